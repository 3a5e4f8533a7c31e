"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line (stall samples, instructions)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "" or r[0] == "0":
        continue
    key = (cur.split("/")[-1], int(r[0]))
    a = agg[key]
    a[0] += num(r[4])
    a[1] += num(r[7])
    a[3] = r[1]
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            a[2][h] += num(r[j])
tot = sum(v[0] for v in agg.values())
toti = sum(v[1] for v in agg.values())
print("total samples", tot, "total warp instructions", toti)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{a[6:]}:{b}" for a, b in v[2].most_common(3))
    print(f"{k[0][-16:]}:{k[1]:4d} {v[0]:6d} {100 * v[0] / max(tot, 1):5.1f}% inst={v[1]:8d} [{top}] {v[3].strip()[:80]}")

if len(sys.argv) > 3:
    # coarse regions: "name:lo-hi,..." over simon_kernel.cu ; everything else by file
    regs = []
    for part in sys.argv[3].split(","):
        nm, rng = part.split(":")
        lo, hi = rng.split("-")
        regs.append((nm, int(lo), int(hi)))
    out = collections.defaultdict(lambda: [0, 0])
    for (f, ln), v in agg.items():
        nm = f
        if f == "simon_kernel.cu":
            nm = "cu:other"
            for r in regs:
                if r[1] <= ln <= r[2]:
                    nm = r[0]
        out[nm][0] += v[0]
        out[nm][1] += v[1]
    print("--- regions (samples, warp instructions)")
    for nm, v in sorted(out.items(), key=lambda kv: -kv[1][1]):
        print(f"{nm:28s} samples {v[0]:7d} {100 * v[0] / tot:5.1f}%   inst {v[1]:10d} {100 * v[1] / toti:5.1f}%")
