"""Summarise one kernel of an .ncu-rep (`ncu --set full`) as JSON: duration, DRAM bytes, issue activity, stall mix.

usage: python tools/ncu_summary.py REPORT.ncu-rep [launch_index] > profiles/rNN_ncu_full_summary.json
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--launch-skip", str(skip), "--launch-count", "1", "--page", "raw", "--csv"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def val(name):
    u, v = m[name]
    return float(v.replace(",", "")) * scale.get(u, 1.0)


keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
raw = {k: {"unit": m[k][0], "value": m[k][1]} for k in keep if k in m}
stalls = {k.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(m[k][1].replace(",", ""))
          for k in m if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
out = {
    "report": rep, "launch_index": skip, "kernel": m.get("Kernel Name", ("", ""))[1],
    "gpu_time_duration_s": val("gpu__time_duration.sum"),
    "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
    "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
    "warp_stalls_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10]),
    "raw": raw,
}
print(json.dumps(out, indent=1))
