#!/usr/bin/env python
"""bench.py — pod-placement decisions/s on the 10k-node synthetic cluster (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA engine (through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference path

A "step" is one pass of the hot path over one batch: the whole ordered pod list of workload C3
(10,000 nodes x 100,000 pods, full predicate set) placed from the empty cluster state.
  value  : decisions/s, inputs resident in HBM, device-timed (CUDA events on the engine's stream), max over ranks
  e2e    : same metric through the C-ABI calls a caller makes with HOST buffers:
           simon_snapshot_upload + simon_pods_upload + simon_schedule(out_node -> host), wall clock
  N > 1  : a single scenario does not shard (each placement depends on the previous one) -> N independent
           replicas (distinct seeds), one per GPU, no data-path collective; scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "pod-placement decisions/sec on 10k-node synthetic cluster"
ALGO_BYTES_PER_NODE_DECISION = 180      # SURVEY.md section 8(d), C3 full predicate set


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(args, seed_no):
    from simon_b200 import simulator, synth
    from simon_b200.compiler import compile_cluster
    t0 = time.time()
    cluster, apps = synth.make_c3(n_nodes=args.nodes, n_workloads=args.workloads, replicas=args.replicas,
                                  n_apps=10, seed_no=seed_no)
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    return p, c, time.time() - t0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.th = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads(args, c=None, first_sched=0) -> int:
    """Host threads for the CPU port: --cpu-threads, or the count that is fastest on a short sample of this workload
    (the per-node loops of one cycle are short, so more threads than that only add fork/join cost)."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if args.cpu_threads > 0:
        return max(1, min(64, args.cpu_threads))
    if c is None or ncpu <= 1:
        return 1
    from oracle.binding import Oracle
    best_t, best_v = 1, 0.0
    n = min(1500, c.pods_dims["n_pods"] - first_sched)
    for t in (1, 2, 4, 8, 16, 32, 64):
        if t > ncpu:
            break
        o = Oracle(c, threads=t)
        o.schedule(0, first_sched)
        t0 = time.perf_counter()
        o.schedule(first_sched, n)
        v = n / (time.perf_counter() - t0)
        o.close()
        if v > best_v:
            best_t, best_v = t, v
    return best_t


def run_reference(args):
    """The reference's CPU path, restated (oracle/simon_oracle.c) — there is no Go toolchain to build the original."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import Oracle
    p, c, _ = build_workload(args, 3)
    P = c.pods_dims["n_pods"]
    first_sched = int(np.argmax(c.pods["pod_fixed_node"] == -1))
    sample = min(args.cpu_sample, P - first_sched)
    o = Oracle(c, threads=cpu_threads(args, c, first_sched))
    times = []
    for step in range(args.warmup + args.steps):
        o.reset()
        o.schedule(0, first_sched)                 # pre-bound pods: accounting only, not timed
        t0 = time.perf_counter()
        o.schedule(first_sched, sample)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = sample * len(times) / tot
    cb = {"value": value, "unit": "decisions/s", "cores": o.threads, "kind": "port",
          "sample": f"first {sample} scheduled pods of the {P}-pod list per step (oracle/simon_oracle.c, per-node loops shared "
                    f"between {o.threads} host threads, the fastest count on this box)"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "decisions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": workload_config(args, c), "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, c):
    return {"workload": f"C3: {args.nodes} nodes x {c.pods_dims['n_pods']} pods ({args.workloads} workloads x {args.replicas}), "
                        "full predicate set (taints, nodeSelector/affinity, inter-pod (anti)affinity, topology spread), seed 0x51514D4F4E03",
            "nodes": args.nodes, "pods": int(c.pods_dims["n_pods"]), "pod_classes": int(c.pods_dims["n_classes"]),
            "parallelism": "1 scenario per GPU (replicas)", "l2": "flushed between timed steps (256 MiB write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--nodes", type=int, default=10000)
    ap.add_argument("--workloads", type=int, default=1000)
    ap.add_argument("--replicas", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=30000)
    ap.add_argument("--cluster-ctas", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU port (0 = all cores, capped at 64)")
    ap.add_argument("--no-batch", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gpu" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl gpu needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from simon_b200.engine import Engine
    p, c, host_s = build_workload(args, 3 + rank)          # replicas: distinct seeds per rank
    P = int(c.pods_dims["n_pods"])
    D = int((c.pods["pod_fixed_node"] == -1).sum())        # decisions = pods that go through filter+score (pre-bound pods excluded)
    eng = Engine(c, device=local, cluster_ctas=args.cluster_ctas, threads_per_cta=args.threads)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.replay(1)
    sampler = ClockSampler(local)
    launches0 = eng.launch_count()
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    ms_total = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ms_total += eng.replay(1)
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * D * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers -> C ABI -> host results ----
    from simon_b200 import abi
    import ctypes as C
    from simon_b200.engine import lib
    L = lib()
    h2d = sum(a.nbytes for a in eng._keep)
    out_node = np.empty(P, np.int32)
    n_fail = C.c_uint32(0)
    e2e_times = []
    for step in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        assert L.simon_snapshot_upload(eng.h, C.byref(eng.snap)) == 0
        assert L.simon_pods_upload(eng.h, C.byref(eng.pods)) == 0
        assert L.simon_schedule(eng.h, 0, P, out_node.ctypes.data, None, None, None, 0, C.byref(n_fail)) == 0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step >= 2:
            e2e_times.append(dt)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * D * len(e2e_times) / float(te.item())
    placed = int((out_node >= 0).sum())

    # ---- batch of independent what-if replicas on ONE GPU: one 16-CTA cluster each.  A cluster lives inside one GPC; on this
    # part 7 such clusters are co-resident (tools/batch_scale.py: 1..7 scenarios take the same time, the 8th starts a second wave) ----
    batch = None
    if not args.no_batch and world == 1:
        nb = 7
        act = np.arange(int(c.n_nodes), dtype=np.uint32)
        eng.run_scenarios([act] * nb)                                   # warm-up
        flush.zero_()
        torch.cuda.synchronize()
        res, _ = eng.run_scenarios([act] * nb)
        bms = eng.last_kernel_ms()
        same = all(r["n_scheduled"] == placed - (P - D) or r["n_scheduled"] == placed for r in res)
        batch = {"scenarios": nb, "value": nb * D / (bms * 1e-3), "unit": "decisions/s", "ms": bms, "identical_counts": bool(same),
                 "note": "simon_scenarios_run: 7 independent replicas of the same workload placed concurrently on one GPU "
                         "(the capacity-planning batch shape); aggregate, not the headline value"}

    if rank == 0:
        peak, peak_src = peaks()
        per_gpu_dps = D * args.steps / (ms_total * 1e-3)
        achieved = per_gpu_dps * ALGO_BYTES_PER_NODE_DECISION * args.nodes / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")))
            traffic = prof["dram_bytes_per_launch"]      # dram__bytes_read.sum + dram__bytes_write.sum of one --set full capture
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_NODE_DECISION * args.nodes * D,
                "note": f"algorithmic bytes = {ALGO_BYTES_PER_NODE_DECISION} B/node-decision x {args.nodes} nodes x {D} decisions per launch; "
                        "the snapshot is cluster-resident in shared memory, so DRAM traffic << algorithmic bytes"}
        cb = None
        if not args.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N = 1 only
            from oracle.binding import Oracle
            first_sched = int(np.argmax(c.pods["pod_fixed_node"] == -1))
            o = Oracle(c, threads=cpu_threads(args, c, first_sched))
            sample = min(args.cpu_sample, P - first_sched)
            o.schedule(0, first_sched)
            t0 = time.perf_counter()
            ref_nodes, _, _, _ = o.schedule(first_sched, sample)
            dt = time.perf_counter() - t0
            agree = bool(np.array_equal(ref_nodes, out_node[first_sched:first_sched + sample]))
            cb = {"value": sample / dt, "unit": "decisions/s", "cores": o.threads, "kind": "port",
                  "sample": f"first {sample} scheduled pods of the same pod list (oracle/simon_oracle.c, per-node loops shared "
                            f"between {o.threads} host threads, the fastest count on this box)",
                  "placements_identical_on_sample": agree}
        line = {"metric": METRIC, "value": value, "unit": "decisions/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic", "config": workload_config(args, c),
                "roofline": roof, "cpu_baseline": cb, "batch": batch,
                "e2e": {"value": e2e_value, "unit": "decisions/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(4 * P + 8),
                        "path": "simon_snapshot_upload + simon_pods_upload + simon_schedule(host out_node), wall clock"},
                "gpu_launches": int(launches), "clocks": clocks, "decisions_per_step": D, "prebound_pods_per_step": P - D, "placed": placed, "unschedulable": int((out_node == -1).sum()),
                "wall_s_timed_region": wall, "host_compile_s": host_s}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
