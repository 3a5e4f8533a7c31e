"""Small driver for ncu: one C3 placement pass (reduced pod count so that ncu's replays stay short)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10000)
ap.add_argument("--workloads", type=int, default=1000)
ap.add_argument("--replicas", type=int, default=100)
ap.add_argument("--pods", type=int, default=6000, help="scheduled pods placed in the profiled launch")
ap.add_argument("--cluster-ctas", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--phase-timers", action="store_true", help="use the kernel variants with per-phase clock64 timers")
a = ap.parse_args()
if a.phase_timers:
    os.environ["SIMON_PROFILE"] = "1"

import numpy as np
from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster
from simon_b200.engine import Engine

cluster, apps = synth.make_c3(n_nodes=a.nodes, n_workloads=a.workloads, replicas=a.replicas, n_apps=10, seed_no=3)
p = simulator.plan(cluster, apps)
c = compile_cluster(p.nodes, p.pods, p.ctx)
first = int(np.argmax(c.pods["pod_fixed_node"] == -1))
with Engine(c, device=0, cluster_ctas=a.cluster_ctas, threads_per_cta=a.threads) as eng:
    eng.schedule(0, first, download=False)           # pre-bound pods (accounting only)
    for _ in range(a.reps):
        eng.schedule(first, a.pods, download=False)
        print(f"placed {a.pods} pods in {eng.last_kernel_ms():.3f} ms -> {a.pods / eng.last_kernel_ms() * 1e3:.0f} decisions/s")
    print("stats", eng.stats())
