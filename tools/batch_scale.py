"""Run time of 1..n concurrent scenarios (one thread-block cluster each) on one GPU, for a few cluster geometries."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster
from simon_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--geoms", default="16x320,8x320")
ap.add_argument("--counts", default="1,7,8,14,16,18")
a = ap.parse_args()
cluster, apps = synth.make_c3(n_nodes=10000, n_workloads=1000, replicas=100, n_apps=10, seed_no=3)
p = simulator.plan(cluster, apps)
c = compile_cluster(p.nodes, p.pods, p.ctx)
D = int((c.pods["pod_fixed_node"] == -1).sum())
act = np.arange(c.n_nodes, dtype=np.uint32)
for g in a.geoms.split(","):
    cs, tpb = [int(x) for x in g.split("x")]
    try:
        with Engine(c, device=0, cluster_ctas=cs, threads_per_cta=tpb) as eng:
            for nb in [int(x) for x in a.counts.split(",")]:
                eng.run_scenarios([act] * nb)
                res, _ = eng.run_scenarios([act] * nb)
                ms = eng.last_kernel_ms()
                el = [round(r["elapsed_ms"]) for r in res]
                print(f"geom {cs}x{tpb}: {nb} scenarios: {ms:.1f} ms -> {round(nb * D / ms * 1e3)} decisions/s; per-scenario elapsed min {min(el)} max {max(el)}", flush=True)
    except Exception as e:
        print(f"geom {cs}x{tpb}: {e}", flush=True)
