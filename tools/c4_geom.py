"""C4 (256 scenarios x 5,000 pending pods x 2,000-2,032 nodes): run time of the batch for a few cluster geometries."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
from simon_b200 import capacity, synth
from simon_b200.engine import Engine

cluster, apps, specs = synth.make_c4(n_nodes=2000, n_workloads=50, replicas=100)
ss = capacity.build_scenarios(cluster, apps, specs, list(range(1, 33)))
lists = [sc.nodes for sc in ss.scenarios]
for cs, tpb in ((0, 0), (1, 320), (2, 320), (2, 256), (4, 256), (4, 320), (4, 192), (8, 256), (8, 128), (16, 128)):
    try:
        with Engine(ss.compiled, device=0, cluster_ctas=cs, threads_per_cta=tpb) as eng:
            eng.run_scenarios(lists)
            t0 = time.perf_counter()
            res, _ = eng.run_scenarios(lists)
            dt = time.perf_counter() - t0
            print(f"geom {cs}x{tpb}: kernel {eng.last_kernel_ms():.1f} ms, call {dt * 1e3:.1f} ms, unscheduled of first/last {res[0]['n_unscheduled']}/{res[-1]['n_unscheduled']}", flush=True)
    except Exception as e:
        print(f"geom {cs}x{tpb}: {str(e)[:100]}", flush=True)
