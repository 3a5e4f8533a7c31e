import os, sys
ROOT = "/root/repo"
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster
from simon_b200.engine import Engine
cluster, apps = synth.make_c3(n_nodes=10000, n_workloads=1000, replicas=100, n_apps=10, seed_no=3)
p = simulator.plan(cluster, apps)
c = compile_cluster(p.nodes, p.pods, p.ctx)
D = int((c.pods["pod_fixed_node"] == -1).sum())
act = np.arange(c.n_nodes, dtype=np.uint32)
with Engine(c, device=0) as eng:
    for nb in (1, 2, 4, 6, 7, 8):
        eng.run_scenarios([act] * nb)
        res, _ = eng.run_scenarios([act] * nb)
        ms = eng.last_kernel_ms()
        print(nb, "scenarios:", round(ms, 1), "ms ->", round(nb * D / ms * 1e3), "decisions/s; per-scenario elapsed", [round(r["elapsed_ms"]) for r in res])
