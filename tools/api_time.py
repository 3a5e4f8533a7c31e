"""Where the wall clock of simulator.Simulate() goes on the C3 objects (one untimed call first)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
from simon_b200 import simulator, synth

for rep in range(3):
    cluster, apps = synth.make_c3(n_nodes=10000, n_workloads=200, replicas=500, n_apps=10, seed_no=3)
    t0 = time.perf_counter()
    res = simulator.Simulate(cluster, apps)
    dt = time.perf_counter() - t0
    print(json.dumps({"simulate_s": round(dt, 4), "parts": {k: round(v, 4) for k, v in simulator.Simulate.last_timing.items()},
                      "placed": sum(len(n.Pods) for n in res.NodeStatus), "unscheduled": len(res.UnscheduledPods)}), flush=True)
