"""Single-CTA clusters (CTA-local reductions, no DSMEM exchange) against the default geometry on small clusters.
    python tools/solo_bench.py
For each case: decisions/s with cluster_ctas = default / 1 / 2, placements compared with the oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-simulator_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from simon_b200 import simulator, synth  # noqa: E402
from simon_b200.compiler import compile_cluster  # noqa: E402
from simon_b200.engine import Engine  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

CASES = [("c2 1000x10000 fit-only", lambda: synth.make_c2()),
         ("c3-shape 1000 nodes", lambda: synth.make_c3(n_nodes=1000, n_workloads=100, replicas=100, n_apps=4, seed_no=21)),
         ("c3-shape 1200 nodes", lambda: synth.make_c3(n_nodes=1200, n_workloads=120, replicas=100, n_apps=4, seed_no=22)),
         ("c3-shape 600 nodes", lambda: synth.make_c3(n_nodes=600, n_workloads=60, replicas=100, n_apps=4, seed_no=23))]
for name, mk in CASES:
    cluster, apps = mk()
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    ref = Oracle(c).schedule()[0]
    D = int((c.pods["pod_fixed_node"] == -1).sum())
    row = {"case": name, "nodes": c.n_nodes, "decisions": D}
    for cs, thr in ((0, 0), (1, 0), (1, 320), (2, 0)):
        try:
            with Engine(c, device=0, cluster_ctas=cs, threads_per_cta=thr) as eng:
                out = eng.schedule()[0]
                eng.replay(2)
                ms = eng.replay(3) / 3
            row[f"cs{cs}_t{thr}"] = {"decisions_per_s": round(D / ms * 1e3), "ms": round(ms, 3), "identical_to_oracle": bool(np.array_equal(out, ref))}
        except Exception as e:      # noqa: BLE001
            row[f"cs{cs}_t{thr}"] = {"error": str(e)[:120]}
    print(json.dumps(row))
