"""Geometry sweep for ONE small scenario: decisions/s for every (CTAs per cluster, threads per CTA) that fits, against the automatic choice.
    python tools/geom_sweep.py"""
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

CASES = [("c2 1000 nodes fit-only", lambda: synth.make_c2()),
         ("c3-shape 600 nodes", lambda: synth.make_c3(n_nodes=600, n_workloads=60, replicas=100, n_apps=4, seed_no=23)),
         ("c3-shape 2500 nodes", lambda: synth.make_c3(n_nodes=2500, n_workloads=100, replicas=100, n_apps=4, seed_no=24)),
         ("c3-shape 5000 nodes", lambda: synth.make_c3(n_nodes=5000, n_workloads=200, replicas=100, n_apps=4, seed_no=25))]
for name, mk in CASES:
    cluster, apps = mk()
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    D = int((c.pods["pod_fixed_node"] == -1).sum())
    ref = None
    rows = []
    for cs in (0, 1, 2, 4, 8, 16):
        for thr in ((0,) if cs == 0 else (64, 128, 256, 320)):
            try:
                with Engine(c, device=0, cluster_ctas=cs, threads_per_cta=thr) as eng:
                    out = eng.schedule()[0]
                    eng.replay(1)
                    ms = eng.replay(3) / 3
                if ref is None:
                    ref = out
                rows.append((round(D / ms * 1e3), cs, thr, bool(np.array_equal(out, ref))))
            except Exception as e:      # noqa: BLE001
                pass
    rows.sort(reverse=True)
    print(json.dumps({"case": name, "nodes": c.n_nodes, "decisions": D, "auto": [r for r in rows if r[1] == 0], "best5": rows[:5],
                      "all_same_placements": all(r[3] for r in rows)}))
