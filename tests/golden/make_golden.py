"""Regenerates tests/golden/*.json.

There is no Go toolchain in this environment, so these vectors are NOT outputs of the reference binary: they are
(1) hand-derived known answers for the plugin arithmetic (kat_scores.json, derivations inline below) and
(2) the placements of the hand-written 'simple' scenario (tests/fixtures/simple, modelled on the reference's only test,
    pkg/simulator/core_test.go:32-362) on which the two independent CPU restatements (oracle/pyref.py on objects,
    oracle/simon_oracle.c on compiled columns) agree.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
HERE = os.path.dirname(os.path.abspath(__file__))


def simple_case():
    from simon_b200 import objects as O, simulator
    from simon_b200.compiler import compile_cluster
    cluster = O.create_cluster_resource_from_cluster_config(os.path.join(ROOT, "tests/fixtures/simple/cluster"))
    app = O.AppResource("simple", O.get_object_from_yaml_content(
        O.get_yaml_content_from_directory(os.path.join(ROOT, "tests/fixtures/simple/app"))))
    p = simulator.plan(cluster, [app])
    return p, compile_cluster(p.nodes, p.pods, p.ctx)


# Hand-derived plugin scores.  Node: 4 CPU / 8Gi / 110 pods, one running pod requesting 1000m / 2Gi.
# Incoming pod: 500m / 1Gi.
#   LeastAllocated : cpu (4000-1500)*100/4000 = 62 ; mem (8192-3072)*100/8192 = 62 (62.5 truncated) ; (62+62)/2 = 62
#   Balanced       : cpuFraction 1500/4000 = 0.375 ; memFraction 3072/8192 = 0.375 ; (1-0)*100 = 100
#   Simon raw      : cpu share 0.5/(4-0.5) = 0.142857.. ; mem share 1Gi/(8Gi-1Gi) = 0.142857.. ; int64(100*0.142857) = 14
# Second node: 8 CPU / 8Gi, empty.
#   LeastAllocated : cpu (8000-500)*100/8000 = 93 ; mem (8192-1024)*100/8192 = 87 ; (93+87)/2 = 90
#   Balanced       : |0.0625 - 0.125| = 0.0625 ; (1-0.0625)*100 = 93.75 -> 93
#   Simon raw      : cpu 0.5/7.5 = 0.0666 ; mem 1/7 = 0.142857 ; -> 14
# Simon normalised: raw equal on both nodes -> range 0 -> 0 for both.
# TaintToleration 100 / NodeAffinity 0 / InterPodAffinity 0 / PodTopologySpread (no constraints) 100 / extra 1,000,000.
#   total node A = 100 + 62 + 0 + 0 + 2*100 + 100 + 2*0 + 1000000 = 1000462
#   total node B =  93 + 90 + 0 + 0 + 2*100 + 100 + 2*0 + 1000000 = 1000483   -> node B wins
KAT = {
    "nodes": [{"name": "a", "cpu": "4", "memory": "8Gi"}, {"name": "b", "cpu": "8", "memory": "8Gi"}],
    "running": [{"node": "a", "cpu": "1000m", "memory": "2Gi"}],
    "pod": {"cpu": "500m", "memory": "1Gi"},
    "expect": {
        "a": {"ba": 100, "la": 62, "ip": 0, "na": 0, "pts": 100, "tt": 100, "sm": 0, "ex": 1000000, "total": 1000462},
        "b": {"ba": 93, "la": 90, "ip": 0, "na": 0, "pts": 100, "tt": 100, "sm": 0, "ex": 1000000, "total": 1000483},
    },
    "winner": "b",
}


def main():
    import numpy as np
    from oracle.binding import Oracle
    from util import run_pyref
    p, c = simple_case()
    out, _, fc, fp = Oracle(c).schedule()
    ref = run_pyref(p, c)
    assert np.array_equal(out, ref), "the two CPU restatements disagree on the simple scenario"
    placements = [{"workload": [r.tmpl.workload_kind, r.tmpl.workload_namespace, r.workload_name], "ordinal": r.ordinal,
                   "node": c.node_names[n] if n >= 0 else None} for r, n in zip(p.pods, out)]
    json.dump({"node_order": c.node_names, "placements": placements}, open(os.path.join(HERE, "simple_placements.json"), "w"), indent=1)
    json.dump(KAT, open(os.path.join(HERE, "kat_scores.json"), "w"), indent=1)
    print("wrote", len(placements), "placements")


if __name__ == "__main__":
    main()
