"""Parity hunt: GPU engine vs C oracle on many seeds of synth.make_mix (placements, scores, failure histograms, state)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from simon_b200.engine import Engine
from util import run_oracle
from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster


def extra_daemonsets(cluster, seed):
    """0-3 more DaemonSets with random host ports / tolerations / node selectors / sizes (the pinned-pod paths)."""
    rng = np.random.default_rng(seed)
    keys = sorted({k for n in cluster.Nodes for k in ((n.get("metadata") or {}).get("labels") or {}) if "hostname" not in k})
    for k in range(int(rng.integers(0, 4))):
        c = {"name": "c", "image": f"ds{k}:v1", "resources": {"requests": {"cpu": str(rng.choice(["50m", "500m", "3"])), "memory": "64Mi"}}}
        if rng.random() < 0.6:
            port = int(9000 + rng.integers(0, 3))
            c["ports"] = [{"containerPort": port, "hostPort": port}]
        spec = {"containers": [c]}
        if rng.random() < 0.6:
            spec["tolerations"] = [{"operator": "Exists"}]
        if keys and rng.random() < 0.4:
            key = str(rng.choice(keys))
            vals = sorted({str(((n.get("metadata") or {}).get("labels") or {}).get(key)) for n in cluster.Nodes if key in ((n.get("metadata") or {}).get("labels") or {})})
            if vals:
                spec["nodeSelector"] = {key: str(rng.choice(vals))}
        lab = {"ds": f"sweep-{k}"}
        cluster.DaemonSets.append({"apiVersion": "apps/v1", "kind": "DaemonSet", "metadata": {"name": f"sweep-{k}", "namespace": "kube-system"},
                                   "spec": {"selector": {"matchLabels": lab}, "template": {"metadata": {"labels": lab}, "spec": spec}}})


a, b = int(sys.argv[1]), int(sys.argv[2])
bad_seeds = []
n_pinned = 0
for seed in range(a, b):
    cluster, apps = synth.make_mix(seed_no=seed, n_nodes=20 + (seed % 5) * 25, n_workloads=20 + (seed % 7) * 10, max_replicas=4 + seed % 6)
    extra_daemonsets(cluster, seed)
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    n_pinned += int((c.pods["pod_pin_node"] >= 0).sum())
    (ref, rscore, rfc, rfp), rstate = run_oracle(c)
    with Engine(c, device=0, record_scores=True, pin_fast=(seed % 3 != 0)) as eng:      # every third seed: the general path
        out, score, fc, fp = eng.schedule()
        st = eng.state()
    ok = np.array_equal(out, ref) and np.array_equal(fc, rfc) and np.array_equal(fp, rfp)
    sched = ref >= 0
    ok = ok and np.array_equal(score[sched & (rscore > 0)], rscore[sched & (rscore > 0)])
    ok = ok and all(np.array_equal(st[k], rstate[k]) for k in rstate)
    if not ok:
        bad_seeds.append(seed)
        print("MISMATCH seed", seed, np.nonzero(out != ref)[0][:5])
print("seeds", a, b, "pinned pods", n_pinned, "mismatching:", bad_seeds)
