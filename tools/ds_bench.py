"""DaemonSet-heavy cluster: 10,000 nodes, C3 workloads plus k DaemonSets (k x 10,000 pinned pods): placement time, with the oracle
compared on a prefix."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster
from simon_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10000)
ap.add_argument("--ds", type=int, default=5)
ap.add_argument("--check", type=int, default=30000)
a = ap.parse_args()
cluster, apps = synth.make_c3(n_nodes=a.nodes, n_workloads=200, replicas=max(1, a.nodes // 20), n_apps=10, seed_no=3)
for k in range(a.ds):
    cluster.DaemonSets.append({"apiVersion": "apps/v1", "kind": "DaemonSet", "metadata": {"name": f"agent-{k}", "namespace": "kube-system"},
        "spec": {"selector": {"matchLabels": {"ds": f"agent-{k}"}}, "template": {"metadata": {"labels": {"ds": f"agent-{k}"}},
                 "spec": {"tolerations": [{"operator": "Exists"}],
                          "containers": [{"name": "c", "image": f"agent:{k}", "resources": {"requests": {"cpu": "50m", "memory": "64Mi"}},
                                          "ports": [{"containerPort": 9100 + k, "hostPort": 9100 + k}]}]}}}})
t0 = time.perf_counter(); p = simulator.plan(cluster, apps); t1 = time.perf_counter()
c = compile_cluster(p.nodes, p.pods, p.ctx); t2 = time.perf_counter()
n_ds = sum(1 for r in p.pods if r.tmpl.workload_kind == "DaemonSet")
print(f"pods {len(p.pods)} (DaemonSet pods {n_ds}), classes {c.pods_dims['n_classes']}, plan {t1 - t0:.2f} s, compile {t2 - t1:.2f} s", flush=True)
with Engine(c, device=0) as eng:
    eng.schedule()
    eng.reset()
    t0 = time.perf_counter(); out, _s, fc, fp = eng.schedule(); dt = time.perf_counter() - t0
    st = eng.stats()
print(f"schedule: {dt * 1e3:.1f} ms wall, kernel {eng.last_kernel_ms() if hasattr(eng, 'last_kernel_ms') else -1:.1f} ms; placed {(out >= 0).sum()}, unschedulable {(out == -1).sum()}; "
      f"class switches {st.get('class_switches')}", flush=True)
if a.check:
    from oracle.binding import Oracle
    o = Oracle(c, threads=16)
    a.check = min(a.check, len(out))
    ref = o.schedule(0, a.check)[0]
    o.close()
    print("oracle prefix", a.check, "identical:", bool(np.array_equal(ref, out[:a.check])), flush=True)
