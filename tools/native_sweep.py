"""Parity hunt for the native host side on a GPU: simon_host_simulate (C++ expansion / sorts / compiler / result) against the Python
mirror simulator.Simulate on random feature mixes (DaemonSets, GPU share, images, taints, (anti)affinity, spread constraints).

    python tools/native_sweep.py FIRST LAST        # seeds [FIRST, LAST)

Prints one line per mismatch and a summary; exit code 1 on any mismatch.
"""
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-simulator_b200"))
sys.path.insert(0, ROOT)

from simon_b200 import native_host, simulator, synth  # noqa: E402


def one(seed):
    cluster, apps = synth.make_mix(seed_no=seed, n_nodes=30 + 10 * (seed % 5), n_workloads=20 + seed % 15, max_replicas=4 + seed % 4,
                                   with_images=bool(seed % 3 == 0))
    res = simulator.Simulate(copy.deepcopy(cluster), copy.deepcopy(apps))
    nat = native_host.simulate_native(cluster, apps)
    t = nat["templates"]
    keys = []
    for ti, name, o in zip(nat["podTemplate"], nat["podName"], nat["podOrdinal"]):
        tt = t[ti]
        keys.append((tt["kind"], tt["namespace"], name if tt["kind"] == "Pod" else tt["workload"], o))
    ok = nat["nodes"] == [(s.Node.get("metadata") or {}).get("name") for s in res.NodeStatus]
    for i, st in enumerate(res.NodeStatus):
        ok = ok and [keys[k] for k in nat["nodeStatus"][i]] == [r.key() for r in st.Pods]
    ok = ok and [(keys[u["pod"]], u["reason"]) for u in nat["unscheduledPods"]] == [(u.Pod.key(), u.Reason) for u in res.UnscheduledPods]
    return ok, len(keys), len(nat["unscheduledPods"])


def main():
    first, last = int(sys.argv[1]), int(sys.argv[2])
    t0 = time.time()
    bad = pods = fails = 0
    for seed in range(first, last):
        ok, n, f = one(seed)
        pods += n
        fails += f
        if not ok:
            bad += 1
            print(f"seed {seed}: MISMATCH between simon_host_simulate and simulator.Simulate")
    print(f"native_sweep seeds [{first}, {last}): {last - first} clusters, {pods} pods, {fails} unschedulable pods (messages compared), "
          f"{bad} mismatching clusters, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
