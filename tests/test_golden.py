"""Oracle pinned against the golden vectors of tests/golden (hand-derived KATs + the 'simple' scenario), and the
facts the reference's own test pins (pkg/simulator/core_test.go:364-591): zero unschedulable pods and per-workload
pod counts."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _simple():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import simple_case
    return simple_case()


def _kat_case():
    from simon_b200 import simulator
    from simon_b200.compiler import compile_cluster
    from simon_b200.objects import AppResource, ResourceTypes
    kat = json.load(open(os.path.join(HERE, "golden", "kat_scores.json")))
    cluster = ResourceTypes()
    for n in kat["nodes"]:
        cluster.Nodes.append({"kind": "Node", "metadata": {"name": n["name"], "labels": {"kubernetes.io/hostname": n["name"]}},
                              "status": {"allocatable": {"cpu": n["cpu"], "memory": n["memory"], "pods": "110"}}})
    for i, r in enumerate(kat["running"]):
        cluster.Pods.append({"kind": "Pod", "metadata": {"name": f"run-{i}", "namespace": "default"},
                             "spec": {"nodeName": r["node"], "containers": [{"name": "c", "image": "x", "resources": {"requests": {"cpu": r["cpu"], "memory": r["memory"]}}}]}})
    app = AppResource("kat", ResourceTypes())
    app.Resource.Pods.append({"kind": "Pod", "metadata": {"name": "incoming", "namespace": "default"},
                              "spec": {"containers": [{"name": "c", "image": "x", "resources": {"requests": kat["pod"]}}]}})
    p = simulator.plan(cluster, [app])
    return kat, p, compile_cluster(p.nodes, p.pods, p.ctx)


def test_oracle_reproduces_hand_derived_plugin_scores():
    from oracle.binding import Oracle
    kat, p, c = _kat_case()
    o = Oracle(c)
    o.enable_dump()
    out, score, _, _ = o.schedule()
    code, sc = o.last_detail()
    names = ["ba", "la", "ip", "na", "pts", "tt", "sm", "ex", "total"]
    for node, exp in kat["expect"].items():
        i = c.node_index(node)
        got = dict(zip(names, [int(x) for x in sc[i][:9]]))
        assert got == exp, (node, got, exp)
    assert c.node_names[out[-1]] == kat["winner"]
    assert int(score[-1]) == kat["expect"][kat["winner"]]["total"]


def test_simple_scenario_matches_golden_and_reference_pinned_facts():
    from util import run_oracle, run_pyref
    p, c = _simple()
    (out, _, fc, fp), _ = run_oracle(c)
    gold = json.load(open(os.path.join(HERE, "golden", "simple_placements.json")))
    assert c.node_names == gold["node_order"]
    got = [c.node_names[n] if n >= 0 else None for n in out]
    assert got == [g["node"] for g in gold["placements"]]
    np.testing.assert_array_equal(out, run_pyref(p, c))
    # what checkResult asserts in the reference: no unschedulable pod, pod count per workload
    assert (out == -1).sum() == 0
    counts = {}
    for r in p.pods:
        k = (r.tmpl.workload_kind, r.workload_name)
        counts[k] = counts.get(k, 0) + 1
    assert counts[("ReplicaSet", "busybox-deploy")] == 4
    assert counts[("StatefulSet", "busybox-sts")] == 4
    assert counts[("ReplicaSet", "calico-rs")] == 2
    assert counts[("Job", "pi")] == 3
    assert counts[("DaemonSet", "busybox-ds")] == 3          # master-1 is tainted and the DaemonSet does not tolerate it
    assert counts[("DaemonSet", "kube-proxy")] == 4
    assert counts[("DaemonSet", "worker-agent")] == 1
    assert sum(1 for r in p.pods if r.tmpl.workload_kind == "Pod") == 5


def test_unschedulable_reason_string_format():
    """FitError.Error() histogram format (generic_scheduler.go:72-90) + Simulator.update wrapper (simulator.go:465)."""
    from oracle.binding import Oracle
    from simon_b200 import simulator
    from simon_b200.compiler import compile_cluster
    from simon_b200.objects import AppResource, ResourceTypes
    cluster = ResourceTypes()
    for i, taint in enumerate([None, {"key": "dedicated", "value": "gpu", "effect": "NoSchedule"}]):
        n = {"kind": "Node", "metadata": {"name": f"n{i}", "labels": {"kubernetes.io/hostname": f"n{i}"}},
             "spec": {"taints": [taint]} if taint else {}, "status": {"allocatable": {"cpu": "2", "memory": "4Gi", "pods": "110"}}}
        cluster.Nodes.append(n)
    app = AppResource("a", ResourceTypes())
    app.Resource.Pods.append({"kind": "Pod", "metadata": {"name": "big", "namespace": "default"},
                              "spec": {"containers": [{"name": "c", "image": "x", "resources": {"requests": {"cpu": "3", "memory": "8Gi"}}}]}})
    p = simulator.plan(cluster, [app])
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    out, _, fc, fp = Oracle(c).schedule()
    assert list(out) == [-1]
    msg = simulator.format_fit_error(c, p.pods[0], fc[0])
    assert msg == ("failed to schedule pod (default/big): Unschedulable: 0/2 nodes are available: "
                   "1 Insufficient cpu, 1 Insufficient memory, 1 node(s) had taint {dedicated: gpu}, that the pod didn't tolerate.")


@pytest.mark.gpu
def test_simulate_api_on_simple_scenario_gpu():
    """simulator.Simulate() end to end on the GPU engine: same placements as the golden fixture."""
    from simon_b200 import objects as O, simulator
    root = os.path.dirname(HERE)
    cluster = O.create_cluster_resource_from_cluster_config(os.path.join(root, "tests/fixtures/simple/cluster"))
    app = O.AppResource("simple", O.get_object_from_yaml_content(O.get_yaml_content_from_directory(os.path.join(root, "tests/fixtures/simple/app"))))
    res = simulator.Simulate(cluster, [app], simulator.DisablePTerm(True))
    assert res.UnscheduledPods == []
    gold = json.load(open(os.path.join(HERE, "golden", "simple_placements.json")))
    want = {}
    for g in gold["placements"]:
        want[(tuple(g["workload"]), g["ordinal"])] = g["node"]
    got = {}
    for st in res.NodeStatus:
        for rec in st.Pods:
            got[((rec.tmpl.workload_kind, rec.tmpl.workload_namespace, rec.workload_name), rec.ordinal)] = st.Node["metadata"]["name"]
    assert got == want


def _plugin_kats():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import kat_plugins
    return kat_plugins.CASES


KAT_NAMES = ["taints_and_preferred_node_affinity", "default_topology_spread", "preferred_inter_pod_affinity",
             "hard_spread_single_survivor", "simon_packs_onto_the_smallest_node", "gpu_share_per_device_fit", "self_affinity_series",
             "image_locality", "prefer_avoid_pods"]
KAT_GPU_LATE = {"gpu_share_per_device_fit", "image_locality", "prefer_avoid_pods"}      # run on the GPU by tests/test_zz_*.py


def _kat_cluster(kat):
    from simon_b200 import simulator
    from simon_b200.compiler import compile_cluster
    from simon_b200.objects import AppResource, ResourceTypes
    cluster = ResourceTypes()
    cluster.Nodes.extend(kat["nodes"])
    cluster.Pods.extend(kat["running"])
    cluster.Services.extend(kat["services"])
    app = AppResource("kat", ResourceTypes())
    pods = kat["pod"] if isinstance(kat["pod"], list) else [kat["pod"]]
    app.Resource.Pods.extend(pods)
    apps = [app]
    if kat.get("pod2"):                      # a second app: placed after the first one (pkg/simulator/core.go:105-116)
        app2 = AppResource("kat2", ResourceTypes())
        app2.Resource.Pods.extend(kat["pod2"])
        apps.append(app2)
    p = simulator.plan(cluster, apps)
    return p, compile_cluster(p.nodes, p.pods, p.ctx), len(pods) + len(kat.get("pod2") or [])


def _kat_expected(kat, n_in):
    winners = kat.get("winners", [kat.get("winner")])
    scores = kat.get("scores", [kat.get("winner_score", kat["expect"].get(kat.get("winner"), {}).get("total"))])
    assert len(winners) == n_in and len(scores) == n_in
    return winners, scores


@pytest.mark.parametrize("name", KAT_NAMES)
def test_oracle_reproduces_hand_derived_plugin_kats(name):
    """tests/golden/kat_plugins.py: expected numbers derived by hand from the reference's formulas (file:line there)."""
    from oracle.binding import Oracle
    from oracle.pyref import PyRef
    kat = _plugin_kats()[name]
    p, c, n_in = _kat_cluster(kat)
    winners, scores = _kat_expected(kat, n_in)
    o = Oracle(c)
    o.enable_dump()
    out, score, _, _ = o.schedule()
    code, sc = o.last_detail()
    names = ["ba", "la", "ip", "na", "pts", "tt", "sm", "ex", "total"]
    for node, exp in kat["expect"].items():          # per-plugin scores of the (single) incoming pod
        i = c.node_index(node)
        got = dict(zip(names, [int(x) for x in sc[i][:9]]))
        assert got == exp, (node, got, exp)
    for node, cde in kat.get("infeasible", {}).items():
        assert int(code[c.node_index(node)]) == cde
    got_w = [c.node_names[n] if n >= 0 else None for n in out[-n_in:]]
    assert got_w == winners
    assert [int(x) for x in score[-n_in:]] == scores
    # the independent object-level restatement picks the same nodes
    ref = PyRef(c.node_objs, services=p.ctx.services, replicasets=p.ctx.replicasets, statefulsets=p.ctx.statefulsets,
                creation_order=c.node_orig_index)
    py = ref.run([x.tmpl.pod for x in p.pods], [x.node_name for x in p.pods])
    assert [c.node_names[n] if n >= 0 else None for n in py[-n_in:]] == winners


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in KAT_NAMES if n not in KAT_GPU_LATE])
def test_engine_reproduces_hand_derived_plugin_kats(name):
    """The CUDA engine picks the hand-derived winners with the hand-derived totals (tests/golden/kat_plugins.py)."""
    from simon_b200.engine import Engine
    kat = _plugin_kats()[name]
    p, c, n_in = _kat_cluster(kat)
    winners, scores = _kat_expected(kat, n_in)
    with Engine(c, device=0, record_scores=True) as eng:
        out, score, _, _ = eng.schedule()
    assert [c.node_names[n] if n >= 0 else None for n in out[-n_in:]] == winners
    assert [int(x) for x in score[-n_in:]] == scores
