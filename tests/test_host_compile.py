"""Host-side expansion / snapshot compiler: the shortcuts that make it fast must not change what it produces.

  * bare pods with identical content share one validated template (workloads.make_valid_pod_by_pod), but keep their own name,
    node and workload identity (pkg/utils/utils.go:378-463 MakeValidPod runs per pod in the reference);
  * pods created with spec.nodeName form one class per template, not one per node (SCW_NODE_NAME = -3);
  * objects.deep_copy is a structural copy with JSON-round-trip results on JSON-shaped input;
  * the failure message is the same with and without the per-template memo of the node-static reasons.
"""
import copy
import json

import numpy as np

from simon_b200 import objects as O, simulator, synth, workloads
from simon_b200.compiler import compile_cluster, SCW_NODE_NAME


def _pod(name, node=None, cpu="100m", labels=None):
    spec = {"containers": [{"name": "c", "image": "img:v1", "resources": {"requests": {"cpu": cpu, "memory": "64Mi"}}}]}
    if node:
        spec["nodeName"] = node
    return {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": name, "namespace": "ns", "labels": dict(labels or {"app": "a"})}, "spec": spec}


def _node(name, cpu="4"):
    return {"apiVersion": "v1", "kind": "Node", "metadata": {"name": name, "labels": {"kubernetes.io/hostname": name}},
            "status": {"allocatable": {"cpu": cpu, "memory": "8Gi", "pods": "110"}, "capacity": {"cpu": cpu, "memory": "8Gi", "pods": "110"}}}


def test_deep_copy_is_structural_and_matches_a_json_round_trip():
    o = {"a": [1, 2.5, "x", None, True, {"b": []}], "c": {"d": {"e": "f"}}, "t": (1, 2)}
    c = O.deep_copy(o)
    assert c == json.loads(json.dumps(o))
    c["a"][5]["b"].append(1)
    c["c"]["d"]["e"] = "g"
    assert o["a"][5]["b"] == [] and o["c"]["d"]["e"] == "f"


def test_identical_bare_pods_share_a_template_but_keep_their_identity():
    pods = [_pod("p0", "n0"), _pod("p1", "n1"), _pod("p2"), _pod("p3", "n0", cpu="200m"), _pod("p4", "n1", labels={"app": "b"}),
            _pod("p5")]
    before = copy.deepcopy(pods)
    res = O.ResourceTypes(Pods=pods)
    recs = workloads.get_valid_pod_exclude_daemonset(res)
    assert pods == before                                        # inputs untouched
    assert [r.name for r in recs] == ["p0", "p1", "p2", "p3", "p4", "p5"]
    assert [r.node_name for r in recs] == ["n0", "n1", "", "n0", "n1", ""]
    assert [r.workload_name for r in recs] == [r.name for r in recs]
    assert len({r.key() for r in recs}) == 6
    assert recs[0].tmpl is recs[1].tmpl                          # same content, both bound
    assert recs[2].tmpl is recs[5].tmpl                          # same content, both unbound
    assert recs[0].tmpl is not recs[2].tmpl                      # bound and unbound never share (the class records "nodeName set")
    assert recs[3].tmpl is not recs[0].tmpl and recs[4].tmpl is not recs[0].tmpl
    # the shared template went through MakeValidPod like any other
    t = recs[1].tmpl.pod
    assert t["spec"]["schedulerName"] and t["metadata"]["annotations"] == {} and t["status"] == {}
    # and without the memo the same records come out
    solo = [workloads.make_valid_pod_by_pod(p) for p in pods]
    assert [(r.name, r.node_name, r.tmpl.pod["spec"].get("containers")) for r in solo] == \
           [(r.name, r.node_name, r.tmpl.pod["spec"].get("containers")) for r in recs]


def test_pods_created_bound_form_one_class_per_template():
    nodes = [_node(f"n{i}") for i in range(6)]
    pods = [_pod(f"sys-{i}", f"n{i}") for i in range(6)] + [_pod("free-0"), _pod("free-1")]
    p = simulator.plan(O.ResourceTypes(Nodes=nodes, Pods=pods), [])
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    assert c.pods_dims["n_classes"] == 2
    off, blob = c.pods["class_off"], c.pods["class_blob"]
    nn = sorted(int(blob[off[k] + SCW_NODE_NAME]) for k in range(2))
    assert nn == [-3, -1]
    fixed = c.pods["pod_fixed_node"]
    assert [c.node_names[f] for f in fixed[:6]] == [f"n{i}" for i in range(6)] and list(fixed[6:]) == [-1, -1]
    # a name that matches no node is still reported per pod
    p2 = simulator.plan(O.ResourceTypes(Nodes=nodes, Pods=[_pod("lost", "nowhere")]), [])
    c2 = compile_cluster(p2.nodes, p2.pods, p2.ctx)
    assert list(c2.pods["pod_fixed_node"]) == [-2]


def test_failure_message_is_the_same_with_the_static_reason_memo():
    cluster, apps = synth.make_mix(seed_no=5)
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    counts = np.zeros(24, np.uint32)
    counts[3] = 2
    memo = {}
    unbound = [r for r in p.pods if not r.node_name][:40]
    for r in unbound:
        assert simulator.format_fit_error(c, r, counts, static_memo=memo) == simulator.format_fit_error(c, r, counts)
    assert 0 < len(memo) <= len({id(r.tmpl) for r in unbound})
