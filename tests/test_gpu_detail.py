"""Per-node parity (not only the winner): the engine's total score and filter verdict of EVERY node for chosen pods
(simon_debug_set_dump_pod / simon_debug_dump_read) against the oracle's per-node detail, on ordinary clusters and on
adversarial numeric inputs:
  * BalancedAllocation values whose exact (1 - |cpuFraction - memFraction|) * 100 is an integer or within an ulp of one
    (the kernel's reciprocal fast path must fall back to the correctly rounded division there),
  * LeastAllocated with (capacity - requested) * 100 far above 2^52 (petabyte-scale memory),
plus the GPU-share Reserve outputs, the extended state download and repeated scenario runs on one context.
"""
import os
import sys
from fractions import Fraction

import numpy as np
import pytest

from util import make_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _detail_oracle(c, pod):
    from oracle.binding import Oracle
    o = Oracle(c)
    o.enable_dump()
    out, _, _, _ = o.schedule(0, pod + 1)
    code, sc = o.last_detail()
    o.close()
    return out, np.array([int(x[8]) for x in sc], np.int64), np.array(code, np.int32)


def _check_pod(c, pod):
    from simon_b200.engine import Engine
    ref_out, ref_total, ref_code = _detail_oracle(c, pod)
    with Engine(c, device=0) as eng:
        out, tot, code = eng.dump_pod(pod)
    np.testing.assert_array_equal(out, ref_out)
    feas = ref_code == 0
    np.testing.assert_array_equal(code == 0, feas, err_msg=f"pod {pod}: feasible sets differ")
    if feas.sum() > 1:                      # F == 1: "just use it", nothing is scored (generic_scheduler.go:159-165)
        bad = np.nonzero(feas & (tot != ref_total))[0]
        assert len(bad) == 0, f"pod {pod}: totals differ on nodes {bad[:8]}: gpu {tot[bad[:8]]} oracle {ref_total[bad[:8]]}"
    return int(feas.sum())


@pytest.mark.parametrize("kind,kw,pods", [
    ("c3", dict(n_nodes=300, n_workloads=60, replicas=10, n_apps=2, seed_no=21), [40, 333, 599]),
    ("mix", dict(seed_no=103, n_nodes=95, n_workloads=40, max_replicas=7), [30, 90, 150]),
    ("c2", dict(n_nodes=200, n_workloads=20, replicas=20, seed_no=2), [5, 250]),
])
def test_per_node_totals_match_oracle(kind, kw, pods):
    p, c = make_case(kind, **kw)
    first = int(np.argmax(c.pods["pod_fixed_node"] == -1))
    for q in pods:
        pod = min(first + q, c.pods_dims["n_pods"] - 1)
        _check_pod(c, pod)


def _ba_adversarial_caps(req_c, req_m, n_want):
    """Node capacities (mCPU, bytes) for which the exact BalancedAllocation value of a pod requesting (req_c mCPU, req_m bytes)
    on an empty node is an integer, or whose float64 evaluation lands within 1e-12 of one (both sides)."""
    out = []
    for capc in range(req_c + 100, 64000, 50):
        for mult in (2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 100, 125, 128, 160, 200, 250, 256):
            capm = req_m * mult
            exact = (1 - abs(Fraction(req_c, capc) - Fraction(req_m, capm))) * 100
            x = (1.0 - abs(req_c / capc - req_m / capm)) * 100.0
            near = abs(x - round(x)) < 1e-12
            if exact.denominator == 1 or near:
                out.append((capc, capm))
                if len(out) >= n_want:
                    return out
    return out


def test_balanced_allocation_integer_boundaries():
    """Every node is an adversarial (capacity) pair for the pod's request: the exact score sits on / next to an integer, where
    a reciprocal-multiply evaluation and the correctly rounded quotient of Go can truncate differently."""
    from simon_b200 import simulator
    from simon_b200.compiler import compile_cluster
    from simon_b200.objects import AppResource, ResourceTypes
    n_tot = 0
    for req_c, req_mi in ((300, 768), (700, 1536), (1000, 1000), (250, 3 * 1024)):
        req_m = req_mi * 1024 * 1024
        caps = _ba_adversarial_caps(req_c, req_m, 400)
        assert len(caps) >= 50
        cluster = ResourceTypes()
        for i, (capc, capm) in enumerate(caps):
            cluster.Nodes.append({"kind": "Node", "metadata": {"name": f"n{i:04d}", "labels": {"kubernetes.io/hostname": f"n{i:04d}"}},
                                  "spec": {}, "status": {"allocatable": {"cpu": f"{capc}m", "memory": str(capm), "pods": "110"}}})
        app = AppResource("a", ResourceTypes())
        app.Resource.Pods.append({"kind": "Pod", "metadata": {"name": "p", "namespace": "default"},
                                  "spec": {"containers": [{"name": "c", "image": "x", "resources": {"requests": {"cpu": f"{req_c}m", "memory": str(req_m)}}}]}})
        p = simulator.plan(cluster, [app])
        c = compile_cluster(p.nodes, p.pods, p.ctx)
        n_tot += _check_pod(c, 0)
    assert n_tot > 200


def test_least_allocated_beyond_2_pow_52():
    """(capacity - requested) * 100 up to ~2^62: exact integer division must survive the double-reciprocal estimate."""
    from simon_b200 import simulator
    from simon_b200.compiler import compile_cluster
    from simon_b200.objects import AppResource, ResourceTypes
    cluster = ResourceTypes()
    rng = np.random.RandomState(5)
    for i in range(256):
        capm = int(rng.randint(1, 1 << 30)) * (1 << 25) + int(rng.randint(0, 1 << 20))       # up to 2^55 bytes, odd values too
        capc = int(rng.randint(1000, 2000000))
        cluster.Nodes.append({"kind": "Node", "metadata": {"name": f"n{i:04d}", "labels": {"kubernetes.io/hostname": f"n{i:04d}"}},
                              "spec": {}, "status": {"allocatable": {"cpu": f"{capc}m", "memory": str(capm), "pods": "110"}}})
    app = AppResource("a", ResourceTypes())
    for j, (rc, rm) in enumerate(((500, (1 << 40) + 12345), (999, (1 << 52) + 1), (100, 3))):
        app.Resource.Pods.append({"kind": "Pod", "metadata": {"name": f"p{j}", "namespace": "default"},
                                  "spec": {"containers": [{"name": "c", "image": "x", "resources": {"requests": {"cpu": f"{rc}m", "memory": str(rm)}}}]}})
    p = simulator.plan(cluster, [app])
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    for pod in range(3):
        _check_pod(c, pod)


def test_gpu_share_reserve_outputs():
    """Device ids per pod (the alibabacloud.com/gpu-index annotation Reserve writes) and per-device memory in use, KAT F:
    p3 (3Gi) -> g2 device 0, p2 (6Gi) -> g2 device 1, a (6Gi) -> g1 device 0 (tightest fit among two free 8Gi devices is the
    first), b unschedulable, c (8Gi) -> g1 device 1."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import kat_plugins
    from simon_b200.engine import Engine
    from test_golden import _kat_cluster
    kat = kat_plugins.CASES["gpu_share_per_device_fit"]
    p, c, n_in = _kat_cluster(kat)
    with Engine(c, device=0) as eng:
        out, _, _, _ = eng.schedule()
        slots = eng.gpu_slots()
        ext = eng.state_ext()
    names = [r.name for r in p.pods]
    by = {nm: slots[i] for i, nm in enumerate(names)}
    assert by["r1"] == []                       # pre-bound: never reaches Reserve
    assert by["p3"] == [0] and by["p2"] == [1] and by["a"] == [0] and by["b"] == [] and by["c"] == [1]
    gi = 1 << 30
    g1, g2 = c.node_index("g1"), c.node_index("g2")
    assert list(ext["gpu_used"][:2, g1]) == [6 * gi, 8 * gi]
    assert list(ext["gpu_used"][:2, g2]) == [3 * gi, 6 * gi]


def test_scenarios_twice_on_one_context_with_different_node_lists():
    """The cgo pattern of a capacity search: one ctx, simon_scenarios_run called again with other node subsets.  Every call
    starts from an empty state (no class summary or own-score cache of the previous call may survive)."""
    from oracle.binding import Oracle
    from simon_b200.engine import Engine
    p, c = make_case("c3", n_nodes=200, n_workloads=30, replicas=8, n_apps=1, seed_no=31)
    N = c.n_nodes
    rng = np.random.RandomState(3)
    lists_a = [np.arange(N, dtype=np.uint32), np.arange(0, N, 2, dtype=np.uint32)]
    lists_b = [np.sort(rng.choice(N, 120, replace=False)).astype(np.uint32), np.arange(N - 1, -1, -1, dtype=np.uint32)[:150].copy(),
               np.arange(1, N, 2, dtype=np.uint32)]
    o = Oracle(c)
    with Engine(c, device=0) as eng:
        for lists in (lists_a, lists_b, lists_a):
            res, nodes = eng.run_scenarios(lists, want_nodes=True)
            for k, act in enumerate(lists):
                o.reset()
                o.set_active(act)
                ref, _, _, _ = o.schedule()
                np.testing.assert_array_equal(nodes[k], ref, err_msg=f"scenario {k}")
                st = o.state()
                assert res[k]["req_mcpu"] == int(st["req_mcpu"][act.astype(np.int64)].sum())
                assert res[k]["n_unscheduled"] == int((ref == -1).sum())
    o.close()


def test_gpu_share_failure_names_the_nodes():
    """UnscheduledPod.Reason of a pod the Open-Gpu-Share filter rejects everywhere: one "Node:<name>" reason per node
    (open-gpu-share.go:66-79 + FitError.Error, generic_scheduler.go:72-90), sorted like every other reason."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import kat_plugins
    from simon_b200 import simulator
    from simon_b200.objects import AppResource, ResourceTypes
    kat = kat_plugins.CASES["gpu_share_per_device_fit"]
    cluster = ResourceTypes()
    cluster.Nodes.extend(kat["nodes"])
    cluster.Pods.extend(kat["running"])
    a1, a2 = AppResource("kat", ResourceTypes()), AppResource("kat2", ResourceTypes())
    a1.Resource.Pods.extend(kat["pod"])
    a2.Resource.Pods.extend(kat["pod2"])
    res = simulator.Simulate(cluster, [a1, a2], simulator.DisablePTerm(True))
    assert [u.Pod.name for u in res.UnscheduledPods] == ["b"]
    assert res.UnscheduledPods[0].Reason == ("failed to schedule pod (default/b): Unschedulable: 0/2 nodes are available: "
                                             "1 Node:g1, 1 Node:g2.")
