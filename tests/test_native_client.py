"""The C ABI is callable from native code without Python or torch: open-simulator_b200/native/simon_client.cpp."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLIENT = os.path.join(ROOT, "open-simulator_b200", "native", "simon_client")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _fnv1a(out: np.ndarray) -> str:
    h = 1469598103934665603
    for b in np.ascontiguousarray(out, dtype="<i4").tobytes():
        h ^= b
        h = (h * 1099511628211) & ((1 << 64) - 1)
    return f"{h:016x}"


def _dump(tmp_path, **kw):
    from dump_compiled import dump
    from util import make_case
    p, c = make_case("c3", **kw)
    path = os.path.join(tmp_path, "cluster.simc")
    dump(c, path)
    return c, path


def _build():
    import __graft_entry__ as g
    g.build_client()            # never rebuilds libsimon_gpu.so (it may be loaded in this process)
    assert os.path.exists(CLIENT)


def test_native_client_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    _build()
    c, path = _dump(str(tmp_path), n_nodes=30, n_workloads=6, replicas=3, n_apps=1, seed_no=5)
    r = subprocess.run([CLIENT, path], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 2 and "no CPU path" in r.stderr      # refuses, does not fall back


@pytest.mark.gpu
def test_native_client_matches_oracle(tmp_path):
    from util import run_oracle
    _build()
    c, path = _dump(str(tmp_path), n_nodes=400, n_workloads=60, replicas=10, n_apps=2, seed_no=6)
    (ref, _, _, _), _ = run_oracle(c)
    r = subprocess.run([CLIENT, path, "0", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["pods"] == len(ref)
    assert res["placed"] == int((ref >= 0).sum()) and res["unschedulable"] == int((ref == -1).sum())
    assert res["fnv1a"] == _fnv1a(ref)
    assert res["replay_steps"] == 2 and res["replay_ms_per_step"] > 0


def _request(tmp_path, **kw):
    from simon_b200 import native_host, synth
    cluster, apps = synth.make_c3(**kw)
    path = os.path.join(tmp_path, "request.json")
    with open(path, "wb") as fh:
        fh.write(native_host.request_json(cluster, apps))
    return cluster, apps, path


def test_native_client_simulate_fails_loudly_without_a_gpu(tmp_path):
    """simon_client --simulate: objects (JSON) -> simon_host_simulate -> SimulateResult, no Python / torch on the path."""
    import torch
    _build()
    _c, _a, path = _request(str(tmp_path), n_nodes=30, n_workloads=6, replicas=3, n_apps=1, seed_no=5)
    r = subprocess.run([CLIENT, "--simulate", path], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 2 and "no CPU path" in r.stderr
    bad = os.path.join(str(tmp_path), "bad.json")
    with open(bad, "w") as fh:
        fh.write("{\"cluster\": ")
    r = subprocess.run([CLIENT, "--simulate", bad], capture_output=True, text=True)
    assert r.returncode == 3 and "JSON" in r.stderr


@pytest.mark.gpu
def test_native_client_simulate_matches_the_python_api(tmp_path):
    import copy
    from simon_b200 import simulator
    _build()
    cluster, apps, path = _request(str(tmp_path), n_nodes=400, n_workloads=60, replicas=10, n_apps=2, seed_no=6)
    out_path = os.path.join(str(tmp_path), "result.json")
    r = subprocess.run([CLIENT, "--simulate", path, "0", out_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    nat = json.load(open(out_path))
    res = simulator.Simulate(copy.deepcopy(cluster), copy.deepcopy(apps))
    assert summary["pods"] == len(nat["podNode"]) and summary["fnv1a_pod_node"] == _fnv1a(np.array(nat["podNode"], dtype=np.int32))
    assert summary["placed"] == sum(len(s.Pods) for s in res.NodeStatus) and summary["unschedulable"] == len(res.UnscheduledPods)
    assert [u["reason"] for u in nat["unscheduledPods"]] == [u.Reason for u in res.UnscheduledPods]
    assert [len(x) for x in nat["nodeStatus"]] == [len(s.Pods) for s in res.NodeStatus]
