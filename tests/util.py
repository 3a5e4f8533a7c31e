"""Shared helpers for the tests: build a plan + compiled cluster from a synthetic config."""
import numpy as np

from simon_b200 import simulator, synth
from simon_b200.compiler import compile_cluster


def make_case(kind="c3", **kw):
    cluster, apps = {"c3": synth.make_c3, "c2": synth.make_c2, "mix": synth.make_mix}[kind](**kw)
    p = simulator.plan(cluster, apps)
    c = compile_cluster(p.nodes, p.pods, p.ctx)
    return p, c


def run_oracle(c):
    from oracle.binding import Oracle
    o = Oracle(c)
    out = o.schedule()
    st = o.state()
    o.close()
    return out, st


def run_pyref(p, c):
    from oracle.pyref import PyRef
    ref = PyRef(c.node_objs, services=p.ctx.services, replicasets=p.ctx.replicasets, statefulsets=p.ctx.statefulsets,
                creation_order=c.node_orig_index)
    return np.array(ref.run([x.tmpl.pod for x in p.pods], [x.node_name for x in p.pods]), dtype=np.int32)


class CompiledArrays:
    """A compiled cluster restored from stored columns (what Engine / Oracle consume: the arrays of include/simon_gpu.h)."""

    def __init__(self, npz):
        import json
        self.snap = {k[6:]: npz[k] for k in npz.files if k.startswith("snap__")}
        self.pods = {k[6:]: npz[k] for k in npz.files if k.startswith("pods__")}
        self.snap_dims = json.loads(bytes(npz["snap_dims"]).decode())
        self.pods_dims = json.loads(bytes(npz["pods_dims"]).decode())
        self.facts = json.loads(bytes(npz["facts"]).decode()) if "facts" in npz.files else {}
        self.n_nodes = int(self.snap_dims["n_nodes"])
        self.node_names = self.facts.get("node_names", [str(i) for i in range(self.n_nodes)])
