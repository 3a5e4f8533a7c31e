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
