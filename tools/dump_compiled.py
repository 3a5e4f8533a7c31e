"""Write a compiled cluster (the host buffers of include/simon_gpu.h) to one binary file for native clients.

Format "SIMC1": 8-byte magic, 8 x u32 snapshot dims, 6 x u32 podset dims (declaration order of the structs), then for every
array of simon_snapshot and simon_podset in declaration order: u64 byte length + raw little-endian data.
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
from simon_b200 import abi

SNAP_DIMS = ["n_nodes", "n_scalars", "n_label_words", "n_taint_words", "n_topos", "n_node_classes", "n_log"]
PODS_DIMS = ["n_classes", "n_pods", "n_counters", "n_static_rows", "n_extra_rows", "n_static_sigs"]


def dump(compiled, path: str) -> None:
    with open(path, "wb") as f:
        f.write(b"SIMC1\0\0\0")
        f.write(struct.pack("<8I", *[int(compiled.snap_dims[k]) for k in SNAP_DIMS], 0))
        f.write(struct.pack("<6I", *[int(compiled.pods_dims[k]) for k in PODS_DIMS]))
        for k, dt in abi._SNAP_DTYPES.items():
            a = np.ascontiguousarray(compiled.snap[k], dtype=dt)
            f.write(struct.pack("<Q", a.nbytes))
            f.write(a.tobytes())
        for k, dt in abi._PODS_DTYPES.items():
            a = np.ascontiguousarray(compiled.pods[k] if compiled.pods.get(k) is not None else np.zeros(0), dtype=dt)    # absent: empty blob = NULL
            f.write(struct.pack("<Q", a.nbytes))
            f.write(a.tobytes())


if __name__ == "__main__":
    import argparse
    from simon_b200 import simulator, synth
    from simon_b200.compiler import compile_cluster
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--nodes", type=int, default=10000)
    ap.add_argument("--workloads", type=int, default=1000)
    ap.add_argument("--replicas", type=int, default=100)
    a = ap.parse_args()
    cluster, apps = synth.make_c3(n_nodes=a.nodes, n_workloads=a.workloads, replicas=a.replicas, n_apps=10, seed_no=3)
    p = simulator.plan(cluster, apps)
    dump(compile_cluster(p.nodes, p.pods, p.ctx), a.out)
    print("wrote", a.out, os.path.getsize(a.out), "bytes")
