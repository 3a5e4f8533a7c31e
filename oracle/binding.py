"""ctypes binding of oracle/libsimon_oracle.so (TEST INFRASTRUCTURE ONLY — see oracle/simon_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from simon_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsimon_oracle.so")
    src = os.path.join(_HERE, "simon_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "simon_gpu.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsimon_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.simon_oracle_create.restype = C.c_void_p
        _LIB.simon_oracle_create.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.simon_oracle_destroy.argtypes = [C.c_void_p]
        _LIB.simon_oracle_reset.argtypes = [C.c_void_p]
        _LIB.simon_oracle_set_active.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        _LIB.simon_oracle_enable_dump.argtypes = [C.c_void_p]
        _LIB.simon_oracle_last_detail.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.simon_oracle_schedule.restype = C.c_int
        _LIB.simon_oracle_schedule.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.simon_oracle_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        _LIB.simon_oracle_moves_score.restype = C.c_int
        _LIB.simon_oracle_moves_score.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.simon_oracle_set_threads.restype = C.c_int
        _LIB.simon_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
    return _LIB


class Oracle:
    def __init__(self, compiled, threads: int = 1):
        self.c = compiled
        self.snap, self.pods, self._keep = abi.marshal(compiled)
        self.h = lib().simon_oracle_create(C.byref(self.snap), C.byref(self.pods))
        if not self.h:
            raise MemoryError("simon_oracle_create failed")
        self.threads = 1
        if threads > 1:
            self.set_threads(threads)

    def set_threads(self, n: int) -> int:
        """Share the per-node loops between n host threads (spin-waiting pool); placements do not depend on n."""
        self.threads = int(lib().simon_oracle_set_threads(self.h, int(n)))
        return self.threads

    def close(self):
        if self.h:
            lib().simon_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        lib().simon_oracle_reset(self.h)

    def set_active(self, nodes=None):
        if nodes is None:
            lib().simon_oracle_set_active(self.h, None, 0)
        else:
            a = np.ascontiguousarray(nodes, dtype=np.uint32)
            lib().simon_oracle_set_active(self.h, a.ctypes.data, len(a))

    def schedule(self, first=0, count=None, max_fail=None):
        P = self.c.pods_dims["n_pods"]
        count = P - first if count is None else count
        max_fail = count if max_fail is None else max_fail
        out_node = np.full(count, -9, np.int32)
        out_score = np.zeros(count, np.int64)
        fail_counts = np.zeros((max(max_fail, 1), abi.N_FAIL_CODES), np.uint32)
        fail_pod = np.zeros(max(max_fail, 1), np.uint32)
        n_fail = C.c_uint32(0)
        rc = lib().simon_oracle_schedule(self.h, first, count, out_node.ctypes.data, out_score.ctypes.data,
                                         fail_counts.ctypes.data, fail_pod.ctypes.data, max_fail, C.byref(n_fail))
        if rc != 0:
            raise RuntimeError(f"simon_oracle_schedule rc={rc}")
        nf = min(n_fail.value, max_fail)
        return out_node, out_score, fail_counts[:nf], fail_pod[:nf]

    def enable_dump(self):
        lib().simon_oracle_enable_dump(self.h)

    def last_detail(self):
        N = self.c.n_nodes
        code = np.zeros(N, np.int32)
        sc = np.zeros((N, 10), np.int64)
        lib().simon_oracle_last_detail(self.h, code.ctypes.data, sc.ctypes.data)
        return code, sc

    def moves_score(self, moves, placement):
        """Candidate moves [(pod, target)] on the oracle's CURRENT state; placement[pod] = node the pod runs on (or < 0)."""
        a = np.ascontiguousarray(moves, dtype=np.uint32).reshape(-1, 2)
        pl = np.ascontiguousarray(placement, dtype=np.int32)
        gain = np.zeros(max(len(a), 1), np.int32)
        code = np.zeros(max(len(a), 1), np.uint32)
        rc = lib().simon_oracle_moves_score(self.h, len(a), a.ctypes.data, pl.ctypes.data, gain.ctypes.data, code.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"simon_oracle_moves_score rc={rc}")
        return gain[:len(a)], code[:len(a)]

    def state(self):
        N = self.c.n_nodes
        arrs = [np.zeros(N, np.int64) for _ in range(5)] + [np.zeros(N, np.int32)]
        lib().simon_oracle_state(self.h, *[a.ctypes.data for a in arrs])
        return dict(zip(["req_mcpu", "req_mem", "req_eph", "nz_mcpu", "nz_mem", "num_pods"], arrs))
