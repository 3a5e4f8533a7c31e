"""Go math routines the placement path depends on, restated bit-for-bit (host side).

PodTopologySpread weights use math.Log(float64(size+2))
(vendor/k8s.io/kubernetes/pkg/scheduler/framework/plugins/podtopologyspread/scoring.go:279-281).
Go's math.Log on amd64 (pure Go `log`, and the amd64 assembly that transcribes it) is the FreeBSD
e_log.c algorithm; it is NOT guaranteed to equal glibc's log(), so the engine never calls libm:
the host builds the table below and uploads it (simon_snapshot.log_table).
"""
from __future__ import annotations

import math

_LN2HI = 6.93147180369123816490e-01
_LN2LO = 1.90821492927058770002e-10
_L1 = 6.666666666666735130e-01
_L2 = 3.999999999940941908e-01
_L3 = 2.857142874366239149e-01
_L4 = 2.222219843214978396e-01
_L5 = 1.818357216161805012e-01
_L6 = 1.531383769920937332e-01
_L7 = 1.479819860511658591e-01
_SQRT2 = 1.41421356237309504880168872420969808


def go_log(x: float) -> float:
    """Go src/math/log.go `log` (operation order preserved; Python floats are IEEE binary64, unfused)."""
    if x != x or x == math.inf:
        return x
    if x < 0:
        return math.nan
    if x == 0:
        return -math.inf
    f1, ki = math.frexp(x)
    if f1 < _SQRT2 / 2:
        f1 *= 2
        ki -= 1
    f = f1 - 1
    k = float(ki)
    s = f / (2 + f)
    s2 = s * s
    s4 = s2 * s2
    t1 = s2 * (_L1 + s4 * (_L3 + s4 * (_L5 + s4 * _L7)))
    t2 = s4 * (_L2 + s4 * (_L4 + s4 * _L6))
    R = t1 + t2
    hfsq = 0.5 * f * f
    return k * _LN2HI - ((hfsq - (s * (hfsq + R) + k * _LN2LO)) - f)


def log_table(n: int):
    """table[i] = go_log(float(i)) for i in [0, n); table[0] is unused (set to 0)."""
    import numpy as np
    t = np.zeros(n, dtype=np.float64)
    for i in range(1, n):
        t[i] = go_log(float(i))
    return t
