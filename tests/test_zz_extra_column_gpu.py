"""ImageLocality / NodePreferAvoidPods ("extra" column, simon_podset.extra_score) on the GPU engine.

Added after the round's GPU minutes were spent (no earlier test input had a non-empty extra column), so this is the first
GPU run of that gather; the file sorts last so that it cannot mask the rest of the suite under `pytest -x`.
Expected numbers: hand-derived in tests/golden/kat_plugins.py.
"""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["image_locality", "prefer_avoid_pods"])
def test_engine_extra_column_kats(name):
    from simon_b200.engine import Engine
    from test_golden import _kat_cluster, _kat_expected, _plugin_kats
    kat = _plugin_kats()[name]
    p, c, n_in = _kat_cluster(kat)
    assert c.pods_dims["n_extra_rows"] > 0
    winners, scores = _kat_expected(kat, n_in)
    with Engine(c, device=0, record_scores=True) as eng:
        out, score, _, _ = eng.schedule()
    assert [c.node_names[n] if n >= 0 else None for n in out[-n_in:]] == winners
    assert [int(x) for x in score[-n_in:]] == scores


@pytest.mark.parametrize("seed", [500, 503, 511, 520])
def test_engine_matches_oracle_with_node_images(seed):
    import numpy as np
    from simon_b200.engine import Engine
    from util import make_case, run_oracle
    p, c = make_case("mix", seed_no=seed, n_nodes=20 + (seed % 5) * 20, n_workloads=20 + (seed % 7) * 8, with_images=True)
    assert c.pods_dims["n_extra_rows"] > 0
    (ref, rscore, rfc, _), _ = run_oracle(c)
    with Engine(c, device=0, record_scores=True) as eng:
        out, score, fc, _ = eng.schedule()
    np.testing.assert_array_equal(out, ref)
    sched = ref >= 0
    np.testing.assert_array_equal(score[sched & (rscore > 0)], rscore[sched & (rscore > 0)])
    np.testing.assert_array_equal(fc, rfc)
