"""Parity hunt: GPU engine vs C oracle on many seeds of synth.make_mix (placements, scores, failure histograms, state)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from simon_b200.engine import Engine
from util import make_case, run_oracle

a, b = int(sys.argv[1]), int(sys.argv[2])
bad_seeds = []
for seed in range(a, b):
    p, c = make_case("mix", seed_no=seed, n_nodes=20 + (seed % 5) * 25, n_workloads=20 + (seed % 7) * 10, max_replicas=4 + seed % 6)
    (ref, rscore, rfc, rfp), rstate = run_oracle(c)
    with Engine(c, device=0, record_scores=True) as eng:
        out, score, fc, fp = eng.schedule()
        st = eng.state()
    ok = np.array_equal(out, ref) and np.array_equal(fc, rfc) and np.array_equal(fp, rfp)
    sched = ref >= 0
    ok = ok and np.array_equal(score[sched & (rscore > 0)], rscore[sched & (rscore > 0)])
    ok = ok and all(np.array_equal(st[k], rstate[k]) for k in rstate)
    if not ok:
        bad_seeds.append(seed)
        print("MISMATCH seed", seed, np.nonzero(out != ref)[0][:5])
print("seeds", a, b, "mismatching:", bad_seeds)
