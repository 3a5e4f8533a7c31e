"""C5 timing: 1M candidate moves on a 10k-node live snapshot (device-resident replay + one end-to-end call)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
from simon_b200 import moves as M, simulator, synth
from simon_b200.compiler import compile_cluster
from simon_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10000)
ap.add_argument("--running", type=int, default=300000)
ap.add_argument("--moves", type=int, default=1000000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--check", type=int, default=20000, help="moves compared with the oracle")
a = ap.parse_args()
t0 = time.time()
cluster, apps = synth.make_c3(n_nodes=a.nodes, n_workloads=max(10, a.nodes // 10), replicas=10, n_apps=10, seed_no=3)
p = simulator.plan(cluster, apps)
c = compile_cluster(p.nodes, p.pods, p.ctx)
live = synth.make_c5(c, a.running)
print(f"snapshot: {live.n_nodes} nodes, {live.pods_dims['n_pods']} running pods, built in {time.time() - t0:.1f} s", flush=True)
with Engine(live, device=0) as eng:
    out = eng.schedule()[0]
    print("import of the running pods:", round(eng.last_kernel_ms(), 1), "ms")
    mv = M.sample_moves(len(out), live.n_nodes, a.moves, seed=11, placement=out)
    eng.moves_upload(mv, 0)
    eng.moves_replay(3)
    ms = eng.moves_replay(a.steps) / a.steps
    print(f"device-resident: {ms * 1e3:.1f} us per {a.moves} moves -> {a.moves / ms * 1e3 / 1e9:.2f} G moves/s; 280 B/move -> {280 * a.moves / ms * 1e3 / 1e9:.0f} GB/s algorithmic")
    t1 = time.perf_counter()
    eng.moves_upload(mv, 0)
    r = eng.moves_run(k=16, want_arrays=True, want_per_pod=True)
    dt = time.perf_counter() - t1
    print(f"end to end (upload + run + top-16 + downloads): {dt * 1e3:.2f} ms -> {a.moves / dt / 1e6:.1f} M moves/s; feasible {r['n_feasible']}; best {M.decode_key(r['best_key'])}; top3 {r['topk'][:3]}")
    if a.check:
        from oracle.binding import Oracle
        o = Oracle(live)
        o.schedule()
        g, cd = o.moves_score(mv[:a.check], out)
        print("oracle check on", a.check, "moves:", bool(np.array_equal(g, r["gain"][:a.check]) and np.array_equal(cd, r["code"][:a.check])))
