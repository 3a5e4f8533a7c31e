#!/bin/bash
# GPU session: full gpu suite, bench line with all blocks, kernel variants, phase timers, batch residency sweep
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
echo "== full gpu suite"
timeout 2700 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (all blocks)"
timeout 1500 python bench.py 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value']), d['cpu_baseline']['placements_identical']))
for k in ('c2','batch','capacity_search','move_scoring'):
    b=d.get(k); print(k, json.dumps(b)[:700])
"
tail -3 $OUT/bench_$TAG.err
echo "== kernel variants"
for F in 0 2 3; do
  SIMON_FAST=$F timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('SIMON_FAST=$F value', round(d['value']), 'ms', round(d['ms_per_step'],1))"
done
echo "== moves occupancy variants"
for O in 2 4; do SIMON_MOVES_OCC=$O timeout 300 python tools/moves_bench.py --check 0 2>&1 | grep device-resident; done
echo "== phase timers"
SIMON_FAST=2 timeout 600 python tools/profile_run.py --pods 100000 --phase-timers > $OUT/phase_$TAG.txt 2>&1; tail -2 $OUT/phase_$TAG.txt | cut -c1-1400
echo "== batch residency"
timeout 900 python tools/batch_scale.py --geoms 16x320,8x320 --counts 1,7,8,14,16,18 2>&1 | tee $OUT/batch_$TAG.txt | tail -14
