#!/bin/bash
# GPU session: full gpu suite, smoke, bench with all blocks, phase timers, profiles
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
echo "== full gpu suite"
timeout 1800 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (all blocks)"
timeout 1500 python bench.py 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value']), d['cpu_baseline']['placements_identical']))
for k in ('c2','daemonsets','batch','capacity_search','move_scoring'):
    b=d.get(k); print(k, json.dumps(b)[:600])
"
tail -3 $OUT/bench_$TAG.err
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
echo "== phase timers"
timeout 600 python tools/profile_run.py --pods 100000 --phase-timers > $OUT/phase_$TAG.txt 2>&1; tail -2 $OUT/phase_$TAG.txt | cut -c1-1300
bash tools/gpu_prof.sh $TAG
