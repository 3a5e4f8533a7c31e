#!/bin/bash
# GPU session: speed of experiment libraries (headline only)
OUT=gpurun_out; mkdir -p $OUT
for lib in "$@"; do
  echo "-- $lib"
  SIMON_GPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>$OUT/err.txt | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'launches', d['gpu_launches'])
except Exception as e: print('bench failed', t[:200])"
  tail -2 $OUT/err.txt
done
