#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for lib in variants/libsimon_o*.so; do
  echo "-- $lib"
  SIMON_GPU_LIB=$PWD/$lib timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detail.py tests/test_capacity.py -x -q -m gpu -k "not full_size" 2>&1 | tail -1
  for rep in 1 2; do
  SIMON_GPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>$OUT/err.txt | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value', round(d['value']), 'ms', round(d['ms_per_step'],1))
except Exception as e: print('bench failed', t[:200])"
  done
done
