#!/bin/bash
# GPU session: kernel feature variants (compile-time masks) on parity + speed
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
for lib in variants/libsimon_m*.so; do
  M=$(basename $lib .so | sed 's/libsimon_m//')
  echo "-- variant mask $M"
  SIMON_GPU_LIB=$PWD/$lib SIMON_FAST=7 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detail.py tests/test_golden.py tests/test_capacity.py -q -m gpu -k "not full_size" 2>&1 | grep -E "FAILED|passed|failed|mismatch|Mismatch|first" | head -12
  SIMON_GPU_LIB=$PWD/$lib SIMON_FAST=7 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>/dev/null | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('mask $M value', round(d['value']), 'ms', round(d['ms_per_step'],1), d['kernel_stats'])
except Exception as e: print('mask $M bench failed', t[:200])"
done
