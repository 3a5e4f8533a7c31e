#!/bin/bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== moves parity + capacity"
timeout 600 python -m pytest tests/test_moves.py tests/test_capacity.py tests/test_gpu_detail.py -x -q -m gpu 2>&1 | tail -4
echo "== moves bench"
timeout 300 python tools/moves_bench.py 2>&1 | tail -5
SIMON_MOVES_OCC=2 timeout 300 python tools/moves_bench.py --check 0 2>&1 | grep device-resident
echo "== placement variants (compile-time options of the class switch)"
for lib in variants/libsimon_o*.so open-simulator_b200/simon_b200/libsimon_gpu.so; do
  echo "-- $lib"
  SIMON_GPU_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or parity_with_oracle" 2>&1 | tail -1
  SIMON_GPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>$OUT/err.txt | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value', round(d['value']), 'ms', round(d['ms_per_step'],1))
except Exception as e: print('bench failed', t[:200])"
done
