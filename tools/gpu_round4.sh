#!/bin/bash
# GPU session: kernel feature variants (compile-time masks) on parity + speed
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
echo "== ubench (trimmed payload)"; timeout 180 tools/ubench_cluster > $OUT/ubench_$TAG.txt 2>&1; grep -E "cs=16 tpb= 320" $OUT/ubench_$TAG.txt | grep -E "argmax|arg-max|allreduce_w<6>"
echo "== parity of the default build (quick subset + kernel paths)"
timeout 900 python -m pytest tests/test_gpu_detail.py tests/test_gpu_parity.py -x -q -m gpu -k "not full_size" 2>&1 | tail -6
for M in 0 2 4 6 7; do
  echo "-- variant mask $M"
  SIMON_GPU_LIB=$PWD/variants/libsimon_m$M.so SIMON_FAST=7 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or parity_with_oracle" 2>&1 | tail -1
  SIMON_GPU_LIB=$PWD/variants/libsimon_m$M.so SIMON_FAST=7 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blocks 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mask $M value', round(d['value']), 'ms', round(d['ms_per_step'],1), d['kernel_stats'])"
done
