#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
echo "== ds bench, every pod compared with the oracle"
timeout 1200 python tools/ds_bench.py --check 1000000 2>&1 | tee $OUT/ds_bench_full_$TAG.txt | tail -4
echo "== mix sweep 4000..4240"
timeout 1500 python tools/gpu_mix_sweep.py 4000 4240 2>&1 | tee $OUT/mix_sweep2_$TAG.txt | tail -4
