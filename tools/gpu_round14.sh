#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
echo "== mix sweep with random DaemonSets (both pinned-pod paths)"
timeout 1500 python tools/gpu_mix_sweep.py 3000 3160 2>&1 | tee $OUT/mix_sweep_$TAG.txt | tail -6
