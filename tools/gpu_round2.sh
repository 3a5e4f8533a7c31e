#!/bin/bash
# GPU session: micro-benchmark of the reduction variants, move-scoring parity + timing
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
echo "== ubench"; timeout 180 tools/ubench_cluster > $OUT/ubench_$TAG.txt 2>&1; grep -E "cs=16 tpb= 320|cs= 1 tpb= 320" $OUT/ubench_$TAG.txt | grep -E "polling|allreduce_w<6>|self"
echo "== moves parity"
timeout 900 python -m pytest tests/test_moves.py tests/test_gpu_detail.py -x -q -m gpu 2>&1 | tail -15
echo "== moves bench"
timeout 600 python tools/moves_bench.py 2>&1 | tee $OUT/moves_$TAG.txt | tail -8
