#!/bin/bash
# GPU session: co-resident scenario batches with cluster sizes that are not powers of two
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
echo "== Simulate() timing"; timeout 300 python tools/api_time.py 2>&1 | tail -3
timeout 1500 python tools/batch_scale.py --geoms 8x320,9x320,10x320,10x256,12x320,6x320,7x320 --counts 12,14,15,16,18,21 2>&1 | tee $OUT/batch_geoms_$TAG.txt | tail -45
