#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== capacity / scenario parity"
timeout 600 python -m pytest tests/test_capacity.py tests/test_gpu_detail.py tests/test_gpu_parity.py -x -q -m gpu -k "scenario or capacity or auto or geometries" 2>&1 | tail -2
echo "== C4 auto geometry"
timeout 600 python tools/capacity_run.py 2>&1 | tail -1 | cut -c1-400
echo "== batch scaling auto"
timeout 900 python tools/batch_scale.py --geoms 0x0,8x320 --counts 7,14 2>&1 | tail -5
