#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== parity (geometries, scenarios)"
timeout 600 python -m pytest tests/test_capacity.py tests/test_gpu_parity.py -x -q -m gpu -k "scenario or capacity or auto or geometries or mixed" 2>&1 | tail -2
echo "== C4 auto geometry"
timeout 600 python tools/capacity_run.py 2>&1 | tail -1 | cut -c1-300
echo "== batch scaling"
timeout 900 python tools/batch_scale.py --geoms 0x0,8x320 --counts 14 2>&1 | tail -3
