#!/bin/bash
# A/B of a library variant: parity suites under the variant, then alternating headline runs
V=${1:-variants/libsimon_o253.so}
echo "== parity under $V"
SIMON_GPU_LIB=$PWD/$V timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_daemonset_pins.py tests/test_capacity.py tests/test_zz_rare_paths_gpu.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_ab.sh $V
