#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out; mkdir -p $OUT
echo "== pins + parity + capacity + moves"
timeout 1500 python -m pytest tests/test_daemonset_pins.py tests/test_gpu_parity.py tests/test_capacity.py tests/test_moves.py tests/test_native_client.py tests/test_config1.py -x -q -m gpu 2>&1 | tail -8
echo "== ds bench"
timeout 900 python tools/ds_bench.py 2>&1 | tee $OUT/ds_bench_$TAG.txt | tail -5
echo "== headline only"
timeout 900 python bench.py --no-blocks --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']))"
