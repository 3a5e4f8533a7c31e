#!/bin/bash
# A/B of library builds on one box: headline only, alternating
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for lib in "" "$@"; do
  if [ -z "$lib" ]; then unset SIMON_GPU_LIB; name=shipped; else export SIMON_GPU_LIB=$PWD/$lib; name=$lib; fi
  timeout 600 python bench.py --no-blocks --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2))"
done; done
