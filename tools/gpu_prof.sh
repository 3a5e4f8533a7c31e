#!/bin/bash
# Profiling session (B200_PROFILING.md recipe): launch list of the bench command, ncu --set full of the placement kernel and of
# the move kernel.  Numbers printed under ncu are never bench values.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
echo "== launch list (bench, headline only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-blocks > $OUT/${TAG}_bench_under_ncu.log 2>&1
grep -c simon $OUT/${TAG}_launches.csv
echo "== ncu full: placement kernel (100,000 decisions)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:simon_place --launch-skip 0 -c 1 -o $OUT/${TAG}_place_full \
    python tools/profile_run.py --pods 100000 > $OUT/${TAG}_place_full.log 2>&1
tail -2 $OUT/${TAG}_place_full.log
echo "== ncu full: move kernel (1,000,000 moves)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simon_moves_kernel --launch-skip 3 -c 1 -o $OUT/${TAG}_moves_full \
    python tools/moves_bench.py --steps 2 --check 0 > $OUT/${TAG}_moves_full.log 2>&1
tail -2 $OUT/${TAG}_moves_full.log
ls -la $OUT/*.ncu-rep | tail -3
