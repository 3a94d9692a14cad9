#!/usr/bin/env bash
# Full ncu capture of selected kernels of one bench step.  Usage (under gpurun): bash tools/gpu_prof.sh <tag> <kernel regex> [count] [bench args...]
TAG=${1:-p}; RE=${2:-k_bin_count}; CNT=${3:-3}
shift 3 || true
OUT=gpurun_out/$TAG
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 0 -c $CNT -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline "$@" > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log | cut -c1-300
ls -la $OUT
