#!/usr/bin/env bash
# Tuning visit: parity tests, then bench lines over the walk split / CTA shape knobs.   Usage: bash tools/gpu_r2b.sh <tag>
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
HSK_TRACE=1 b c2_trace --steps 3 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_trace.err | tail -60 > $OUT/trace_c2.txt; tail -45 $OUT/trace_c2.txt
for sp in 2 3 4 6; do HSK_WALK_SPLIT=$sp b c2_s$sp --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-e2e; done
HSK_WALK_SPLIT=3 HSK_WALK_MIN=8 b c2_s3m8 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-e2e
for th in 512 1024; do for sp in 2 4; do HSK_BIN_THREADS=$th HSK_WALK_SPLIT=$sp b k55_t${th}_s$sp --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity --no-e2e; done; done
b c2 --steps 50 --warmup 5 --no-cpu-baseline
b ext1 --ext 1 --steps 20 --warmup 3 --no-cpu-baseline --no-parity
ls $OUT | wc -l
