#!/usr/bin/env bash
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
b c2 --steps 50 --warmup 5 --no-cpu-baseline --no-parity
b k55 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity --no-e2e
HSK_TRACE=1 b c3_30Gbp --workload c3_30Gbp --steps 3 --warmup 1 --no-cpu-baseline --no-e2e; grep -E "memory is short" $OUT/bench_c3_30Gbp.err | head -3
