#!/usr/bin/env bash
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
HSK_TRACE=1 b c2_trace --steps 3 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_trace.err | grep -v "bin group" | tail -24 > $OUT/trace_c2.txt; cat $OUT/trace_c2.txt
b c2 --steps 50 --warmup 5 --no-cpu-baseline
HSK_WALK_SPLIT=1 b c2_s1 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-e2e
HSK_BIN_THREADS=256 b c2_t256 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-e2e
HSK_BIN_THREADS=256 HSK_WALK_SPLIT=1 b c2_t256_s1 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-e2e
b c2_150bp --workload c2_150Mbp_150bp --steps 50 --warmup 5 --no-cpu-baseline --no-parity
b k55 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline
HSK_WALK_SPLIT=1 b k55_s1 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity --no-e2e
b ext1 --ext 1 --steps 20 --warmup 3 --no-cpu-baseline --no-parity
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count" -s 0 -c 1 -o $OUT/prof_c2 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count" -s 0 -c 1 -o $OUT/prof_k55 -f \
    python bench.py --k 55 --m 23 --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_k55.log 2>&1
ls $OUT | wc -l
