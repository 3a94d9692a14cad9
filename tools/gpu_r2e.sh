#!/usr/bin/env bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag > $OUT/thp.txt 2>&1; cat $OUT/thp.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
HSK_TRACE=1 b c2_trace --steps 3 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_trace.err | grep -v "bin group" | tail -10 > $OUT/trace_c2.txt; cat $OUT/trace_c2.txt
b c2 --steps 50 --warmup 5 --no-cpu-baseline
b c2_150bp --workload c2_150Mbp_150bp --steps 50 --warmup 5 --no-cpu-baseline --no-parity
b k55 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline
HSK_BIN_THREADS=512 b k55_t512 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity --no-e2e
b ext1 --ext 1 --steps 20 --warmup 3 --no-cpu-baseline --no-parity
timeout 600 python tools/polya_check.py 1.0 > $OUT/polya.log 2>&1; tail -3 $OUT/polya.log
b c3_30Gbp --workload c3_30Gbp --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
HSK_TRACE=1 timeout 900 python tools/big_check.py 3.75 > $OUT/big_check.log 2>&1; grep -v "hsk trace" $OUT/big_check.log | tail -4
ls $OUT | wc -l
