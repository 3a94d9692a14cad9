#!/usr/bin/env bash
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
for ht in 16 8; do
HSK_HOST_THREADS=$ht HSK_TRACE=1 b c2_trace_h$ht --steps 3 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_trace_h$ht.err | grep -v "bin group" | tail -12 > $OUT/trace_c2_h$ht.txt; cat $OUT/trace_c2_h$ht.txt
done
b c2 --steps 50 --warmup 5 --no-cpu-baseline
b c2_150bp --workload c2_150Mbp_150bp --steps 50 --warmup 5 --no-cpu-baseline --no-parity
b k55 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity
