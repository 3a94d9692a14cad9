#!/usr/bin/env bash
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
g++ -O3 -pthread tools/hostbench/host_fill_bench.cpp -o /tmp/hfb && /tmp/hfb > $OUT/host_fill_bench.txt 2>&1; grep -E "threads=16|threads= 8|hardware" $OUT/host_fill_bench.txt
b() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
for ht in 16 8 4; do
HSK_HOST_THREADS=$ht HSK_TRACE=1 b c2_trace_h$ht --steps 3 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_trace_h$ht.err | grep -v "bin group" | tail -11 > $OUT/trace_c2_h$ht.txt; cat $OUT/trace_c2_h$ht.txt
done
ls $OUT | wc -l
