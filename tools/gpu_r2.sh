#!/usr/bin/env bash
# One GPU-box visit of round 2: parity tests, smoke, bench lines (c2, 150 bp, K=55 with both CTA shapes, EXT), reference arm,
# ncu launch list and full capture of the main kernels.   Usage (under gpurun): bash tools/gpu_r2.sh <tag> [quick|noprof]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
b c2 --steps 50 --warmup 5
b c2_150bp --workload c2_150Mbp_150bp --steps 50 --warmup 5 --no-cpu-baseline
b k55 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline
HSK_BIN_THREADS=512 b k55_t512 --k 55 --m 23 --steps 30 --warmup 3 --no-cpu-baseline --no-parity
b ext1 --ext 1 --steps 30 --warmup 3 --no-cpu-baseline
if [ "$2" != "quick" ]; then
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
fi
if [ "$2" != "quick" ] && [ "$2" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_launch.log 2>&1
python tools/launch_summary.py $OUT/launches.csv
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count|k_supermer_count|k_supermer_scatter" -s 0 -c 6 -o $OUT/prof_main -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count" -s 0 -c 2 -o $OUT/prof_k55 -f \
    python bench.py --k 55 --m 23 --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_full_k55.log 2>&1
fi
ls -la $OUT
