#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench (ours + reference arm), ncu launch list, full ncu capture of the main kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; python tools/bench_brief.py $OUT/bench.json; tail -3 $OUT/bench.err
python bench.py --workload c2_150Mbp_150bp --no-cpu-baseline > $OUT/bench_150bp.json 2>> $OUT/bench.err; python tools/bench_brief.py $OUT/bench_150bp.json
if [ "$2" != "quick" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cat $OUT/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
python tools/launch_summary.py $OUT/launches.csv
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count|k_supermer_count|k_supermer_scatter" -s 0 -c 10 -o $OUT/prof_main -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
