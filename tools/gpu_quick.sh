#!/usr/bin/env bash
# Short GPU-box visit: parity tests, two bench lines (no CPU baseline), ncu launch list.
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [notest]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$2" != "notest" ]; then
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
fi
python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python tools/bench_brief.py $OUT/bench.json; tail -3 $OUT/bench.err
python bench.py --workload c2_150Mbp_150bp --no-cpu-baseline > $OUT/bench_150bp.json 2>> $OUT/bench.err; python tools/bench_brief.py $OUT/bench_150bp.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
python tools/launch_summary.py $OUT/launches.csv
