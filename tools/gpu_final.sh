#!/usr/bin/env bash
# Final visit of a round: what the driver runs (GPU tests, smoke, both bench arms) + the ncu evidence for profiles/.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; python tools/bench_brief.py $OUT/bench_c2.json || tail -5 $OUT/bench_c2.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_launch.log 2>&1
python tools/launch_summary.py $OUT/launches.csv
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count|k_supermer_count|k_supermer_scatter" -s 0 -c 3 -o $OUT/prof_c2 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count" -s 0 -c 1 -o $OUT/prof_k55 -f \
    python bench.py --k 55 --m 23 --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_k55.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bin_count" -s 0 -c 1 -o $OUT/prof_ext1 -f \
    python bench.py --ext 1 --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-e2e > $OUT/ncu_ext1.log 2>&1
ls $OUT | wc -l
