#!/usr/bin/env bash
# Multi-GPU visit (under gpurun --gpus N): union-over-ranks parity (engine + C++ API as an MPI job), then bench lines.
# Usage: bash tools/gpu_multi2.sh <tag> <N> [workloads...]
TAG=${1:-m}; N=${2:-2}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 900 $T --master-port 29531 tools/mgpu_check.py > $OUT/mgpu_check.log 2>&1; echo "mgpu_check rc=$?"
grep -E "^PASS|^FAIL|MGPU_CHECK|Error|error" $OUT/mgpu_check.log | tail -24
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q > $OUT/pytest_mgpu.log 2>&1; tail -3 $OUT/pytest_mgpu.log
b() { name=$1; shift; timeout 900 $T --master-port 29532 bench.py --gpus $N "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -8 $OUT/bench_$name.err; }
for w in "$@"; do
  case $w in
    c2) b c2 --steps 20 --warmup 5 ;;
    c2nccl) HSK_EXCHANGE=nccl b c2_nccl --steps 20 --warmup 5 --no-parity ;;
    c3) b c3_share --workload c3_share --steps 5 --warmup 2 ;;
    c4) b c4_share --workload c4_share --steps 5 --warmup 2 ;;
    c5) b c5_share --workload c5_share --steps 5 --warmup 2 ;;
    c3strong) b c3_30Gbp --workload c3_30Gbp --steps 3 --warmup 1 ;;
    ref) timeout 900 $T --master-port 29533 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-300 $OUT/bench_ref.json ;;
  esac
done
ls $OUT | wc -l
