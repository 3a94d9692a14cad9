#!/usr/bin/env bash
# Multi-GPU visit (under gpurun --gpus N): union-over-ranks parity, then bench with both exchange modes.
# Usage: bash tools/gpu_multi.sh <tag> <N> [nocheck]
TAG=${1:-m}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$3" != "nocheck" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_check.py > $OUT/mgpu_check.log 2>&1
grep -v "^W\|^\*\*\*\|OMP_NUM" $OUT/mgpu_check.log | tail -12
fi
for ex in p2p nccl; do
  HSK_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29532 \
      bench.py --gpus $N --steps 10 --warmup 3 2>$OUT/bench_$ex.err > $OUT/bench_$ex.json
  echo "exchange=$ex"; python tools/bench_brief.py $OUT/bench_$ex.json || tail -5 $OUT/bench_$ex.err
done
