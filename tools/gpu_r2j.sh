#!/usr/bin/env bash
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
b2() { name=$1; shift; timeout 900 $T --master-port 29532 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -8 $OUT/bench_$name.err; }
b1() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -8 $OUT/bench_$name.err; }
b2 c4_share_2gpu --workload c4_share --steps 5 --warmup 2
HSK_TRACE=1 b1 c3_30Gbp_1gpu --workload c3_30Gbp --steps 2 --warmup 1 --no-cpu-baseline --no-e2e; grep -E "memory is short" $OUT/bench_c3_30Gbp_1gpu.err | head -3
b2 c3_30Gbp_2gpu --workload c3_30Gbp --steps 2 --warmup 1 --no-e2e
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $OUT/mem.txt
