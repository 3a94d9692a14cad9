mkdir -p gpurun_out/r2rehearsal
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2rehearsal/bench_2gpu.json 2> gpurun_out/r2rehearsal/bench_2gpu.err; echo "ours rc=$?"; python tools/bench_brief.py gpurun_out/r2rehearsal/bench_2gpu.json || tail -5 gpurun_out/r2rehearsal/bench_2gpu.err
timeout 200 $T --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2rehearsal/bench_ref_2gpu.json 2> gpurun_out/r2rehearsal/bench_ref_2gpu.err; echo "ref rc=$?"; cut -c1-160 gpurun_out/r2rehearsal/bench_ref_2gpu.json
