#!/usr/bin/env bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_cxx_api.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo -n "$name: "; python tools/bench_brief.py $OUT/bench_$name.json || tail -5 $OUT/bench_$name.err; }
HSK_TRACE=1 b c2_150bp --workload c2_150Mbp_150bp --steps 5 --warmup 2 --no-cpu-baseline --no-parity; grep "hsk trace" $OUT/bench_c2_150bp.err | grep -v "bin group" | tail -13
for g in 16 64; do HSK_GROUPS=$g b c2_g$g --steps 30 --warmup 3 --no-cpu-baseline --no-parity; done
b c2 --steps 30 --warmup 3 --no-cpu-baseline --no-parity
