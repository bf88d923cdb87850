#!/bin/bash
# quick single-GPU bench lines (no tests): tag as $1, extra bench args in $BRR_BENCH_EXTRA
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-q}
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench_groups.json 2> gpurun_out/r2_${tag}_bench_groups.err; echo "groups rc=$?"
timeout 300 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench_hs.json 2> gpurun_out/r2_${tag}_bench_hs.err; echo "hs rc=$?"
