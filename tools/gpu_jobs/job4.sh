#!/bin/bash
# single GPU: parity suite + bench lines of the three samplers (+ optional extra args per line via $BRR_BENCH_EXTRA)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-j4}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${tag}_pytest.log
tail -4 gpurun_out/r2_${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench_groups.json 2> gpurun_out/r2_${tag}_bench_groups.err; echo "groups rc=$?"
timeout 300 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e $BRR_BENCH_EXTRA > gpurun_out/r2_${tag}_bench_hs.json 2> gpurun_out/r2_${tag}_bench_hs.err; echo "hs rc=$?"
