#!/bin/bash
# the driver's own sequence on one GPU: smoke, bench (default flags), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-fin}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_${tag}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_${tag}_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_${tag}_reference.json 2> gpurun_out/r2_${tag}_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --e2e-steps 30 > gpurun_out/r2_${tag}_groups.json 2> gpurun_out/r2_${tag}_groups.err; echo "groups rc=$?"
timeout 600 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --e2e-steps 30 > gpurun_out/r2_${tag}_hs.json 2> gpurun_out/r2_${tag}_hs.err; echo "hs rc=$?"
