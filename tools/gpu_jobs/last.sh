#!/bin/bash
# one GPU, the round's last call: the default build through the GPU suite and the default bench line (tag, optional extra bench runs skipped)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-k}
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
unset BRR_LIB
timeout 60 python -m pytest tests -m gpu -x -q > gpurun_out/r2_${tag}_pytest.log 2>&1; el "default pytest rc=$?"; tail -1 gpurun_out/r2_${tag}_pytest.log
timeout 40 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 100 > gpurun_out/r2_${tag}_bench_n1.json 2> gpurun_out/r2_${tag}_bench_n1.err; el "bench rc=$?"
python tools/summ.py gpurun_out/r2_${tag}_bench_n1.json 2>/dev/null
