#!/bin/bash
# round 2, GPU call 2: whole GPU suite (dense columns, boundary test, sharded), quick bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2_j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_j2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_j2_bench.json 2> gpurun_out/r2_j2_bench.err; echo "bench rc=$?"
tail -30 gpurun_out/r2_j2_pytest.log
