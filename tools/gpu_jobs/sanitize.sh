#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck) over small chains: one rank and thread ranks of one device
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-san}
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_sharded.py -m gpu -q -x -k "thread_ranks_match" > gpurun_out/r2_${tag}_memcheck_sharded.log 2>&1; echo "memcheck sharded rc=$?"
timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "philox_chain or horseshoe_chain or groups_chain" > gpurun_out/r2_${tag}_memcheck_single.log 2>&1; echo "memcheck single rc=$?"
tail -5 gpurun_out/r2_${tag}_memcheck_sharded.log; tail -5 gpurun_out/r2_${tag}_memcheck_single.log
