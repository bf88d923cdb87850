#!/bin/bash
# 4 GPUs: multi-GPU bench lines of the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-g4}
run() { name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_${tag}_c5_hs_n4 4 --sampler horseshoe --total-rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_c3_groups_n4 4 --sampler groups --total-rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_weak_n4 4 --steps 20 --warmup 3 --no-cpu --no-e2e
