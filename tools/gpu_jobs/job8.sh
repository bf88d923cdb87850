#!/bin/bash
# round 2, 8 GPUs: process-per-GPU parity at world 8 and the multi-GPU bench lines of the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-m8}
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q --durations=6 -k "one_process_per_gpu and 8-" > gpurun_out/r2_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${tag}_pytest.log
tail -4 gpurun_out/r2_${tag}_pytest.log
run() { # name gpus args...
  name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
}
run r2_${tag}_weak_n8 8 --steps 20 --warmup 3 --no-cpu
run r2_${tag}_c5_hs_n8 8 --sampler horseshoe --total-rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_c4_n8 8 --markers 500000 --steps 5 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_c5_hs_n4 4 --sampler horseshoe --total-rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_c3_groups_n4 4 --sampler groups --total-rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_${tag}_weak_n2 2 --steps 20 --warmup 3 --no-cpu --no-e2e
