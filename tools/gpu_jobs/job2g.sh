#!/bin/bash
# 2 GPUs: process-per-GPU parity at world 2 and multi-GPU bench lines of the current code (tag, then "groups" to add the Groups line)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-g2}
timeout 600 python -m pytest tests/test_sharded.py -m gpu -q -k "one_process_per_gpu" > gpurun_out/r2_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${tag}_pytest.log
tail -3 gpurun_out/r2_${tag}_pytest.log
run() { name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_${tag}_weak_n2 2 --steps 20 --warmup 3 --no-cpu --no-e2e
run r2_${tag}_c5_hs_n2 2 --sampler horseshoe --total-rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
[ "$2" = groups ] && run r2_${tag}_c3_groups_n2 2 --sampler groups --total-rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e
python tools/summ.py gpurun_out/r2_${tag}_*.json 2>/dev/null
