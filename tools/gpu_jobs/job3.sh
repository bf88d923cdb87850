#!/bin/bash
# round 2, GPU call 3 (8 GPUs): process-per-GPU parity at world 2/4/8 and the fixed-size multi-GPU configs (J2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_j3_smi.txt
timeout 1500 python -m pytest tests/test_sharded.py -m gpu -q --durations=12 -k one_process_per_gpu > gpurun_out/r2_j3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_j3_pytest.log
tail -5 gpurun_out/r2_j3_pytest.log
run() { # name gpus args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$?"
}
for n in 1 2 4; do run r2_j3_c3_groups_n$n $n --sampler groups --total-rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e; done
for n in 1 2 4 8; do run r2_j3_c5_hs_n$n $n --sampler horseshoe --total-rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e; done
run r2_j3_c4_n8 8 --markers 500000 --steps 5 --warmup 3 --burn 5 --no-cpu --no-e2e
run r2_j3_weak_n8 8 --steps 20 --warmup 3 --no-cpu
