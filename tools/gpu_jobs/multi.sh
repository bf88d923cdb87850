#!/bin/bash
# multi-GPU evidence: one-process-per-GPU parity test + torchrun bench at N GPUs; usage: tools/gpu_jobs/multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.log 2>&1
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q -p no:cacheprovider -k one_process_per_gpu > gpurun_out/multi_test_$N.log 2>&1; tail -5 gpurun_out/multi_test_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
