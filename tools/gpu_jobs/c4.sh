#!/bin/bash
# BASELINE configs[3] shape: N = 400,000 x M = 500,000 on 8 GPUs (one row-sharded chain)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --markers 500000 --steps 5 --warmup 3 --burn 6 --no-cpu --no-e2e > gpurun_out/c4_n8.json 2> gpurun_out/c4_n8.err
cat gpurun_out/c4_n8.json | cut -c1-1500; tail -3 gpurun_out/c4_n8.err
