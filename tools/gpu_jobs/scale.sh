#!/bin/bash
# scaling evidence on one multi-GPU box: torchrun bench at each N given; usage: tools/gpu_jobs/scale.sh 8 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_scale.log 2>&1
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  cat gpurun_out/scale_n$N.json | cut -c1-400; tail -3 gpurun_out/scale_n$N.err
done
