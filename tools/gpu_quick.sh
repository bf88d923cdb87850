#!/bin/bash
# parity (each test in its own process) + one bench line
bash tools/gpu_each.sh "$1" > gpurun_out/each_summary.log 2>&1; tail -4 gpurun_out/each_summary.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
