#!/bin/bash
# round 2, GPU call 1: whole GPU suite, bench lines of the three samplers, stand-alone Gram capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_j1_smi.txt
lscpu | head -20 > gpurun_out/r2_j1_cpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2_j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_j1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_j1_bench.json 2> gpurun_out/r2_j1_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e > gpurun_out/r2_j1_bench_groups.json 2> gpurun_out/r2_j1_bench_groups.err; echo "groups rc=$?"
timeout 300 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e > gpurun_out/r2_j1_bench_hs.json 2> gpurun_out/r2_j1_bench_hs.err; echo "hs rc=$?"
python tools/gram_alone.py 50000 12800 128 > gpurun_out/r2_j1_gram_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc -s 2 -c 1 -o gpurun_out/r2_gram_alone python tools/gram_alone.py 50000 12800 128 > gpurun_out/r2_j1_gram_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r2_j1_pytest.log
cat gpurun_out/r2_j1_bench.json | head -c 1500
