#!/bin/bash
# bench + ncu evidence for one round; writes into gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest "tests/test_gpu_parity.py::test_groups_chain_matches_oracle" -q -p no:cacheprovider 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
SHORT="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu --no-e2e"
$SHORT > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
$SHORT > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'sweep_kernel|gram_tc_kernel' -s 10 -c 2 -o gpurun_out/prof_r1e -f $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
