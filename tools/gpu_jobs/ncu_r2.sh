#!/bin/bash
# round 2 ncu evidence (one GPU): launch list of a short bench run and one --set full capture of the sweep and Gram kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu --no-e2e"
$SHORT > gpurun_out/r2_ncu_plain1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $SHORT > gpurun_out/r2_ncu_list.log 2>&1
echo "list rc=$?"; tail -2 gpurun_out/r2_ncu_list.log
$SHORT > gpurun_out/r2_ncu_plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'sweep_kernel|gram_tc_kernel' -s 10 -c 2 -o gpurun_out/r2_ncu_full -f $SHORT > gpurun_out/r2_ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out | grep r2_ncu
