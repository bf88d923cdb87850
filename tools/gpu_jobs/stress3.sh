#!/bin/bash
# repeat the three-thread-rank parity test in fresh processes (intermittent failures)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-st}; n=${2:-4}
for i in $(seq 1 $n); do
  timeout 120 python -m pytest tests/test_sharded.py -m gpu -q -x -s -k "thread_ranks_match and 3-1500" > gpurun_out/r2_${tag}_run$i.log 2>&1; rc=$?
  echo "run $i rc=$rc $(grep -h -o 'first attempt: .*' gpurun_out/r2_${tag}_run$i.log | head -1 | cut -c1-160) $(grep -h -o 'error 4: .*' gpurun_out/r2_${tag}_run$i.log | head -1 | cut -c1-160) $(tail -1 gpurun_out/r2_${tag}_run$i.log)"
done
