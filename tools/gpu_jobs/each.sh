#!/bin/bash
# Run every GPU test in its own process (a faulting kernel poisons the CUDA context of its process only).
# usage: tools/gpu_jobs/each.sh [pytest -k expression]
mkdir -p gpurun_out
LOG=gpurun_out/gpu_each.log
: > $LOG
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv >> $LOG 2>&1
ids=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; fail=0
for t in $ids; do
  echo "=== $t" >> $LOG
  if timeout 600 python -m pytest "$t" -q -x --timeout 500 -p no:cacheprovider >> $LOG 2>&1; then pass=$((pass+1)); else fail=$((fail+1)); echo "FAILED: $t" | tee -a $LOG; fi
done
echo "passed=$pass failed=$fail" | tee -a $LOG
[ $fail -eq 0 ]
