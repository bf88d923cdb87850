#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in 0 124 140; do
  timeout 300 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 6 --warmup 3 --burn 3 --no-cpu --no-e2e --workers $w > gpurun_out/r2_hsprof_w$w.json 2>/dev/null; echo "hs w=$w rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_hsprof_v2.json 2>/dev/null; echo "v2 rc=$?"
timeout 300 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 6 --warmup 3 --burn 3 --no-cpu --no-e2e > gpurun_out/r2_hsprof_groups.json 2>/dev/null; echo "groups rc=$?"
