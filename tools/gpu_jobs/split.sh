#!/bin/bash
# SM split experiment: sweep workers vs Gram CTAs at the three BASELINE shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in 98 106; do
  timeout 300 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 6 --warmup 3 --burn 3 --no-cpu --no-e2e --workers $w > gpurun_out/r2_split_groups_w$w.json 2>/dev/null; echo "groups w=$w rc=$?"
done
for w in 98 106 124; do
  timeout 300 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 6 --warmup 3 --burn 3 --no-cpu --no-e2e --workers $w > gpurun_out/r2_split_hs_w$w.json 2>/dev/null; echo "hs w=$w rc=$?"
done
for w in 98 107 124; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --workers $w > gpurun_out/r2_split_v2_w$w.json 2>/dev/null; echo "v2 w=$w rc=$?"
done
