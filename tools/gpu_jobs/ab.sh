#!/bin/bash
# one GPU, one box: A/B of library variants (BRR_LIB) on the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
run() {   # tag lib-suffix extra-args...
  local tag=$1 suf=$2; shift 2
  if [ -n "$suf" ]; then export BRR_LIB="$GRAFT_REPO_ROOT/bayesrrcpp_b200/libbayesrr_b200_$suf.so"; else unset BRR_LIB; fi
  timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; el "$tag rc=$?"
}
run head_w112 head
run head_w98 head --workers 98
run la64_w112 la64 --workers 112
run la64_w98 la64
run la128_w112 la128 --workers 112
run la128pp la128pp
run la128_groups la128 --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
run head_groups head --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
run head_hs head --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --burn 5
python tools/summ.py gpurun_out/ab_*.json 2>/dev/null
