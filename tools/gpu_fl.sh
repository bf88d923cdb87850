#!/bin/bash
# one GPU: the default build through the GPU suite and the bench shapes, launch experiments (env), the look-ahead-64 variant beside it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-fl}
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
run() {   # name lib-suffix extra-args...
  local name=$1 suf=$2; shift 2
  if [ -n "$suf" ]; then export BRR_LIB="$GRAFT_REPO_ROOT/bayesrrcpp_b200/libbayesrr_b200_$suf.so"; else unset BRR_LIB; fi
  timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; el "$name rc=$?"
}
unset BRR_LIB
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; el "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
run v2 ""
BRR_KEV_AFTER_TABLES=1 run v2_kev ""
BRR_KEV_AFTER_TABLES=1 BRR_PLAIN_LAUNCH=1 run v2_plain ""
run hs "" --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --burn 5
run groups "" --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
export BRR_LIB="$GRAFT_REPO_ROOT/bayesrrcpp_b200/libbayesrr_b200_la64.so"
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_la64_pytest.log 2>&1; el "la64 parity rc=$?"; tail -1 gpurun_out/${tag}_la64_pytest.log
run la64_v2 la64
run la64_v2_w112 la64 --workers 112
run la64_groups la64 --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
run la64_hs la64 --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --burn 5
run head_v2 head
python tools/summ.py gpurun_out/${tag}_*.json 2>/dev/null
