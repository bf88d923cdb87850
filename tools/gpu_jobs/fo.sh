#!/bin/bash
# one GPU: default build through the GPU suite + bench shapes; round-profile build and HEAD beside it on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-fo}
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
run head_v2 head
run rp128_v2 rp128
run hs "" --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --burn 5
run groups "" --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
run v2_w112 "" --workers 112
python tools/summ.py gpurun_out/${tag}_*.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/fo_rp*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['cycles_per_block']
    print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v) for k,v in c.items() if k in ('gather','serial_pass','publish','eval_cycles','resolve_cycles','prologue_cycles','gather_last_chunk','worker_reduce')})
PY
