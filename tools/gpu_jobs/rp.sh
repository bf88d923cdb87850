#!/bin/bash
# one GPU, one box: round-profile builds (BRR_ROUND_PROFILE=1) of library variants on the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
run() {   # name lib-suffix extra-args...
  local name=$1 suf=$2; shift 2
  export BRR_LIB="$GRAFT_REPO_ROOT/bayesrrcpp_b200/libbayesrr_b200_$suf.so"
  timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/rp_$name.json 2> gpurun_out/rp_$name.err; el "$name rc=$?"
}
run head_w112 rphead
run la64_w112 rp64 --workers 112
run la128_w112 rp128 --workers 112
run la128_w98 rp128
run head_groups rphead --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
run la128_groups rp128 --sampler groups --rows 100000 --markers 200000 --steps 10 --burn 5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/rp_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['cycles_per_block']
    print(f.split('/')[-1], round(d['ms_per_step'],3), {k:round(v) for k,v in c.items() if k in ('gather','serial_pass','publish','eval_cycles','resolve_cycles','prologue_cycles','gather_last_chunk','worker_reduce')}, round(d['markers_per_speculative_window'],2))
PY
