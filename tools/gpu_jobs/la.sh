#!/bin/bash
# one GPU: the default build (look-ahead 96 at 128-marker blocks) through the whole GPU suite and the three bench shapes, then the
# variant builds (BRR_LIB) through the parity file and the default bench line.  Everything lands in gpurun_out/ as it is produced.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[$(( $(date +%s) - t0 )) s] $*"; }
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/la96_pytest.log 2>&1; el "la96 pytest rc=$?"; tail -3 gpurun_out/la96_pytest.log
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 40 > gpurun_out/la96_bench.json 2> gpurun_out/la96_bench.err; el "la96 bench rc=$?"
timeout 120 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e > gpurun_out/la96_hs.json 2> gpurun_out/la96_hs.err; el "la96 hs rc=$?"
timeout 150 python bench.py --sampler groups --rows 100000 --markers 200000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e > gpurun_out/la96_groups.json 2> gpurun_out/la96_groups.err; el "la96 groups rc=$?"
for v in la128 la64; do
  export BRR_LIB="$GRAFT_REPO_ROOT/bayesrrcpp_b200/libbayesrr_b200_$v.so"
  [ -f "$BRR_LIB" ] || continue
  timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${v}_pytest.log 2>&1; el "$v parity rc=$?"; tail -2 gpurun_out/${v}_pytest.log
  timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/${v}_bench.json 2> gpurun_out/${v}_bench.err; el "$v bench rc=$?"
  timeout 120 python bench.py --sampler horseshoe --rows 100000 --markers 100000 --steps 10 --warmup 3 --burn 5 --no-cpu --no-e2e > gpurun_out/${v}_hs.json 2> gpurun_out/${v}_hs.err; el "$v hs rc=$?"
done
unset BRR_LIB
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --workers 112 > gpurun_out/la96_bench_w112.json 2> gpurun_out/la96_bench_w112.err; el "la96 w112 rc=$?"
python tools/summ.py gpurun_out/la*_*.json 2>/dev/null | tail -20
