#!/bin/bash
# quick single-GPU check: the parity tests that exercise the walk + the three bench lines; tag as $1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -q -x > gpurun_out/r2_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_${tag}_pytest.log
tail -3 gpurun_out/r2_${tag}_pytest.log
bash tools/gpu_jobs/quick.sh $tag
