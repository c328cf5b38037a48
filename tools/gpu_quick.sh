#!/bin/bash
# Shorter gpurun call: GPU tests, bench (no reference arm), dense-step launch list.
# Usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
echo "== bench" ; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
bash tools/gpu_profile_dense.sh ${TAG} > /dev/null 2>&1
tail -5 gpurun_out/${TAG}_bench.err
