#!/bin/bash
# One short call for the round-2 tail: the streaming square-adjoint form (eltwise + conv2 tests), the random-graph
# differential test, and the conv2 bench block whose adjoint targets contain that kernel.
# Usage: bash tools/gpu_r4a.sh <tag>
TAG=${1:-r04a}
mkdir -p gpurun_out
echo "== eltwise"; timeout 120 python -m pytest tests/test_gpu_eltwise.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_eltwise.log
echo "== conv2";   timeout 120 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "conv2 or small_layer or avgpool" 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_conv2.log
echo "== fuzz";    timeout 150 python -m pytest tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/${TAG}_pytest_fuzz.log
echo "== bench conv2"; timeout 120 python bench.py --workload conv2 2>gpurun_out/${TAG}_bench_conv2.err | tee gpurun_out/${TAG}_bench_conv2.json | cut -c1-1500
tail -3 gpurun_out/${TAG}_bench_conv2.err
