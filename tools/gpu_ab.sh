#!/bin/bash
# A/B of an environment switch on the dense step: bash tools/gpu_ab.sh <tag> <ENV=1>
TAG=${1:-ab}
mkdir -p gpurun_out
b() { timeout 300 python bench.py --workload dense --no-extras --no-cpu --steps 300 --warmup 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3,2), 'us')"; }
echo "-- default"; b; b
echo "-- $2"; env $2 bash -c "$(declare -f b); b; b"
echo "== trace (default)"
timeout 300 python tools/gemm_trace.py 2>&1 | tail -9 | cut -c1-330
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -q -x -k "not full_size and not conv2" 2>&1 | tail -2
