#!/bin/bash
TAG=${1:-conv}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
echo "== conv tests"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rf --tb=short -k "conv2 and not full_size" 2>&1 | tail -4
fi
echo "== conv bench"
timeout 300 python bench.py --workload conv2 --no-cpu --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['kernels_ms'], d['roofline']['per_kernel_frac'])
"
