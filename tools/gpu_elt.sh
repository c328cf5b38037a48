#!/bin/bash
echo "== eltwise + adam tests"
timeout 900 python -m pytest tests/test_gpu_eltwise.py tests/test_gpu_model.py -m gpu -q -rf --tb=short -k "optimizer or adam or eltwise or checkpoint or fashion" 2>&1 | tail -5
echo "== eltwise bench"
timeout 300 python bench.py --workload eltwise --no-cpu --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v['kernels_ms'],4), round(v['frac'],3), v['launches']) for k,v in d['per_target'].items()})
"
