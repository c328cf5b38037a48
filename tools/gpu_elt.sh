#!/bin/bash
b() { timeout 300 python bench.py --workload eltwise --no-cpu --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v['kernels_ms'],4), round(v['frac'],3)) for k,v in d['per_target'].items()})
"; }
echo "== default"; b
for p in 8 12 9 13; do echo "== policy $p"; EGB_ELT_POLICY=$p b; done
