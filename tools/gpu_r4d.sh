#!/bin/bash
# Store schedule of the scalar-operand stream kernel (variant 0 / 1, two groups per thread) + the tests that cover it
TAG=${1:-r04d}
mkdir -p gpurun_out
{
echo "== variant 0"; timeout 40 python tools/sq_adjoint_probe.py quick
echo "== variant 1"; EGB_ELT_SCALAR_VARIANT=1 timeout 40 python tools/sq_adjoint_probe.py quick
echo "== two groups per thread"; EGB_ELT_UNROLL=2 timeout 40 python tools/sq_adjoint_probe.py quick
} 2>&1 | tee gpurun_out/${TAG}_stream_variants.txt
for v in 0 1; do
  echo "== tests, variant $v"
  EGB_ELT_SCALAR_VARIANT=$v timeout 60 python -m pytest tests/test_gpu_eltwise.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_v$v.log
done
echo "== conv2 (ragged shapes)"; timeout 60 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "conv2 and not full_size" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_conv2.log
echo "== smoke"; timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in 1 0; do
  echo "== bench conv2, variant $v"
  EGB_ELT_SCALAR_VARIANT=$v timeout 60 python bench.py --workload conv2 2>/dev/null > gpurun_out/${TAG}_bench_conv2_v$v.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_conv2_v$v.json').read().strip().splitlines()[-1]); print({k: round(v['target_ms'], 3) for k, v in d['targets'].items()})"
done
echo "== dense"; timeout 40 python bench.py --workload dense --no-extras --no-cpu --steps 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
