#!/bin/bash
# Variant 1 of the scalar-operand stream kernel with its residency bounded at three CTAs per SM
TAG=${1:-r04f}
mkdir -p gpurun_out
export EGB_ELT_SCALAR_VARIANT=1 EGB_ELT_SCALAR_SMEM=73728
echo "== variant 1, 72 KB dynamic shared memory per CTA"
timeout 25 python bench.py --workload conv2 --no-cpu --steps 5 2>/dev/null > gpurun_out/${TAG}_bench_conv2_v1_smem.json
python -c "import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_conv2_v1_smem.json').read().strip().splitlines()[-1]); print({k: (round(v['target_ms'], 3), round(v['all_kernels_ms'], 3)) for k, v in d['targets'].items()})" | tee gpurun_out/${TAG}_summary.txt
timeout 20 python -m pytest tests/test_gpu_eltwise.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
