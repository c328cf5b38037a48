#!/bin/bash
# single GPU: conv2 TMA-store check + source-level ncu of the dense step's contraction kernel
TAG=${1:-r02d}
mkdir -p gpurun_out
echo "== conv tests"
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -rf --tb=short -k "conv2 and not full_size_matches" 2>&1 | tail -5
echo "== conv bench (TMA store)"
timeout 300 python bench.py --workload conv2 --no-cpu --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['kernels_ms'], d['roofline']['per_kernel_frac'], d['targets'])
"
echo "-- without TMA store"
EGB_CONV_NO_TMA_STORE=1 timeout 300 python bench.py --workload conv2 --no-cpu --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['kernels_ms'])
"
echo "== ncu source-level: contraction kernels of one dense train step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3_kernel -s 24 -c 8 -f -o gpurun_out/${TAG}_dense_gemm \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_dense.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_dense.log | cut -c1-200
echo "== ncu conv fwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_conv_fwd \
    python bench.py --workload conv2 --no-cpu --steps 3 > gpurun_out/${TAG}_ncu_conv.log 2>&1
ls -la gpurun_out | tail -5
