#!/bin/bash
# quick single-GPU check of a contraction-kernel change: gemm + dense parity tests, device timeline, dense bench
TAG=${1:-r02f}
mkdir -p gpurun_out
echo "== pytest gemm + dense"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -q -x -rf --tb=short -k "not full_size and not conv2" > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest.log | head -30
echo "== dense step timeline"
timeout 300 python tools/gemm_trace.py 2>&1 | tail -32 | tee gpurun_out/${TAG}_gemm_trace.txt
echo "== dense bench"
timeout 600 python bench.py --workload dense --no-extras --no-cpu --steps 200 --warmup 20 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_classes'], d['roofline'].get('eager_ms_per_step'))"
tail -3 gpurun_out/${TAG}_bench.err
