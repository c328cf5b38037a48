#!/bin/bash
# Round-2 single-GPU gate.   Usage: bash tools/gpu_r2c.sh <tag> [quick]
TAG=${1:-r02c}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q -rf --tb=short --durations=8 > gpurun_out/${TAG}_pytest_full.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest_full.log | head -60
tail -25 gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== dense step timeline"
timeout 300 python tools/gemm_trace.py 2>&1 | tail -32 | tee gpurun_out/${TAG}_gemm_trace.txt
echo "-- keep_intermediates=1"
timeout 300 python tools/gemm_trace.py keep_intermediates=1 2>&1 | grep -E "step us|csk|epi unit"
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-700
tail -5 gpurun_out/${TAG}_bench.err
echo "-- matmul without the tail-wave split"
EGB_GEMM_NO_TAIL_SPLIT=1 timeout 300 python bench.py --workload matmul --no-extras --no-cpu --steps 20 --warmup 5 2>&1 | cut -c1-260
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu: relu-adjoint (out of place) vs sgd-axpy (in place)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:elt_stream_kernel -s 16 -c 2 -f -o gpurun_out/${TAG}_elt_reluadj \
    python bench.py --workload eltwise --steps 5 --no-cpu > gpurun_out/${TAG}_ncu_elt1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:elt_stream_kernel -s 58 -c 2 -f -o gpurun_out/${TAG}_elt_sgd \
    python bench.py --workload eltwise --steps 5 --no-cpu > gpurun_out/${TAG}_ncu_elt2.log 2>&1
ls -la gpurun_out | tail -8
