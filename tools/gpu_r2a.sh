#!/bin/bash
# Round-2 gate, one gpurun call: GPU tests, smoke, bench (both arms), eltwise ncu capture.
# Usage: bash tools/gpu_r2a.sh <tag>
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-3000
tail -5 gpurun_out/${TAG}_bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-600
echo "== ncu eltwise"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:elt_stream_kernel -c 8 -f -o gpurun_out/${TAG}_eltwise \
    python bench.py --workload eltwise --steps 5 --no-cpu > gpurun_out/${TAG}_ncu_eltwise.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_eltwise.log
ls -la gpurun_out | tail -12
