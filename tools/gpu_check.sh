#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list and one full capture of the top kernel.
# Usage (from the repo root, on the GPU box): bash tools/gpu_check.sh <tag> [kernel-regex] [conv]
TAG=${1:-r01}
KREGEX=${2:-gemm_bf16x3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
echo "== bench" ; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-conv --no-dense --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-conv --no-dense > gpurun_out/${TAG}_ncu_full.log 2>&1
# optional third argument "conv": one full capture per tensor-core conv2 kernel as well
if [ "$3" = "conv" ]; then
  echo "== ncu full (conv2)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2_.*tc_kernel -c 6 -f -o gpurun_out/${TAG}_conv \
      python bench.py --workload conv2 --no-cpu > gpurun_out/${TAG}_ncu_conv.log 2>&1
fi
ls -la gpurun_out | tail -20
