#!/bin/bash
# full single-GPU gate: every GPU test, smoke, default bench, reference arm
TAG=${1:-r02full}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q -rf --tb=short --durations=6 > gpurun_out/${TAG}_pytest_full.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest_full.log | head -40
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default)" ; timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-600
tail -3 gpurun_out/${TAG}_bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-300
