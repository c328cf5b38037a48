#!/bin/bash
# Round-2 single-GPU gate: GPU tests (all, no -x), smoke, bench (both arms), eltwise cache-policy probe, ncu of the
# in-place optimizer kernels.   Usage: bash tools/gpu_r2b.sh <tag>
TAG=${1:-r02b}
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1800 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -60 | tee gpurun_out/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-1500
tail -5 gpurun_out/${TAG}_bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-400
echo "== dense step timeline"
timeout 300 python tools/gemm_trace.py 2>&1 | tail -32 | tee gpurun_out/${TAG}_gemm_trace.txt
echo "-- equal stream priorities"
EGB_STREAM_PRIORITY=0 timeout 300 python tools/gemm_trace.py 2>&1 | grep "step us"
echo "== eltwise policy probe"
for mode in 7 0 1 3 5; do
  echo "-- EGB_ELT_POLICY=$mode"
  EGB_ELT_POLICY=$mode timeout 300 python bench.py --workload eltwise --steps 10 --no-cpu 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print({k:(round(v['kernels_ms'],4), round(v['frac'],3)) for k,v in d['per_target'].items()})
"
done 2>&1 | tee gpurun_out/${TAG}_elt_policy.log
echo "== ncu in-place optimizer kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:elt_stream_kernel<(7|16|17|18)' -s 8 -c 8 -f -o gpurun_out/${TAG}_eltwise \
    python bench.py --workload eltwise --steps 5 --no-cpu > gpurun_out/${TAG}_ncu_eltwise.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_eltwise.log | cut -c1-300
ls -la gpurun_out | tail -12
