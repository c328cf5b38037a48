#!/bin/bash
TAG=${1:-r03r}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_lat_kernel -s 18 -c 5 -f -o gpurun_out/${TAG}_gemm_lat \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_gemm_lat.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_rows_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_head \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adam_fused_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_adam \
    python bench.py --workload eltwise --no-cpu --steps 5 > gpurun_out/${TAG}_ncu_adam.log 2>&1
timeout 300 python tools/gemm_trace.py 2>&1 | tail -30 > gpurun_out/${TAG}_gemm_trace.txt
ls -la gpurun_out | grep ${TAG}
