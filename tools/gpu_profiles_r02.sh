#!/bin/bash
# Round-2 profile captures (one GPU): launch lists of the timed regions + full ncu captures of the new kernels.
TAG=${1:-r02z}
mkdir -p gpurun_out
echo "== launch list: default bench (timed region, no host legs)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-conv --no-eltwise > gpurun_out/${TAG}_launches_bench.log 2>&1
tail -1 gpurun_out/${TAG}_launches_bench.log | cut -c1-200
echo "== launch list: dense step (eager)"
bash tools/gpu_profile_dense.sh ${TAG} > /dev/null 2>&1
echo "== ncu full: the contraction launches of one dense step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_lat_kernel -s 18 -c 5 -f -o gpurun_out/${TAG}_gemm_lat \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_gemm_lat.log 2>&1
echo "== ncu full: head kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_rows_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_head \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_head.log 2>&1
echo "== ncu full: matmul 2-CTA kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3_2cta_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_matmul \
    python bench.py --workload matmul --no-extras --no-cpu --no-e2e --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_matmul.log 2>&1
echo "== ncu full: conv2 d_filters"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_dw_tc_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_conv_dw \
    python bench.py --workload conv2 --no-cpu --steps 3 > gpurun_out/${TAG}_ncu_conv_dw.log 2>&1
ls -la gpurun_out | grep ${TAG} | tail -12
