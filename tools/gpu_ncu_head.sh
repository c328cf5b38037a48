#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_rows_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_head \
    python bench.py --workload dense --no-extras --no-cpu --steps 6 --warmup 3 > gpurun_out/${TAG}_ncu_head.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_head.log | cut -c1-200
ncu -i gpurun_out/${TAG}_head.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc|Issue Slots Busy|No Eligible|One or More|Registers Per|Achieved Occupancy|Theoretical Occ|Stall|L1/TEX Hit|L2 Hit|Warp Cycles Per Issued|Block Limit|Waves" | head -40
