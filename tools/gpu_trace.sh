#!/bin/bash
# device timeline of the dense step only (few seconds of GPU time)
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 300 python tools/gemm_trace.py "${@:2}" 2>&1 | tail -32 | grep -v "^  #" | tee gpurun_out/${TAG}_gemm_trace.txt
