#!/bin/bash
# Cache-policy sweep of the streaming maps on 512 MiB and 3.2 GB tensors + one ncu capture of the square-adjoint kernel
TAG=${1:-r04b}
mkdir -p gpurun_out
timeout 60 python tools/sq_adjoint_probe.py 2>&1 | tee gpurun_out/${TAG}_stream_policy.txt
timeout 80 ncu --set full --clock-control none --import-source on -k regex:elt_stream -s 2 -c 1 -o gpurun_out/${TAG}_sqadj python tools/sq_adjoint_probe.py once > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
