#!/bin/bash
TAG=${1:-r04c}
mkdir -p gpurun_out
timeout 80 ncu --set full --clock-control none --import-source on -k regex:elt_stream -c 2 -o gpurun_out/${TAG}_sqadj python tools/sq_adjoint_probe.py once > gpurun_out/${TAG}_ncu.log 2>&1
grep -v "^==PROF" gpurun_out/${TAG}_ncu.log | tail -30
