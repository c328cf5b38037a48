#!/bin/bash
# Store schedule x trigger placement of the streaming kernels inside the conv2 adjoint targets (graph replay with
# programmatic edges), and the eltwise bench block under the late trigger
TAG=${1:-r04e}
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0"; do
  set -- $cfg
  echo "== variant $1, late trigger $2"
  EGB_ELT_SCALAR_VARIANT=$1 EGB_ELT_LATE_TRIGGER=$2 timeout 40 python bench.py --workload conv2 --no-cpu --steps 5 2>/dev/null > gpurun_out/${TAG}_bench_conv2_v$1_l$2.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_conv2_v$1_l$2.json').read().strip().splitlines()[-1]); print({k: (round(v['target_ms'], 3), round(v['all_kernels_ms'], 3)) for k, v in d['targets'].items()})"
done 2>&1 | tee gpurun_out/${TAG}_summary.txt
echo "== eltwise bench, late trigger"
EGB_ELT_LATE_TRIGGER=1 timeout 40 python bench.py --workload eltwise --no-cpu --steps 10 2>/dev/null > gpurun_out/${TAG}_bench_eltwise_l1.json
python -c "import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_eltwise_l1.json').read().strip().splitlines()[-1]); print({k: (round(v['target_ms'], 4), round(v['frac'], 3)) for k, v in d['per_target'].items()})" | tee -a gpurun_out/${TAG}_summary.txt
