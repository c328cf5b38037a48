#!/bin/bash
TAG=${1:-r02w}; N=${2:-8}
mkdir -p gpurun_out
run() { EGB_EXCHANGE_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_classes'].get('exchange'))"; grep "exchange 1 trace rank 0" gpurun_out/${TAG}_$1.err | head -2; }
for v in "$@"; do :; done
echo "-- backoff 0"; EGB_DP_BACKOFF_NS=0 run 29701
echo "-- backoff 200"; EGB_DP_BACKOFF_NS=200 run 29702
echo "-- backoff 400 (default)"; run 29703
echo "-- backoff 1000"; EGB_DP_BACKOFF_NS=1000 run 29704
