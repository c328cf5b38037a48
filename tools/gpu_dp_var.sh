#!/bin/bash
TAG=${1:-r02w}; N=${2:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_classes'].get('exchange'))"; }
echo "-- one exchange"; EGB_DP_ONE_EXCHANGE=1 run 29701
echo "-- two, early part 64 CTAs"; EGB_DP_EARLY_CTAS=64 run 29702
echo "-- two, early part 16 CTAs"; EGB_DP_EARLY_CTAS=16 run 29703
echo "-- two, early part 32 CTAs (default)"; run 29704
