#!/bin/bash
TAG=${1:-r02w}; N=${2:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 300 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_$1.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['ms_per_step']*1e3,2), 'us', round(d['value']/1e6,1), 'M/s')"; }
echo "-- default (two exchanges, early part 64 CTAs)"; run 29701
echo "-- early part 96 CTAs"; EGB_DP_EARLY_CTAS=96 run 29702
echo "-- early part 128 CTAs"; EGB_DP_EARLY_CTAS=128 run 29703
echo "-- one exchange"; EGB_DP_ONE_EXCHANGE=1 run 29704
echo "-- default again"; run 29705
