#!/bin/bash
# 2-GPU: exchange phase trace + DP bench (peer exchange) ; usage: gpurun --gpus 2 -- bash tools/gpu_dp2.sh <tag> [tests]
TAG=${1:-r02q}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
if [ "$2" == "tests" ]; then
  echo "== pytest test_gpu_dp"
  timeout 900 python -m pytest tests/test_gpu_dp.py -q -rf --tb=short > gpurun_out/${TAG}_pytest_dp.log 2>&1
  grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest_dp.log | head -20
fi
echo "== bench --gpus 2 peer exchange (with phase stamps)"
EGB_EXCHANGE_TRACE=1 run 2 29611 --steps 200 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_classes'])"
grep "exchange trace" gpurun_out/${TAG}_n2.err | head -8
tail -2 gpurun_out/${TAG}_n2.err | cut -c1-300
