#!/bin/bash
# N-GPU DP bench (peer exchange) with phase stamps ; usage: gpurun --gpus N -- bash tools/gpu_dpN.sh <tag> <N> [tests]
TAG=${1:-r02v}
N=${2:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
if [ "$3" == "tests" ]; then
  echo "== pytest test_gpu_dp"
  timeout 900 python -m pytest tests/test_gpu_dp.py -q -rf --tb=short -k "matches_global_batch" > gpurun_out/${TAG}_pytest_dp.log 2>&1
  grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest_dp.log | head -20
fi
for n in $N; do
echo "== bench --gpus $n peer exchange"
EGB_EXCHANGE_TRACE=1 run $n 2961$n --steps 200 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_n$n.err | tee gpurun_out/${TAG}_bench_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_classes'].get('exchange'))"
grep "exchange . trace rank 0" gpurun_out/${TAG}_n$n.err | head -4
done
