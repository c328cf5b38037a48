#!/bin/bash
# Multi-GPU gate: data-parallel parity tests + the DP dense bench at N ranks (peer-memory exchange and the NCCL arm).
# Usage (gpurun --gpus N): bash tools/gpu_dp.sh <tag> <N>
TAG=${1:-r02dp}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== pytest test_gpu_dp"
timeout 900 python -m pytest tests/test_gpu_dp.py -q -rf --tb=short > gpurun_out/${TAG}_pytest_dp.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/${TAG}_pytest_dp.log | head -40
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
echo "== bench --gpus 1 (dense primary on a multi-GPU box)"
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --no-cpu --no-extras 2>gpurun_out/${TAG}_n1.err | tee gpurun_out/${TAG}_bench_n1.json | cut -c1-400
echo "== bench --gpus $N peer exchange"
run $N 29611 --steps 200 --warmup 20 --no-cpu 2>gpurun_out/${TAG}_n${N}.err | tee gpurun_out/${TAG}_bench_n${N}.json | cut -c1-400
tail -3 gpurun_out/${TAG}_n${N}.err
echo "== bench --gpus $N NCCL arm"
EGB_DP_NCCL=1 run $N 29612 --steps 200 --warmup 20 --no-cpu --no-extras 2>>gpurun_out/${TAG}_n${N}.err | tee gpurun_out/${TAG}_bench_n${N}_nccl.json | cut -c1-400
echo "== reference arm --gpus $N"
run $N 29613 --impl reference --steps 5 --warmup 1 2>>gpurun_out/${TAG}_n${N}.err | tee gpurun_out/${TAG}_bench_n${N}_ref.json | cut -c1-300
ls -la gpurun_out | tail -6
