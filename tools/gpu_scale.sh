#!/bin/bash
# scaling points on one multi-GPU box: bash tools/gpu_scale.sh <tag> "<N list>"   (first N with the extra workloads)
TAG=${1:-scale}
mkdir -p gpurun_out
first=1
for n in $2; do
  extra="--no-extras"; [ $first == 1 ] && extra=""
  first=0
  echo "== bench --gpus $n $extra"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 200 --warmup 20 --no-cpu $extra 2>gpurun_out/${TAG}_n$n.err | tee gpurun_out/${TAG}_bench_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['metric'], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],5), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
for k in ('dense_strong','matmul_replicas'):
    if k in d: print('  ', k, 'ms', round(d[k]['ms_per_step'],5), 'value', round(d[k]['value']))
"
  tail -2 gpurun_out/${TAG}_n$n.err | cut -c1-200
done
