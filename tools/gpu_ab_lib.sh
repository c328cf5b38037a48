#!/bin/bash
# A/B of two builds of the library on one box: exprgrad_b200/libegb200_old.bin vs libegb200_new.bin
b() { timeout 300 python bench.py --workload dense --no-extras --no-cpu --steps 300 --warmup 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3,2), 'us')"; }
for round in 1 2; do
  for v in old new; do cp exprgrad_b200/libegb200_$v.bin exprgrad_b200/libegb200.so; echo -n "$v: "; b; done
done
cp exprgrad_b200/libegb200_new.bin exprgrad_b200/libegb200.so
