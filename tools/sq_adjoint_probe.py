"""Streaming maps on conv2-sized tensors (3.2 GB) vs 512 MiB ones: the square-adjoint form and relu under the cache
policies of eltwise_stream.cu (EGB_ELT_POLICY is read per launch). Usage: python tools/sq_adjoint_probe.py [once|quick]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL, gpu as G
ctx = eg.new_gpu_context()
once = len(sys.argv) > 1 and sys.argv[1] == 'once'      # one launch per kernel (ncu)
quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'    # 3.2 GB tensors, default policy only
x = F.input("x", [-1, 64]); s = F.input("s", [1])
r = F.Fun(); it = F.Iter("it"); r.raw[it] += s[0] * x.raw[it] + s[0] * x.raw[it]; r.copy_shape(x)
pm = eg.compile(r.target("sqadj", "gpu"), PL.relu(x).target("relu", "gpu"), gpu=ctx)
ds = eg.alloc_tensor(ctx, (1,)); ds.write(np.full(1, 0.5, np.float32))
for rows in ([12616704] if once or quick else [2097152, 12616704]):          # 512 MiB and 256*222*222 rows of 64 (3.2 GB)
    dx = eg.alloc_tensor(ctx, (rows, 64))
    dx.fill(1.0)
    n = rows * 64
    for name, args in (("sqadj", {"x": dx, "s": ds}), ("relu", {"x": dx})):
        for pol in ([None] if once or quick else [None, "0", "4", "5", "1"]):
            if pol is None: os.environ.pop("EGB_ELT_POLICY", None)
            else: os.environ["EGB_ELT_POLICY"] = pol
            pm.apply(name, args)
            reps = 1 if once else 5
            e0, e1 = G.GpuEvent(ctx), G.GpuEvent(ctx); e0.record()
            for _ in range(reps): pm.apply(name, args, sync=False)
            e1.record(); ms = e0.elapsed_ms(e1) / reps
            print(f"{name:6s} n={n:10d} policy={pol or 'default'}: {ms:7.3f} ms {8 * n / ms / 1e6:7.0f} GB/s", flush=True)
            if pol is None and once:
                print(pm.describe_plan(), flush=True)
    dx.buffer.dealloc()
