"""Large elementwise / broadcast / reduction kernels through the model API (for ncu captures and GB/s)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL, gpu as G
ctx = eg.new_gpu_context()
R, C = 65536, 2048                       # 134 M elements = 512 MiB per tensor
x = F.input("x"); b = F.input("b")
relu = PL.relu(x)
bias = F.Fun(); y_, x_ = F.Iter("y"), F.Iter("x"); bias[y_, x_] += x[y_, x_] + b[x_]
mse = F.Fun(); it = F.Iter("it"); mse[0] += F.sq(x.raw[it] - relu.raw[it]) / F.to_scalar(x.shape[0])
cs = F.Fun(); y_, x_ = F.Iter("y"), F.Iter("x"); cs[x_] += x[y_, x_]
sgd = F.Fun(); it = F.Iter("it"); sgd.raw[it] += x.raw[it] + (-relu.raw[it]) * 0.01; sgd.copy_shape(x)
pm = eg.compile(relu.target("relu", "gpu"), bias.target("bias", "gpu"), mse.target("mse", "gpu"), cs.target("colsum", "gpu"),
                sgd.target("axpy", "gpu"), gpu=ctx)
dx = eg.alloc_tensor(ctx, (R, C)); db = eg.alloc_tensor(ctx, (C,))
dx.write(np.random.default_rng(0).uniform(-1, 1, (R, C)).astype(np.float32)); db.write(np.ones(C, np.float32))
n = R * C
for name, args, bytes_ in [("relu", {"x": dx}, 8 * n), ("bias", {"x": dx, "b": db}, 8 * n), ("mse", {"x": dx}, 8 * n + 8 * n),
                           ("colsum", {"x": dx}, 4 * n), ("axpy", {"x": dx}, 8 * n + 12 * n)]:
    pm.apply(name, args)
    e0, e1 = G.GpuEvent(ctx), G.GpuEvent(ctx); e0.record()
    for _ in range(5): pm.apply(name, args, sync=False)
    e1.record(); ms = e0.elapsed_ms(e1) / 5
    print(f"{name:8s} {ms:8.3f} ms  {bytes_ / ms / 1e6:8.0f} GB/s (algorithmic bytes of the whole target)", flush=True)
    print(pm.describe_plan())
