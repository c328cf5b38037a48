"""A/B timing of the 4096^3 contraction inside the full step (split + split + gemm), several repeats."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL, gpu as G
import graphs as GR
ctx = eg.new_gpu_context()
n = 4096
rng = np.random.default_rng(0)
a = rng.uniform(0, 1, (n, n)).astype(np.float32); b = rng.uniform(0, 1, (n, n)).astype(np.float32)
da, db = eg.alloc_tensor(ctx, (n, n)), eg.alloc_tensor(ctx, (n, n)); da.write(a); db.write(b)
m = eg.compile(*GR.matmul(F, PL), gpu=ctx)
for rep in range(3):
    for _ in range(5): m.apply("c", {"a": da, "b": db}, sync=False)
    ctx.synchronize()
    e0, e1 = G.GpuEvent(ctx), G.GpuEvent(ctx)
    e0.record()
    for _ in range(100): m.apply("c", {"a": da, "b": db}, sync=False)
    e1.record(); ms = e0.elapsed_ms(e1) / 100
    G.set_timing(ctx, True)
    for _ in range(20): m.apply("c", {"a": da, "b": db}, sync=False)
    k, kn = G.kernel_time(ctx, "gemm"); s_, sn = G.kernel_time(ctx, "split")
    G.set_timing(ctx, False)
    print(f"step {ms*1e3:.1f} us  gemm {k/kn*1e3:.1f} us  split {s_/sn*1e3:.1f} us x{sn//20}  -> {2*n**3/ms/1e9:.1f} TFLOP/s", flush=True)
