"""A/B timing of the dense train step under model options (same box, same process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL, gpu as GG
import graphs as G
ctx = eg.new_gpu_context()
x, y, p = G.dense_inputs(1024)
dx, dy = eg.alloc_tensor(ctx, x.shape), eg.alloc_tensor(ctx, y.shape); dx.write(x); dy.write(y)
a = {"x": dx, "y": dy}
opts = [dict(), dict(splitk=0), dict(rowchain=0), dict(concurrent=0), dict(fuse=0)]
for o in opts:
    pm = eg.compile(*G.dense_net(F, PL), gpu=ctx)
    for k, v in o.items(): pm.set_option(k, v)
    for _ in range(20): pm.apply("train", a, sync=False)
    ctx.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = GG.GpuEvent(ctx), GG.GpuEvent(ctx); e0.record()
        for _ in range(300): pm.apply("train", a, sync=False)
        e1.record(); best = min(best, e0.elapsed_ms(e1) / 300 * 1e3)
    print(o, f"{best:.1f} us", flush=True)
    if not o: print(pm.describe_plan())
    pm.free()
