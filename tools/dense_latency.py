"""Isolated-step latency vs back-to-back throughput of the dense train step (CUDA events)."""
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
pm = eg.compile(*G.dense_net(F, PL), gpu=ctx)
if os.environ.get("EGB_DP_FORCE"):
    from exprgrad_b200 import dist as D
    D.set_data_parallel(pm, D.Comm(ctx, rank=0, world=1))
for kv in sys.argv[1:]:
    k, v = kv.split("="); pm.set_option(k, int(v))
for _ in range(20): pm.apply("train", a, sync=False)
ctx.synchronize()
lat = []
for rep in range(20):
    e0, e1 = GG.GpuEvent(ctx), GG.GpuEvent(ctx)
    ctx.synchronize(); e0.record(); pm.apply("train", a, sync=False); e1.record(); ctx.synchronize()
    lat.append(e0.elapsed_ms(e1) * 1e3)
e0, e1 = GG.GpuEvent(ctx), GG.GpuEvent(ctx); e0.record()
for _ in range(300): pm.apply("train", a, sync=False)
e1.record(); ctx.synchronize()
print(pm.describe_plan()) if os.environ.get("EGB_DP_FORCE") else None
print("isolated step us: min %.1f median %.1f | back-to-back us/step %.1f" % (min(lat), sorted(lat)[10], e0.elapsed_ms(e1) / 300 * 1e3))
