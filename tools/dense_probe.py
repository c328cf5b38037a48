import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL
import graphs as G
ctx = eg.new_gpu_context()
pm = eg.compile(*G.dense_net(F, PL), gpu=ctx, seed=0)
pm.set_option("graphs", 0)
x, y, params = G.dense_inputs(1024)
dx, dy = eg.alloc_tensor(ctx, x.shape), eg.alloc_tensor(ctx, y.shape)
dx.write(x); dy.write(y)
for _ in range(4):
    pm.apply("train", {"x": dx, "y": dy})
print(pm.describe_plan())
