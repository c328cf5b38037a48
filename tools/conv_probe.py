"""Runs each conv2 target once at BASELINE config 4 size (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL
import graphs as G
ctx = eg.new_gpu_context()
pm = eg.compile(*G.conv2_net(F, PL, filters=(64, 3, 3, 3)), gpu=ctx, seed=0)
img = eg.alloc_tensor(ctx, (256, 224, 224, 3))
img.write(np.random.default_rng(0).uniform(0, 1, (256, 224, 224, 3)).astype(np.float32))
for t in ("conv", "dw", "dimg"):
    pm.apply(t, {"img": img})
print("done")
