import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import frontend as F, layers as PL
import graphs as G
import oracle as o
from oracle import layers as OL
ctx = eg.new_gpu_context()
x, y, params = G.dense_inputs(300)
om = o.compile(*G.dense_net(o, OL, ct="threads"), seed=0)
ids = sorted(om.params)
for t, v in zip(ids, params): om.params[t][...] = v
om.apply("train", {"x": x, "y": y})
ref = [om.params[t].copy() for t in ids]
for fuse in (1, 0):
    for sk in (1, 0):
        pm = eg.compile(*G.dense_net(F, PL), gpu=ctx, seed=0)
        pm.set_option("fuse", fuse); pm.set_option("splitk", sk)
        for t, v in zip(ids, params): pm.params[t] = v
        pm.apply("train", {"x": x, "y": y})
        errs = []
        for t, r, v in zip(ids, ref, params):
            g = pm.params[t]
            errs.append(float(np.abs((g - v) - (r - v)).max() / max(np.abs(r - v).max(), 1e-30)))
        print("fuse", fuse, "splitk", sk, ["%.1e" % e for e in errs], flush=True)
        if fuse == 0 and sk == 1: print(pm.describe_plan())
        pm.free()
