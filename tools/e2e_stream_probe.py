"""Times model.call('c', host a, host b, out=host c) at 4096^3 for several row-block sizes (EGB_STREAM_BLOCK_KIB)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, time
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, gpu as G, layers as PL
    import graphs as GR
    ctx = eg.new_gpu_context()
    n = 4096
    model = eg.compile(*GR.matmul(F, PL), gpu=ctx)
    ha, hb, hc = G.pinned_empty((n, n)), G.pinned_empty((n, n)), G.pinned_empty((n, n))
    rng = np.random.default_rng(0)
    ha[...] = rng.uniform(0, 1, (n, n)); hb[...] = rng.uniform(0, 1, (n, n))
    for _ in range(3): model.call("c", {"a": ha, "b": hb}, out=hc)
    t0 = time.perf_counter()
    for _ in range(10): model.call("c", {"a": ha, "b": hb}, out=hc)
    print("block KiB", os.environ.get("EGB_STREAM_BLOCK_KIB"), "ms/call", (time.perf_counter() - t0) / 10 * 1e3, flush=True)
else:
    for kib in sys.argv[1:] or ["2048", "4096", "8192", "16384", "65536"]:
        env = dict(os.environ, EGB_STREAM_BLOCK_KIB=kib)
        subprocess.run([sys.executable, __file__, "one"], env=env)
