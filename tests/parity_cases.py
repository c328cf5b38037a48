"""Shared helpers for the GPU parity tests and __graft_entry__.smoke(): seeded inputs, the oracle
side (CPU restatement under oracle/) and the comparison metric of SURVEY.md 8(d):
  Index/shape results  - exact equality
  fp32 tensors         - max|gpu - ref| / max|ref| <= 1e-4 per tensor."""
import numpy as np

TOL = 1e-4


def norm_err(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(got, ref, tol=TOL, what=""):
    e = norm_err(got, ref)
    assert e <= tol, f"{what}: normalised max error {e:.3e} > {tol}"
    return e


_oracle_matmul = None


def oracle_matmul(a, b):
    """c[y,x] ++= a[y,it]*b[it,x] through the oracle (sequential fp32 accumulation in `it` order,
    exprgrad/layers/base.nim:27-28 lowered as exprgrad/llvmgen.nim:277-297)."""
    global _oracle_matmul
    import oracle as o
    if _oracle_matmul is None:
        c = o.Fun(); x, y, it = o.Iter("x"), o.Iter("y"), o.Iter("it")
        c[y, x] += o.input("a")[y, it] * o.input("b")[it, x]
        _oracle_matmul = o.compile(c.target("c"))
    return _oracle_matmul.call("c", {"a": a, "b": b})


def gemm_f32(ctx, a, b, ta=0, tb=0, flags=0, c0=None, bias=None, alpha=1.0):
    """Raw C-ABI contraction on host arrays (blocking copies around one asynchronous launch)."""
    import ctypes
    import exprgrad_b200 as eg
    from exprgrad_b200._ffi import check, lib
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    da, db = eg.alloc_tensor(ctx, a.shape), eg.alloc_tensor(ctx, b.shape)
    dc = eg.alloc_tensor(ctx, (M, N))
    da.write(a); db.write(b)
    dc.write(c0 if c0 is not None else np.zeros((M, N), np.float32))
    dbias = None
    if bias is not None:
        dbias = eg.alloc_tensor(ctx, bias.shape); dbias.write(bias)
    check(lib.egb_gemm_f32(ctx.handle, ta, tb, M, N, K, da.buffer.device_ptr, a.shape[1], db.buffer.device_ptr,
                           b.shape[1], dc.buffer.device_ptr, N, flags,
                           dbias.buffer.device_ptr if dbias else None, ctypes.c_float(alpha)))
    out = dc.read()
    for t in (da, db, dc, dbias):
        if t is not None:
            t.buffer.dealloc()
    return out


def smoke_case():
    """One small invocation of the hot path on cuda:0, checked against the oracle: a dense-net train
    step (forward contractions on tcgen05, adjoints, SGD) through the model C-ABI."""
    import oracle as o
    from oracle import layers as OL
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as PL
    import graphs as G
    sizes = (64, 48, 32, 10)
    ctx = eg.new_gpu_context()
    om = o.compile(*G.dense_net(o, OL, sizes), seed=1)
    pm = eg.compile(*G.dense_net(F, PL, sizes), gpu=ctx, seed=1)
    x, y, params = G.dense_inputs(96, sizes)
    for tid, v in zip(sorted(om.params), params):
        om.params[tid][...] = v
        pm.params[tid] = v
    n0 = ctx.launch_count
    for _ in range(2):
        om.apply("train", {"x": x, "y": y})
        pm.apply("train", {"x": x, "y": y})
    worst = 0.0
    for tid in sorted(om.params):
        worst = max(worst, assert_close(pm.params[tid], om.params[tid], what=f"smoke: param tensor{tid - 1}"))
    e = assert_close(pm.call("loss", {"x": x, "y": y}), om.call("loss", {"x": x, "y": y}), what="smoke: loss")
    launches = ctx.launch_count - n0
    assert launches >= 20, launches
    print(f"smoke ok: dense {sizes} train step x2 vs oracle: params err {worst:.2e}, loss err {e:.2e}, "
          f"{launches} kernel launches")
    pm.free()
    ctx.destroy()
