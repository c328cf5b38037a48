"""Backends under test for the reference known-answer vectors:
  oracle        the CPU restatement (oracle/) - pins the oracle to the reference's own tests
  b200          the product (libegb200.so through exprgrad_b200) in its default mode
  b200_strict   the product in bit-exact mode (sequential accumulation, no tensor cores)
Both expose the same DSL surface so that one test body runs against all of them."""
import types

import numpy as np

NAMES = ["Fun", "Iter", "input", "param", "select", "sq", "to_scalar", "ShapeError", "RuntimeError_"]


def oracle_backend():
    import oracle as o
    from oracle import layers as L
    ns = types.SimpleNamespace(name="oracle", o=o, L=L, exact=True, exact_libm=True)
    for n in NAMES:
        setattr(ns, n, getattr(o, n))
    return ns


_ctx = None


def b200_backend(strict):
    global _ctx
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as L, model as M
    if _ctx is None:
        _ctx = eg.new_gpu_context()

    def compile_(*graphs, scalar="float32", seed=0, **kw):
        # float64 models are out of scope of the B200 backend; the reference vectors that use them hold
        # small integers, which fp32 represents exactly.
        return M.compile(*graphs, gpu=_ctx, seed=seed, strict=strict)

    o = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("_")})
    o.compile = compile_
    o.ShapeError = eg.ShapeError
    o.RuntimeError_ = eg.RuntimeError_
    o.layers = L
    ns = types.SimpleNamespace(name="b200_strict" if strict else "b200", o=o, L=L, exact=True, exact_libm=False)
    for n in NAMES:
        setattr(ns, n, getattr(o, n))
    return ns


def make_backend(name):
    if name == "oracle":
        return oracle_backend()
    return b200_backend(strict=(name == "b200_strict"))


def close_libm(got, ref, be):
    """The reference compares transcendental results with the host libm using exact ==; CUDA's
    expf/sinf/... differ from glibc in the last ulp, so the device backends use a 4-ulp tolerance."""
    if be.exact_libm:
        return np.array_equal(got, ref)
    return np.allclose(got, ref, rtol=5e-7, atol=1e-7)
