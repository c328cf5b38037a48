"""Backends under test for the reference known-answer vectors:
  oracle        the CPU restatement (oracle/) - pins the oracle to the reference's own tests
  b200          the product (libegb200.so through exprgrad_b200) in its default mode
  b200_strict   the product in bit-exact mode (sequential accumulation, no tensor cores)
  b200_host     the product's host side only (parser, passes, shape inference, lowering) with tests/ip_interp.py in
                place of the device kernel - CPU tier
All expose the same DSL surface so that one test body runs against all of them."""
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


def host_backend():
    """The product's HOST side without a device (CPU tier): exprgrad_b200's graph builder -> the C ABI's parser, passes
    and shape inference -> the device programs lower.cpp builds (egb_program_lower_dump), executed by the sequential
    interpreter of tests/ip_interp.py in place of the CUDA kernel that would interpret them."""
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as L
    from exprgrad_b200.model import Program
    from ip_interp import run_target

    class HostModel:
        def __init__(self, graphs, seed):
            self.program = Program.from_graphs(graphs)
            self.program.compile()
            rng = np.random.default_rng(seed)
            self.state = {}
            for tid in range(1, self.program.tensor_count() + 1):
                info = self.program.tensor_info(tid)
                if info["kind"] == "param":      # the reference's default init range (parser.nim:714)
                    self.state[tid] = rng.uniform(-0.1, 0.1, info["shape"]).astype(np.float32)
                elif info["kind"] == "cache":
                    self.state[tid] = np.zeros(info["shape"], np.float32)
            self.epoch = 0
            self._dumps = {}

        def call(self, target, args=None):
            args = {k: np.ascontiguousarray(v, np.float32) for k, v in (args or {}).items()}
            out = run_target(self.program, target, args, self.state, strict=True, epoch=self.epoch, cache=self._dumps)
            return None if out is None else np.array(out)

        def apply(self, target, args=None):
            self.call(target, args)

        def fit(self, target, args, batch_size=32):
            if not args:
                raise eg.RuntimeError_("Model.fit requires at least one input tensor.")
            first = next(iter(args.values()))
            self.epoch += 1
            for b in range(first.shape[0] // batch_size):
                self.call(target, {k: v[b * batch_size:(b + 1) * batch_size] for k, v in args.items()})

    def compile_(*graphs, scalar="float32", seed=0, **kw):
        gs = []
        for g in graphs:
            gs.extend(g) if isinstance(g, (list, tuple)) else gs.append(g)
        return HostModel(gs, seed)

    o = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("_")})
    o.compile = compile_
    o.ShapeError = eg.ShapeError
    o.RuntimeError_ = eg.RuntimeError_
    o.layers = L
    ns = types.SimpleNamespace(name="b200_host", o=o, L=L, exact=True, exact_libm=False)
    for n in NAMES:
        setattr(ns, n, getattr(o, n))
    return ns


def make_backend(name):
    if name == "oracle":
        return oracle_backend()
    if name == "b200_host":
        return host_backend()
    return b200_backend(strict=(name == "b200_strict"))


def close_libm(got, ref, be):
    """The reference compares transcendental results with the host libm using exact ==; CUDA's
    expf/sinf/... differ from glibc in the last ulp, so the device backends use a 4-ulp tolerance."""
    if be.exact_libm:
        return np.array_equal(got, ref)
    return np.allclose(got, ref, rtol=5e-7, atol=1e-7)
