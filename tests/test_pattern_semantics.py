"""CPU-only: every kernel the planner's matchers claim (contraction, conv2, map forms) computes what the IR kernel
computes. For each such kernel of the reference-shaped graphs, of seeded random graphs and of random layer stacks, the
pattern's meaning (tests/pattern_semantics.py, from the operands / constants the specialised device kernel would be
launched with) is compared with the generic lowering of the same kernel (tests/ip_interp.py) on random tensor contents."""
import collections

import numpy as np
import pytest

import fuzz_graphs as FG
import graphs as G
from ip_interp import run_program
from pattern_semantics import apply_conv2, apply_eltwise, apply_gemm

SEEN = collections.Counter()


def _check_target(prog, target, shapes, epoch=1, label=""):
    dump = prog.lower_dump(target, shapes, strict=True, epoch=epoch)
    rng = np.random.default_rng(len(dump) * 7919 + sum(sum(s) for s in shapes.values()))
    for p in dump:
        rec = next((k for k in ("gemm", "conv2", "eltwise") if k in p), None)
        if rec is None:
            continue
        ids = {p["write"]["tensor"]} | {rd["tensor"] for rd in p["reads"]}
        sizes = {}
        for tid in ids:
            info = prog.tensor_info(tid)
            shape = info["shape"] if info["kind"] in ("param", "cache") else prog.infer_shapes(target, shapes, tensor_id=tid)
            sizes[tid] = int(np.prod(shape)) if shape else 1
        positive = rec == "eltwise" and p["eltwise"]["form"] == "adam-step"      # sqrt(v / c2): second moments are >= 0
        base = {tid: (rng.uniform(0.05, 1, n) if positive else rng.uniform(-1, 1, n)).astype(np.float32) for tid, n in sizes.items()}
        a = {tid: v.copy() for tid, v in base.items()}
        b = {tid: v.copy() for tid, v in base.items()}
        run_program(p, a)
        if rec == "gemm":
            apply_gemm(p["gemm"], b)
        elif rec == "conv2":
            apply_conv2(p["conv2"], b)
        else:
            apply_eltwise(p["eltwise"], p["write"]["tensor"], b, epoch)
        w = p["write"]["tensor"]
        name = rec if rec != "eltwise" else "eltwise " + p["eltwise"]["form"]
        SEEN[name] += 1
        if rec == "eltwise":
            assert np.array_equal(a[w], b[w]), f"{label} {target} kernel {p['kernel']} ({name}): max |diff| {np.abs(a[w] - b[w]).max()}"
        else:
            scale = max(np.abs(a[w]).max(), 1e-30)
            err = np.abs(a[w].astype(np.float64) - b[w]).max() / scale
            assert err < 2e-6, f"{label} {target} kernel {p['kernel']} ({name}): {err:.2e}"
        for tid in ids - {w}:
            assert np.array_equal(a[tid], base[tid]) and np.array_equal(b[tid], base[tid])


def test_reference_shaped_graphs():
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    prog = Program.from_graphs(G.dense_net(F, PL, (12, 8, 6, 4))).compile()
    _check_target(prog, "train", {"x": [9, 12], "y": [9, 4]}, label="dense")
    prog = Program.from_graphs(G.conv2_net(F, PL, filters=(4, 3, 2, 3))).compile()
    for t in ("conv", "dw", "dimg"):
        _check_target(prog, t, {"img": [2, 7, 6, 3]}, label="conv2")
    prog = Program.from_graphs(G.fashion_net(F, PL)).compile()
    _check_target(prog, "train", {"x": [3, 12, 12, 1], "y": [3, 10]}, epoch=3, label="fashion")
    prog = Program.from_graphs(G.xor_net(F, PL)).compile()
    _check_target(prog, "train", {"x": [4, 2], "y": [4, 1]}, label="xor")
    for name in ("gemm", "conv2", "eltwise bias-row-add", "eltwise relu", "eltwise relu-adjoint", "eltwise sgd-axpy",
                 "eltwise leakyRelu", "eltwise leakyRelu-adjoint", "eltwise adam-m", "eltwise adam-v", "eltwise adam-step",
                 "eltwise sigmoid", "eltwise sigmoid-adjoint", "eltwise square-adjoint"):
        assert SEEN[name] > 0, name


@pytest.mark.parametrize("seed", list(range(40)))
def test_random_graphs(seed):
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    graphs, what = FG.random_net(F, PL, seed)
    prog = Program.from_graphs(graphs).compile()
    shapes = {"a": [5, FG.COLS], "b": [5, FG.COLS], "v": [FG.COLS]}
    for tid in range(1, prog.tensor_count() + 1):
        pass
    used = {prog.tensor_info(t)["name"] for t in range(1, prog.tensor_count() + 1) if prog.tensor_info(t)["kind"] == "input"}
    shapes = {k: v for k, v in shapes.items() if k in used}
    for target in ("out", "loss", "da", "train"):
        try:
            _check_target(prog, target, shapes, label=f"seed {seed} ({what})")
        except Exception as e:
            if "is not a target" in str(e):
                continue
            raise


@pytest.mark.parametrize("seed", list(range(30)))
def test_random_layer_stacks(seed):
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    graphs, what, sh = FG.random_cnn(F, PL, seed)
    prog = Program.from_graphs(graphs).compile()
    _check_target(prog, "train", {"x": [2] + sh["x"][1:], "y": [2, sh["outs"]]}, epoch=2, label=f"cnn {seed} ({what})")


def test_every_map_form_is_covered():
    """the forms the graphs above do not contain: tanh and its adjoint, tensor subtraction, division by a constant"""
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    x = F.input("x", [-1, 6]); y = F.input("y", [-1, 6])
    h = PL.tanh(x)
    loss = F.Fun(); it = F.Iter("it"); loss[0] += F.sq(h.raw[it])
    d = PL.sub(x, y); d.copy_shape(x)
    prog = Program.from_graphs([loss.backwards().grad(x).target("dx", "gpu"), d.target("diff", "gpu"),
                                PL.divide(x, 3.0).target("div", "gpu"), PL.leaky_relu(x, 0.2).target("leaky", "gpu")]).compile()
    _check_target(prog, "dx", {"x": [5, 6]}, label="tanh")
    _check_target(prog, "diff", {"x": [5, 6], "y": [5, 6]}, label="sub")
    _check_target(prog, "div", {"x": [5, 6]}, label="div")
    _check_target(prog, "leaky", {"x": [5, 6]}, label="leaky")
    for name in ("eltwise tanh", "eltwise tanh-adjoint", "eltwise sub", "eltwise div-const", "eltwise leakyRelu"):
        assert SEEN[name] > 0, name


@pytest.mark.parametrize("seed", list(range(30)))
def test_random_index_graphs(seed):
    """indices that only LOOK like a map or a contraction (strided, divided, wrapped, shifted by explicit bounds): whatever
    the matchers still claim must compute what the IR kernel computes"""
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    graphs, what = FG.random_index_net(F, PL, seed)
    try:
        prog = Program.from_graphs(graphs).compile()
    except (eg.GradientError, eg.ShapeError):
        pytest.skip("compile-time error of the reference (see tests/test_passes_fuzz.py)")
    used = {prog.tensor_info(t)["name"] for t in range(1, prog.tensor_count() + 1) if prog.tensor_info(t)["kind"] == "input"}
    shapes = {k: v for k, v in {"a": [5, 12], "v": [12]}.items() if k in used}
    for target in ("out", "loss", "da"):
        _check_target(prog, target, shapes, label=f"index graph {seed} ({what})")
