"""CPU-only: the committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle)
against the oracle as it builds and runs on THIS host - with the thread-pool compile target where the fixtures were
generated single-threaded - and the library's native shape inference against the frozen integer tables. Guards the
checker itself: another gcc, glibc or CPU must not move the numbers the device path is compared with."""
import json
import os

import numpy as np
import pytest

import graphs as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# fp32 results that involve libm (exp / ln) may differ in the last ulp between glibc versions; everything else is exact
LIBM_TOL = 2e-6


def _load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def _close(got, want, tol, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    if tol == 0:
        assert np.array_equal(got, want), f"{what}: max |diff| {np.abs(got - want).max()}"
    else:
        err = np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30)
        assert err <= tol, f"{what}: normalised max error {err:.2e}"


def test_matmul_golden_is_exact():
    import oracle as o
    from oracle import layers as OL
    g = _load("matmul")
    m = o.compile(*G.matmul(o, OL, ct="threads"))
    _close(m.call("c", {"a": g["a"], "b": g["b"]}), g["c"], 0, "matmul")
    # and the definition itself: sequential fp32 accumulation in `it` order (llvmgen.nim:277-297)
    c = np.zeros((37, 29), np.float32)
    for it in range(53):
        c += g["a"][:, it:it + 1] * g["b"][it:it + 1, :]
    assert np.array_equal(c, g["c"])


def test_conv2_golden_is_exact():
    import oracle as o
    from oracle import layers as OL
    g = _load("conv2")
    m = o.compile(*G.conv2_net(o, OL, ct="threads"), seed=1)
    m.params[sorted(m.params)[0]][...] = g["filters"]
    for t in ("conv", "loss", "dw", "dimg"):
        _close(m.call(t, {"img": g["img"]}), g[t], 0, f"conv2 {t}")


def test_dense_step_golden():
    import oracle as o
    from oracle import layers as OL
    from golden.make_golden import DENSE_SIZES
    g = _load("dense_step")
    m = o.compile(*G.dense_net(o, OL, DENSE_SIZES, ct="threads"), seed=1)
    tids = sorted(m.params)
    for i, tid in enumerate(tids):
        m.params[tid][...] = g[f"param{i}_before"]
    _close(m.call("predict", {"x": g["x"]}), g["predict"], LIBM_TOL, "predict")
    _close(m.call("loss", {"x": g["x"], "y": g["y"]}), g["loss"], LIBM_TOL, "loss")
    for _ in range(2):
        m.apply("train", {"x": g["x"], "y": g["y"]})
    for i, tid in enumerate(tids):
        _close(m.params[tid], g[f"param{i}_after"], LIBM_TOL, f"param{i} after two steps")


def test_xor_and_adam_golden():
    import oracle as o
    from oracle import layers as OL
    g = _load("xor")
    m = o.compile(*G.xor_net(o, OL, rate=0.1, ct="threads"), seed=1)
    for i, tid in enumerate(sorted(m.params)):
        m.params[tid][...] = g[f"param{i}_before"]
    X = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float32); Y = np.array([[0], [1], [1], [0]], np.float32)
    losses = []
    for step in range(100):
        m.apply("train", {"x": X, "y": Y})
        if step % 10 == 9:
            losses.append(float(m.call("loss", {"x": X, "y": Y})[0]))
    _close(np.array(losses, np.float32), g["losses"], 1e-5, "xor loss trajectory")
    for i, tid in enumerate(sorted(m.params)):
        _close(m.params[tid], g[f"param{i}_after"], 1e-5, f"xor param{i}")
    g = _load("fashion_adam")
    m = o.compile(*G.fashion_net(o, OL, ct="threads"), seed=1)
    for i, tid in enumerate(sorted(m.params)):
        m.params[tid][...] = g[f"param{i}_before"]
    for _ in range(2):
        m.epoch += 1
        m.apply("train", {"x": g["x"], "y": g["y"]})
    for i, tid in enumerate(sorted(m.params)):
        _close(m.params[tid], g[f"param{i}_after"], 1e-5, f"adam param{i}")
    for i, tid in enumerate(sorted(m.caches)):
        _close(m.caches[tid], g[f"cache{i}_after"], 1e-5, f"adam cache{i}")


@pytest.mark.parametrize("backend", ["oracle", "library"])
def test_shape_tables_are_bit_exact(backend):
    """integer work (passes.nim:1386-1436): the frozen tables against the oracle and against csrc/passes.cpp"""
    table = json.load(open(os.path.join(GOLD, "shapes.json")))
    assert len(table) == 5
    for case in table:
        want = {int(t): s for t, s in case["shapes"].items()}
        if backend == "oracle":
            import oracle as o
            from oracle import layers as OL
            from oracle.passes import compile_program, infer_shapes
            prog = o.ir.to_program(G.ALL[case["graph"]](o, OL))
            compile_program(prog)
            got = infer_shapes(prog, case["target"], {prog.inputs[k]: v for k, v in case["inputs"].items()})
            assert {t: list(s) for t, s in got.items()} == want, case["graph"]
        else:
            from exprgrad_b200 import frontend as F, layers as PL
            from exprgrad_b200.model import Program
            prog = Program.from_graphs(G.ALL[case["graph"]](F, PL)).compile()
            for tid, shape in want.items():
                assert prog.infer_shapes(case["target"], case["inputs"], tensor_id=tid) == shape, (case["graph"], tid)
