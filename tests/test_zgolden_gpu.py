"""GPU parity against COMMITTED vectors (tests/golden/*.npz; generated from the oracle by tests/golden/make_golden.py):
the same comparisons the live-oracle tests make, but nothing under oracle/ is imported here - the device path stands
against frozen numbers. Bars: SURVEY 8(d), normalised max error 1e-4 per tensor (parameter UPDATES 2e-3: they are
differences of nearly equal fp32 numbers). (The file name sorts behind the live-oracle GPU tests on purpose: this file was added
after the last GPU run of the round, and `pytest -x` should reach everything that has run on a B200 before it.)"""
import os

import numpy as np
import pytest

import graphs as G
from parity_cases import assert_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ctx():
    import exprgrad_b200 as eg
    c = eg.new_gpu_context()
    yield c
    c.destroy()


@pytest.mark.parametrize("strict", [False, True])
def test_matmul_and_conv2_golden(ctx, strict):
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    g = _load("matmul")
    pm = M.compile(*G.matmul(F, PL), gpu=ctx, strict=strict)
    got = pm.call("c", {"a": g["a"], "b": g["b"]})
    if strict:      # the reference's own accumulation order, un-contracted: bit for bit
        assert np.array_equal(got, g["c"])
    else:
        assert_close(got, g["c"], what="matmul")
    pm.free()
    g = _load("conv2")
    pm = M.compile(*G.conv2_net(F, PL), gpu=ctx, seed=1, strict=strict)
    pm.params[pm.params.ids()[0]] = g["filters"]
    for t in ("conv", "loss", "dw", "dimg"):
        got = pm.call(t, {"img": g["img"]})
        assert_close(got, g[t], tol=1e-6 if strict else 1e-4, what=f"conv2 {t} (strict={strict})")
    pm.free()


def test_dense_step_golden(ctx):
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    from golden.make_golden import DENSE_SIZES
    g = _load("dense_step")
    pm = M.compile(*G.dense_net(F, PL, DENSE_SIZES), gpu=ctx, seed=1)
    tids = pm.params.ids()
    for i, tid in enumerate(tids):
        pm.params[tid] = g[f"param{i}_before"]
    assert_close(pm.call("predict", {"x": g["x"]}), g["predict"], what="predict")
    assert_close(pm.call("loss", {"x": g["x"], "y": g["y"]}), g["loss"], what="loss")
    for _ in range(2):
        pm.apply("train", {"x": g["x"], "y": g["y"]})
    for i, tid in enumerate(tids):
        assert_close(pm.params[tid], g[f"param{i}_after"], what=f"param{i} after two steps")
        assert_close(pm.params[tid] - g[f"param{i}_before"], g[f"param{i}_after"] - g[f"param{i}_before"], tol=2e-3,
                     what=f"update of param{i}")
    pm.free()


def test_xor_golden(ctx):
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    g = _load("xor")
    pm = M.compile(*G.xor_net(F, PL, rate=0.1), gpu=ctx, seed=1)
    for i, tid in enumerate(pm.params.ids()):
        pm.params[tid] = g[f"param{i}_before"]
    X = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float32); Y = np.array([[0], [1], [1], [0]], np.float32)
    for _ in range(100):
        pm.apply("train", {"x": X, "y": Y}, sync=False)
    for i, tid in enumerate(pm.params.ids()):
        assert_close(pm.params[tid], g[f"param{i}_after"], tol=1e-4, what=f"xor param{i} after 100 steps")
    pm.free()
