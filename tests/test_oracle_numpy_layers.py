"""CPU tier: the remaining layers of exprgrad/layers/base.nim and dnn.nim - the ones tests/test_oracle_numpy.py does
not reach - through the oracle against closed-form numpy (fp64) with hand-derived adjoints: tanh, avgpool2, maxpool2 and
its customGrad, upsample2, binaryCrossEntropy, the tensor arithmetic layers, transpose, dropout's scaling and adam.
Together with the token-identity of the library's compiled programs (tests/test_passes_parity.py, test_passes_fuzz.py)
this pins both copies of the layer library (oracle/layers.py, exprgrad_b200/layers.py) on something neither was
derived from."""
import numpy as np
import pytest

TOL = 2e-6


def err(got, ref):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def oracle():
    import oracle as o
    from oracle import layers as OL
    return o, OL


def _with_loss(o, layer_out, wrt):
    """loss = sum(w * out) with a fixed random weighting w (an input), so that d loss / d out = w"""
    w = o.input("w")
    loss = o.Fun(); it = o.Iter("it")
    loss[0] += layer_out.raw[it] * w.raw[it]
    return [layer_out.target("out", "cpu"), loss.backwards().grad(wrt).target("grad", "cpu")]


def test_tanh_and_its_adjoint(oracle):
    o, OL = oracle
    x = o.input("x", [-1, 5])
    m = o.compile(*_with_loss(o, OL.tanh(x), x), openmp=False)
    rng = np.random.default_rng(0)
    xv = rng.uniform(-2, 2, (4, 5)).astype(np.float32); wv = rng.uniform(-1, 1, (4, 5)).astype(np.float32)
    t = np.tanh(xv.astype(np.float64))
    assert err(m.call("out", {"x": xv}), t) < TOL                                     # dnn.nim:35-40
    assert err(m.call("grad", {"x": xv, "w": wv}), wv * (1 - t * t)) < 5e-6


def test_pooling_and_upsampling(oracle):
    o, OL = oracle
    rng = np.random.default_rng(1)
    img = rng.uniform(-1, 1, (2, 6, 4, 3)).astype(np.float32)
    img[0, 0, 0, 0] = img[0, 1, 1, 0] = 0.99                                           # a tie inside one 2x2 window
    blocks = img.astype(np.float64).reshape(2, 3, 2, 2, 2, 3)                          # [n, y, dy, x, dx, c]
    # avgpool2 (dnn.nim:73-79) and its adjoint: every input of a window receives w / 4
    x = o.input("img", [-1, 6, 4, 3])
    m = o.compile(*_with_loss(o, OL.avgpool2(x), x), openmp=False)
    wv = rng.uniform(-1, 1, (2, 3, 2, 3)).astype(np.float32)
    assert err(m.call("out", {"img": img}), blocks.mean(axis=(2, 4))) < TOL
    up = np.repeat(np.repeat(wv.astype(np.float64), 2, axis=1), 2, axis=2)
    assert err(m.call("grad", {"img": img, "w": wv}), up / 4) < TOL
    # maxpool2 (dnn.nim:58-71): the customGrad sends the output gradient to EVERY input equal to the window's maximum
    x = o.input("img", [-1, 6, 4, 3])
    m = o.compile(*_with_loss(o, OL.maxpool2(x), x), openmp=False)
    mx = blocks.max(axis=(2, 4))
    assert err(m.call("out", {"img": img}), mx) < TOL
    mx_up = np.repeat(np.repeat(mx, 2, axis=1), 2, axis=2)
    want = np.where(img.astype(np.float64) == mx_up, up, 0.0)
    got = m.call("grad", {"img": img, "w": wv})
    assert err(got, want) < TOL
    assert got[0, 0, 0, 0] == got[0, 1, 1, 0] == wv[0, 0, 0, 0]                        # both tied inputs
    # upsample2 (dnn.nim:81-88, `withShape`) and its adjoint: the sum over each 2x2 block of w
    x = o.input("img", [-1, 6, 4, 3])
    m = o.compile(*_with_loss(o, OL.upsample2(x), x), openmp=False)
    w2 = rng.uniform(-1, 1, (2, 12, 8, 3)).astype(np.float32)
    assert err(m.call("out", {"img": img}), np.repeat(np.repeat(img.astype(np.float64), 2, axis=1), 2, axis=2)) < TOL
    assert err(m.call("grad", {"img": img, "w": w2}), w2.astype(np.float64).reshape(2, 6, 2, 4, 2, 3).sum(axis=(2, 4))) < TOL


def test_losses(oracle):
    o, OL = oracle
    rng = np.random.default_rng(2)
    p = rng.uniform(0.05, 0.95, (6, 3)).astype(np.float32); y = rng.integers(0, 2, (6, 3)).astype(np.float32)
    p64, y64 = p.astype(np.float64), y.astype(np.float64)
    for name, layer, value, grad in [
            ("binaryCrossEntropy", OL.binary_cross_entropy,                                  # base.nim:60-64
             -(y64 * np.log(p64) + (1 - y64) * np.log(1 - p64)).sum() / 6, (-(y64 / p64) + (1 - y64) / (1 - p64)) / 6),
            ("crossEntropy", OL.cross_entropy, -(y64 * np.log(p64)).sum() / 6, -(y64 / p64) / 6),   # base.nim:66-67
            ("mse", OL.mse, ((p64 - y64) ** 2).sum() / 6, 2 * (p64 - y64) / 6)]:             # base.nim:57-58
        pi, yi = o.input("p", [-1, 3]), o.input("y", [-1, 3])
        loss = layer(pi, yi)
        m = o.compile(loss.target("loss", "cpu"), loss.backwards().grad(pi).target("grad", "cpu"), openmp=False)
        assert err(m.call("loss", {"p": p, "y": y}), [value]) < 5e-6, name
        assert err(m.call("grad", {"p": p, "y": y}), grad) < 5e-6, name


def test_tensor_arithmetic_and_transpose(oracle):
    o, OL = oracle
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (3, 4)).astype(np.float32); b = rng.uniform(-1, 1, (3, 4)).astype(np.float32)
    b[0, 0] = a[0, 0]
    cases = [("add", lambda x, y: OL.add(x, y), a + b), ("sub", lambda x, y: OL.sub(x, y), a - b),
             ("min", lambda x, y: OL.minimum(x, y), np.minimum(a, b)), ("max", lambda x, y: OL.maximum(x, y), np.maximum(a, b)),
             ("scale", lambda x, y: OL.scale(x, 2.5), a * np.float32(2.5)), ("divide", lambda x, y: OL.divide(x, 3.0), a / np.float32(3.0))]
    for name, build, want in cases:                                                          # base.nim:19-25
        x, y = o.input("a", [-1, 4]), o.input("b", [-1, 4])
        r = build(x, y); r.copy_shape(x)
        m = o.compile(r.target("r", "cpu"), openmp=False)
        args = {"a": a, "b": b} if name in ("add", "sub", "min", "max") else {"a": a}
        assert np.array_equal(m.call("r", args), want), name
    x = o.input("a", [-1, 4])
    m = o.compile(OL.transpose(x).target("t", "cpu"), openmp=False)                        # base.nim:32-33
    assert np.array_equal(m.call("t", {"a": a}), a.T)


def test_dropout_scaling(oracle):
    """dnn.nim:96-100: select(prob <= rand, x / (1 - prob), 0) with rand ~ U(0, 1) drawn per call"""
    o, OL = oracle
    x = o.input("x", [-1, 50])
    m = o.compile(OL.dropout(x, 0.25).target("y", "cpu"), seed=0, openmp=False)
    xv = np.random.default_rng(4).uniform(0.5, 1.5, (40, 50)).astype(np.float32)
    y = m.call("y", {"x": xv})
    kept = y != 0
    assert np.allclose(y[kept], (xv / np.float32(0.75))[kept], rtol=1e-6)
    assert 0.65 < kept.mean() < 0.85


def test_adam_two_steps(oracle):
    """base.nim:40-53: m += m (b1 - 1) + (1 - b1) g; v += v (b2 - 1) + (1 - b2) g^2;
    p += -eta (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps), t = epoch()"""
    o, OL = oracle
    p = o.param([7], name="p"); g = o.input("g", [7])
    eff = o.Fun("Effect", effect=p)
    OL.adam(0.01)(eff, g)
    mdl = o.compile(eff.target("step", "cpu"), seed=0, openmp=False)
    tid = sorted(mdl.params)[0]
    p0 = np.random.default_rng(5).uniform(-1, 1, 7).astype(np.float32)
    mdl.params[tid][...] = p0
    pw, mw, vw = p0.astype(np.float64), np.zeros(7), np.zeros(7)
    for t in (1, 2):
        gv = np.random.default_rng(10 + t).uniform(-1, 1, 7).astype(np.float32)
        mdl.epoch = t
        mdl.apply("step", {"g": gv})
        g64 = gv.astype(np.float64)
        mw = 0.9 * mw + 0.1 * g64
        vw = 0.999 * vw + 0.001 * g64 * g64
        pw = pw - 0.01 * (mw / (1 - 0.9 ** t)) / (np.sqrt(vw / (1 - 0.999 ** t)) + 1e-8)
    assert err(mdl.params[tid], pw) < 5e-6
    caches = [mdl.caches[c] for c in sorted(mdl.caches)]
    assert min(err(caches[0], mw), err(caches[1], mw)) < 5e-6 and min(err(caches[0], vw), err(caches[1], vw)) < 5e-5
