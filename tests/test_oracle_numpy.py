"""CPU tier: the oracle (the parity anchor of every GPU test) against independent closed-form numpy (fp64)
implementations of the BASELINE workloads' layers. The reference's own known-answer tests
(tests/test_reference_vectors.py) pin the oracle on tiny cases; nothing in the reference pins C2/C3/C4-shaped
results, so these tests check the oracle's *semantics* there: layer definitions from
exprgrad/layers/base.nim and dnn.nim, adjoints derived by hand - not by the oracle's autodiff."""
import numpy as np
import pytest

import graphs as G

TOL = 2e-5   # fp32 sequential accumulation vs fp64 closed form, normalised by max|ref|


def err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def oracle():
    import oracle as o
    from oracle import layers as OL
    return o, OL


def test_matmul_matches_numpy(oracle):
    """base.nim:27-28 / matmul_gpu.nim:32: c[y,x] ++= a[y,it] * b[it,x]."""
    o, OL = oracle
    m = o.compile(*G.matmul(o, OL, ct="threads"))
    rng = np.random.default_rng(0)
    for (M, K, N) in [(5, 7, 3), (64, 129, 33)]:
        a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        b = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        assert err(m.call("c", {"a": a, "b": b}), a.astype(np.float64) @ b.astype(np.float64)) < TOL


def _dense_forward(x, params):
    """dnn.nim:19-27 (dense + bias, relu = select(0 <= h, h, 0)), dnn.nim:90-94 (softmax without max-subtraction)."""
    acts, pre = [x.astype(np.float64)], []
    n_layers = len(params) // 2
    for i in range(n_layers):
        h = acts[-1] @ params[2 * i].astype(np.float64) + params[2 * i + 1].astype(np.float64)
        pre.append(h)
        acts.append(np.where(0.0 <= h, h, 0.0) if i < n_layers - 1 else h)
    e = np.exp(pre[-1])
    p = e / e.sum(1, keepdims=True)
    return acts, pre, p


def test_dense_net_step_matches_numpy(oracle):
    """BASELINE config 3 at reduced size: predict, loss (base.nim:66-67: sum(-y * ln(p)) / shape[0]) and one
    gradientDescent step (base.nim:37-38: P += (0 - g) * rate) against hand-derived gradients:
    dh_last = (p * sum_x(y) - y) / N, dW = a^T dh, db = column sums of dh, da = dh W^T masked by 0 <= h."""
    o, OL = oracle
    sizes, rate, B = (20, 16, 12, 5), 0.01, 9
    m = o.compile(*G.dense_net(o, OL, sizes=sizes, rate=rate, ct="threads"), seed=0)
    x, y, params = G.dense_inputs(B, sizes)
    ids = sorted(m.params)
    for tid, v in zip(ids, params):
        m.params[tid][...] = v
    acts, pre, p = _dense_forward(x, params)
    assert err(m.call("predict", {"x": x}), p) < TOL
    loss = -(y.astype(np.float64) * np.log(p)).sum() / B
    assert err(m.call("loss", {"x": x, "y": y}), [loss]) < TOL
    dh = (p * y.astype(np.float64).sum(1, keepdims=True) - y) / B
    want = [None] * len(params)
    for i in reversed(range(len(params) // 2)):
        want[2 * i] = params[2 * i].astype(np.float64) - rate * (acts[i].T @ dh)
        want[2 * i + 1] = params[2 * i + 1].astype(np.float64) - rate * dh.sum(0)
        if i > 0:
            dh = (dh @ params[2 * i].astype(np.float64).T) * (0.0 <= pre[i - 1])
    m.apply("train", {"x": x, "y": y})
    for tid, w, before in zip(ids, want, params):
        got = m.params[tid]
        assert err(got, w) < TOL
        # the update itself (not just the parameter) against the closed form
        assert err(got.astype(np.float64) - before, w - before) < 2e-3


@pytest.mark.parametrize("shape,filters", [((2, 7, 6, 3), (4, 3, 3, 3)), ((1, 5, 9, 2), (3, 2, 3, 2))])
def test_conv2_forward_and_adjoints_match_numpy(oracle, shape, filters):
    """dnn.nim:45-49: out[n,y,x,f] ++= img[n,y+dy,x+dx,c] * w[f,dy,dx,c] (NHWC, valid); with loss = sum(out^2):
    d_out = 2 out, d_w[f,dy,dx,c] = sum d_out * img (shifted), d_img = scatter of d_out * w."""
    o, OL = oracle
    m = o.compile(*G.conv2_net(o, OL, ct="threads", filters=filters), seed=0)
    rng = np.random.default_rng(1)
    w = rng.uniform(-2, 2, filters).astype(np.float32)
    m.params[sorted(m.params)[0]][...] = w
    img = rng.uniform(0, 1, shape).astype(np.float32)
    N, H, W, C = shape
    F, KH, KW, _ = filters
    OH, OW = H - KH + 1, W - KW + 1
    i64, w64 = img.astype(np.float64), w.astype(np.float64)
    out = np.zeros((N, OH, OW, F))
    for dy in range(KH):
        for dx in range(KW):
            out += np.einsum("nyxc,fc->nyxf", i64[:, dy:dy + OH, dx:dx + OW, :], w64[:, dy, dx, :])
    assert err(m.call("conv", {"img": img}), out) < TOL
    assert err(m.call("loss", {"img": img}), [(out ** 2).sum()]) < TOL
    d_out = 2.0 * out
    d_w = np.zeros(filters)
    d_img = np.zeros(shape)
    for dy in range(KH):
        for dx in range(KW):
            d_w[:, dy, dx, :] = np.einsum("nyxf,nyxc->fc", d_out, i64[:, dy:dy + OH, dx:dx + OW, :])
            d_img[:, dy:dy + OH, dx:dx + OW, :] += np.einsum("nyxf,fc->nyxc", d_out, w64[:, dy, dx, :])
    assert err(m.call("dw", {"img": img}), d_w) < TOL
    assert err(m.call("dimg", {"img": img}), d_img) < TOL


def test_xor_net_forward_matches_numpy(oracle):
    """BASELINE config 1 (examples/xor/xor.nim:20-28): dense -> leakyRelu (select(0 <= h, 1, leak = 0.01) * h, dnn.nim:29-30) -> dense ->
    sigmoid (1 / (1 + exp(-h))); mse = sum((a - b)^2) / shape[0] (base.nim:57-58)."""
    o, OL = oracle
    graphs = G.xor_net(o, OL, ct="threads")
    m = o.compile(*graphs, seed=0)
    x = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float32)
    y = np.array([[0], [1], [1], [0]], np.float32)
    ids = sorted(m.params)
    vals = [m.params[t].astype(np.float64) for t in ids]
    by_shape = {v.shape: v for v in vals}
    w1, b1, w2, b2 = by_shape[(2, 4)], by_shape[(4,)], by_shape[(4, 1)], by_shape[(1,)]
    h1 = x @ w1 + b1
    a1 = np.where(0.0 <= h1, 1.0, 0.01) * h1
    pred = 1.0 / (1.0 + np.exp(-(a1 @ w2 + b2)))
    loss = ((pred - y) ** 2).sum() / 4
    assert err(m.call("loss", {"x": x, "y": y}), [loss]) < TOL
