"""GPU parity of the specialised streaming map kernels (csrc/eltwise_stream.cu) that run the fixed elementwise /
optimizer forms of exprgrad/layers/base.nim and dnn.nim: against the oracle, and against the generic loop-nest
kernel they replace (option eltwise=0), on sizes with and without a 4-element tail."""
import numpy as np
import pytest

from parity_cases import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import exprgrad_b200 as eg
    c = eg.new_gpu_context()
    yield c
    c.destroy()


def _act_net(d, L, act):
    x = d.input("x", [-1, -1])
    h = getattr(L, act)(x)
    loss = d.Fun(); it = d.Iter("it")
    loss[0] += d.sq(h.raw[it])
    return [h.target("y", "gpu"), loss.backwards().grad(x).target("dx", "gpu")]


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "sigmoid", "tanh"])
@pytest.mark.parametrize("shape", [(33, 129), (64, 256), (1, 3)])
def test_activations_and_adjoints(ctx, act, shape):
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    x = np.random.default_rng(sum(shape)).uniform(-3, 3, shape).astype(np.float32)
    x[0, 0] = 0.0                                   # the relu mask is `0 <= x` (true at zero, dnn.nim:26-27)
    om = o.compile(*_act_net(o, OL, act), seed=0)
    outs = {}
    for elt in (1, 0):
        pm = M.compile(*_act_net(F, PL, act), gpu=ctx, seed=0)
        pm.set_option("eltwise", elt)
        outs[elt] = (pm.call("y", {"x": x}), pm.call("dx", {"x": x}))
        plan = pm.describe_plan()
        assert (" eltwise eltwise " in plan) == bool(elt), plan
        pm.free()
    for i, target in enumerate(("y", "dx")):
        ref = om.call(target, {"x": x})
        assert_close(outs[1][i], ref, tol=2e-6, what=f"{act} {target} vs oracle")
        # same operations in the same order as the generic kernel
        assert_close(outs[1][i], outs[0][i], tol=1e-7, what=f"{act} {target} vs generic kernel")


def _optimizer_step(d, L, opt, n):
    p = d.param([n], name="p"); g = d.input("g", [n])
    eff = d.Fun("Effect", effect=p)
    opt(eff, g)
    return [eff.target("step", "gpu")]


@pytest.mark.parametrize("n", [10, 4096, 100003])
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_optimizer_updates(ctx, opt, n):
    """gradientDescent (base.nim:37-38) and adam (base.nim:40-53, caches + epoch()) as single streaming launches."""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    mk = (lambda L: L.gradient_descent(0.05)) if opt == "sgd" else (lambda L: L.adam(0.01))
    om = o.compile(*_optimizer_step(o, OL, mk(OL), n), seed=2)
    pm = M.compile(*_optimizer_step(F, PL, mk(PL), n), gpu=ctx, seed=2)
    tid = pm.params.ids()[0]
    p0 = np.random.default_rng(1).uniform(-1, 1, n).astype(np.float32)
    om.params[tid][...] = p0; pm.params[tid] = p0
    for step in range(3):
        g = np.random.default_rng(10 + step).uniform(-1, 1, n).astype(np.float32)
        om.epoch += 1; om.apply("step", {"g": g})
        pm.set_option("epoch", step + 1); pm.apply("step", {"g": g})
    assert " eltwise eltwise " in pm.describe_plan() and " interp " not in pm.describe_plan(), pm.describe_plan()
    if opt == "adam":   # first moment, second moment and step of a parameter run as one pass over it
        assert "in one pass" in pm.describe_plan() and pm.describe_plan().count(" eltwise eltwise ") == 1, pm.describe_plan()
    assert_close(pm.params[tid], om.params[tid], tol=2e-6, what=f"{opt} parameter after 3 steps")
    assert_close(pm.params[tid] - p0, om.params[tid] - p0, tol=1e-4, what=f"{opt} update")
    for cid in sorted(om.caches):
        assert_close(pm.caches[cid], om.caches[cid], tol=2e-6, what=f"{opt} cache")
    pm.free()


def test_bias_row_add_and_arithmetic(ctx):
    from exprgrad_b200 import frontend as F, layers as PL, model as M
    rng = np.random.default_rng(4)
    a = rng.uniform(-1, 1, (37, 64)).astype(np.float32); b = rng.uniform(-1, 1, (37, 64)).astype(np.float32)
    bias = rng.uniform(-1, 1, 64).astype(np.float32)

    def net():
        x = F.input("a", [-1, 64]); y = F.input("b", [-1, 64]); bp = F.param([64], name="bias")
        r = F.Fun(); i, j = F.Iter("y"), F.Iter("x")
        r[i, j] += x[i, j]
        i, j = F.Iter("y"), F.Iter("x")
        r[i, j] += bp[j]
        # a raw write with two reads carries no shape constraint (passes.nim:1059-1095): copyShape like a user would
        s2, d2 = PL.add(x, y), PL.sub(x, y)
        s2.copy_shape(x); d2.copy_shape(x)
        return [r.target("biased", "gpu"), s2.target("sum", "gpu"), d2.target("diff", "gpu"),
                PL.scale(x, 2.5).target("scaled", "gpu"), PL.divide(x, 3.0).target("divided", "gpu")]
    pm = M.compile(*net(), gpu=ctx, seed=0)
    pm.params[pm.params.ids()[0]] = bias
    assert np.array_equal(pm.call("biased", {"a": a}), a + bias)
    assert "bias-row-add" in pm.describe_plan()
    assert np.array_equal(pm.call("sum", {"a": a, "b": b}), a + b)
    assert np.array_equal(pm.call("diff", {"a": a, "b": b}), a - b)
    assert np.array_equal(pm.call("scaled", {"a": a}), a * np.float32(2.5))
    assert np.array_equal(pm.call("divided", {"a": a}), a / np.float32(3.0))
    pm.free()


@pytest.mark.parametrize("n_rows", [1, 33, 4096])
def test_square_adjoint_reads_one_seed_element(ctx, n_rows):
    """d[i] = s[k] * x[i] + s[k] * x[i] (derive of sq under a scalar loss, passes.nim:399-403) with the seed at a
    non-zero offset of its tensor, on sizes with and without a 4-element tail: streaming kernel vs numpy (the same two
    products and one addition in fp32) and vs the generic kernel."""
    from exprgrad_b200 import frontend as F, model as M
    rng = np.random.default_rng(n_rows)
    x = rng.uniform(-2, 2, (n_rows, 63)).astype(np.float32)
    s = rng.uniform(-2, 2, 4).astype(np.float32)

    def net():
        xi = F.input("x", [-1, 63]); si = F.input("s", [4])
        r = F.Fun(); it = F.Iter("it")
        r.raw[it] += si[2] * xi.raw[it] + xi.raw[it] * si[2]
        r.copy_shape(xi)
        return [r.target("y", "gpu")]
    want = s[2] * x + s[2] * x
    for elt in (1, 0):
        pm = M.compile(*net(), gpu=ctx, seed=0)
        pm.set_option("eltwise", elt)
        got = pm.call("y", {"x": x, "s": s})
        assert ("square-adjoint" in pm.describe_plan()) == bool(elt), pm.describe_plan()
        assert np.array_equal(got, want), f"eltwise={elt}: max diff {np.abs(got - want).max()}"
        pm.free()


def test_sigmoid_head_runs_in_the_contraction_epilogue(ctx):
    """The xor net's sigmoid (examples/xor/xor.nim:20-28) and a tanh layer fuse into the contraction that feeds
    them, like relu / leakyRelu do; results match the unfused plan and the oracle."""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL, model as M

    def net(d, L):
        x = d.input("x", [-1, 24])
        h = L.tanh(L.dense(x, 24, 16))
        return [L.sigmoid(L.dense(h, 16, 8)).target("predict", "gpu")]
    om = o.compile(*net(o, OL), seed=3)
    x = np.random.default_rng(0).uniform(-1, 1, (50, 24)).astype(np.float32)
    ref = om.call("predict", {"x": x})
    outs = []
    for fuse in (1, 0):
        pm = M.compile(*net(F, PL), gpu=ctx, seed=3)
        pm.set_option("fuse", fuse)
        for tid in sorted(om.params):
            pm.params[tid] = om.params[tid]
        outs.append(pm.call("predict", {"x": x}))
        plan = pm.describe_plan()
        if fuse:
            assert plan.count("+2 fused") == 2 and " eltwise eltwise " not in plan, plan   # bias + activation per layer
        pm.free()
    assert_close(outs[0], ref, tol=1e-5, what="fused tanh/sigmoid epilogues vs oracle")
    assert_close(outs[0], outs[1], tol=1e-6, what="fused vs unfused")


def test_unrecognised_softmax_variant_falls_back_with_a_note(ctx):
    """A max-subtracted ("safe") softmax is not the reference's softmax (dnn.nim:90-94): the fused row kernel must
    not claim it; the plan still runs (generic kernels) and matches numpy."""
    from exprgrad_b200 import frontend as F, model as M
    x = F.input("x", [-1, 10])
    mx = F.Fun(); i, j = F.Iter("y"), F.Iter("x")
    mx[i] += F.select(x[i, j] >= 100.0, x[i, j], 0.0)          # stand-in for a row statistic that is subtracted first
    sums = F.Fun(); i, j = F.Iter("y"), F.Iter("x")
    sums[i] += F.exp(x[i, j] - mx[i])
    r = F.Fun(); i, j = F.Iter("y"), F.Iter("x")
    r[i, j] += F.exp(x[i, j] - mx[i]) / sums[i]
    pm = M.compile(r.target("p", "gpu"), gpu=ctx, seed=0)
    v = np.random.default_rng(0).uniform(-2, 2, (64, 10)).astype(np.float32)
    got = pm.call("p", {"x": v})
    e = np.exp(v.astype(np.float64))
    assert_close(got, e / e.sum(1, keepdims=True), tol=1e-5, what="shifted softmax")
    assert "softmax_xent" not in pm.describe_plan()
    pm.free()
