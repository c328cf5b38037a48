"""CPU-only: what csrc/lower.cpp hands to the generic loop-nest kernel, executed by an independent sequential
interpreter (tests/ip_interp.py) and compared with the oracle - on the reference-shaped graphs and on seeded random
ones. Checks the lowering (loop partition, flattened accesses, index arithmetic for `y div 2`-style
indices, register program, literal pool, f64 constant folding like passes.nim:1656-1706) without a device; the device
kernel that interprets the same programs is compared with the same oracle in the GPU tier."""
import numpy as np
import pytest

import fuzz_graphs as FG
import graphs as G
from ip_interp import run_target
from parity_cases import norm_err

# the same operations in the same order as the oracle; numpy's fp32 sin / exp / log / pow may differ from glibc's in
# the last ulp, and an ill-conditioned random graph amplifies that like any rounding (bound scales with the oracle's own
# fp32-vs-float64 difference, as in tests/test_gpu_fuzz.py)
TOL = 2e-6


def _oracle(graph_fn, seed_params, inputs, scalar="float32"):
    import oracle as o
    from oracle import layers as OL
    om = o.compile(*graph_fn(o, OL), scalar=scalar, seed=1, openmp=False)
    return om


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", list(range(0, 36)))
def test_random_graph_lowering_matches_oracle(seed, strict):
    if not strict and seed % 3:
        pytest.skip("the non-strict lowering (loop merging, fast-path flags) runs on every third graph")
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    res = {}
    for scalar in ("float32", "float64"):
        graphs, what = FG.random_net(o, OL, seed, ct="cpu")
        om = o.compile(*graphs, scalar=scalar, seed=0, openmp=False)
        inputs = {k: v for k, v in FG.random_inputs(np, seed, rows=4).items() if k in om.program.inputs}
        names = {om.program.tdef(t).name: t for t in om.params}
        for k, v in FG.random_params(np, seed).items():
            if k in names:
                om.params[names[k]][...] = v
        out = {t: np.array(om.call(t, inputs)) for t in om.program.targets if t != "train"}
        if "train" in om.program.targets:
            om.apply("train", inputs)
            for tid in sorted(om.params):
                out[f"param{tid}"] = np.array(om.params[tid])
        res[scalar] = out
    ref, ref64 = res["float32"], res["float64"]
    prog = Program.from_graphs(FG.random_net(F, PL, seed)[0]).compile()
    params = FG.random_params(np, seed)
    state0 = {tid: params[k].copy() for k, tid in names.items()}
    for t in sorted(ref):
        if t.startswith("param"):
            continue
        got = run_target(prog, t, inputs, dict(state0), strict=strict)
        _compare(got, ref[t], ref64[t], f"seed {seed} ({what}) target {t} strict={strict}")
    if any(t.startswith("param") for t in ref):
        state = {tid: v.copy() for tid, v in state0.items()}
        run_target(prog, "train", inputs, state, strict=strict)
        for tid in state:
            _compare(state[tid], ref[f"param{tid}"], ref64[f"param{tid}"], f"seed {seed} ({what}) param{tid} strict={strict}")


def _compare(got, ref, ref64, what):
    if not np.any(ref64):
        assert not np.any(got), what + ": expected zeros"
        return
    cond = norm_err(ref, ref64)
    e = norm_err(got, ref)
    assert e <= max(TOL, 10 * cond), f"{what}: normalised max error {e:.2e} (conditioning {cond:.1e})"


def test_reference_shaped_graphs_lowering_matches_oracle():
    """dense step (contractions, bias, relu, softmax + crossEntropy, SGD), conv2 forward / d_filters / d_images (scatter
    accumulate), conv + leakyRelu + maxpool (customGrad, `y div 2` indices) + reshape + adam with epoch()."""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    # dense step
    sizes = (6, 5, 4, 3)
    om = o.compile(*G.dense_net(o, OL, sizes, ct="cpu"), seed=1, openmp=False)
    x, y, params = G.dense_inputs(5, sizes)
    tids = sorted(om.params)
    for i, tid in enumerate(tids):
        om.params[tid][...] = params[i]
    prog = Program.from_graphs(G.dense_net(F, PL, sizes)).compile()
    state = {tid: params[i].copy() for i, tid in enumerate(tids)}
    assert norm_err(run_target(prog, "predict", {"x": x}, dict(state)), om.call("predict", {"x": x})) <= TOL
    assert norm_err(run_target(prog, "loss", {"x": x, "y": y}, dict(state)), om.call("loss", {"x": x, "y": y})) <= TOL
    om.apply("train", {"x": x, "y": y})
    run_target(prog, "train", {"x": x, "y": y}, state)
    for tid in tids:
        assert norm_err(state[tid], om.params[tid]) <= TOL, f"dense param tensor{tid - 1}"
    # conv2: no libm call anywhere -> bit for bit
    om = o.compile(*G.conv2_net(o, OL, ct="cpu"), seed=1, openmp=False)
    w = np.random.default_rng(1).uniform(-2, 2, (4, 3, 3, 3)).astype(np.float32)
    tid = sorted(om.params)[0]
    om.params[tid][...] = w
    img = np.random.default_rng(0).uniform(0, 1, (2, 6, 5, 3)).astype(np.float32)
    prog = Program.from_graphs(G.conv2_net(F, PL)).compile()
    for t in ("conv", "loss", "dw", "dimg"):
        got = run_target(prog, t, {"img": img}, {tid: w.copy()})
        assert np.array_equal(got, om.call(t, {"img": img})), f"conv2 {t}"
    # fashion net, adam, two epochs
    om = o.compile(*G.fashion_net(o, OL, ct="cpu"), seed=1, openmp=False)
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (3, 12, 12, 1)).astype(np.float32)
    y = np.zeros((3, 10), np.float32); y[np.arange(3), rng.integers(0, 10, 3)] = 1
    prog = Program.from_graphs(G.fashion_net(F, PL)).compile()
    state = {tid: np.array(v) for tid, v in om.params.items()}
    state.update({tid: np.array(v) for tid, v in om.caches.items()})
    for epoch in (1, 2):
        om.epoch = epoch
        om.apply("train", {"x": x, "y": y})
        run_target(prog, "train", {"x": x, "y": y}, state, epoch=epoch)
    for tid in om.params:
        assert norm_err(state[tid], om.params[tid]) <= 1e-5, f"adam param tensor{tid - 1}"
    for tid in om.caches:
        assert norm_err(state[tid], om.caches[tid]) <= 1e-5, f"adam cache tensor{tid - 1}"


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_layer_stack_lowering_matches_oracle(seed):
    _layer_stack_case(seed, strict=True)


@pytest.mark.parametrize("seed", list(range(1, 24, 2)))
def test_random_layer_stack_default_lowering_matches_oracle(seed):
    """the same through the DEFAULT lowering (what a device runs unless strict mode is on): merged loops, fast-path flags"""
    _layer_stack_case(seed, strict=False)


def _layer_stack_case(seed, strict):
    """Random stacks of the reference's layers (tests/fuzz_graphs.py random_cnn): strided and divided indices of the
    pooling layers, the customGrad adjoint of maxpool2, `withShape` upsampling, the reshape generator, adam's caches and
    epoch() - one train step through the lowered programs against the oracle."""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    graphs, what, sh = FG.random_cnn(o, OL, seed, ct="cpu")
    if "dropout" in what:
        pytest.skip("dropout draws a random tensor per call (model.nim:310-314): nothing to compare element-wise")
    rng = np.random.default_rng(seed)
    n = 2
    x = rng.uniform(0, 1, [n] + sh["x"][1:]).astype(np.float32)
    y = rng.uniform(0.1, 0.9, (n, sh["outs"])).astype(np.float32)
    res = {}
    for scalar in ("float32", "float64"):
        om = o.compile(*FG.random_cnn(o, OL, seed, ct="cpu")[0], scalar=scalar, seed=3, openmp=False)
        out = {"predict": np.array(om.call("predict", {"x": x})), "loss": np.array(om.call("loss", {"x": x, "y": y}))}
        start = {tid: np.array(v, np.float32) for tid, v in om.params.items()}
        om.epoch = 1
        om.apply("train", {"x": x, "y": y})
        for tid in om.params:
            out[f"param{tid}"] = np.array(om.params[tid])
        for tid in om.caches:
            out[f"cache{tid}"] = np.array(om.caches[tid])
        res[scalar] = (out, start, sorted(om.caches))
    (ref, start, cache_ids), (ref64, _, _) = res["float32"], res["float64"]
    prog = Program.from_graphs(FG.random_cnn(F, PL, seed)[0]).compile()
    state = {tid: v.copy() for tid, v in start.items()}
    for tid in cache_ids:
        state[tid] = np.zeros(prog.tensor_info(tid)["shape"], np.float32)
    _compare(run_target(prog, "predict", {"x": x}, dict(state), strict=strict), ref["predict"], ref64["predict"], f"cnn {seed} ({what}) predict")
    _compare(run_target(prog, "loss", {"x": x, "y": y}, dict(state), strict=strict), ref["loss"], ref64["loss"], f"cnn {seed} ({what}) loss")
    run_target(prog, "train", {"x": x, "y": y}, state, strict=strict, epoch=1)
    for tid in start:
        _compare(state[tid], ref[f"param{tid}"], ref64[f"param{tid}"], f"cnn {seed} ({what}) param tensor{tid - 1}")
    for tid in cache_ids:
        _compare(state[tid], ref[f"cache{tid}"], ref64[f"cache{tid}"], f"cnn {seed} ({what}) cache tensor{tid - 1}")


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", list(range(40)))
def test_random_index_graph_lowering_matches_oracle(seed, strict):
    """explicit loop bounds, strided / divided / modular / wrapped accesses, iterator and shape values in expressions,
    array literals, scatter writes: value, loss and gradient through the lowered programs, both lowering modes"""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    rng = np.random.default_rng(seed)
    full = {"a": rng.uniform(-1, 1, (4, 12)).astype(np.float32), "v": rng.uniform(-1, 1, 12).astype(np.float32)}
    res = {}
    for scalar in ("float32", "float64"):
        graphs, what = FG.random_index_net(o, OL, seed, ct="cpu")
        try:
            om = o.compile(*graphs, scalar=scalar, seed=0, openmp=False)
        except (o.GradientError, o.ShapeError):
            pytest.skip("a compile-time error of the reference (no derive rule for ArrayRead, underconstrained scatter "
                        "shape): error parity is checked in tests/test_passes_fuzz.py")
        inputs = {k: v for k, v in full.items() if k in om.program.inputs}
        res[scalar] = {t: np.array(om.call(t, inputs)) for t in ("out", "loss", "da")}
    prog = Program.from_graphs(FG.random_index_net(F, PL, seed)[0]).compile()
    for t in ("out", "loss", "da"):
        got = run_target(prog, t, inputs, {}, strict=strict)
        _compare(got, res["float32"][t], res["float64"][t], f"index graph {seed} ({what}) target {t} strict={strict}")


def test_rarely_used_opcodes_lowering_matches_oracle():
    """Boolean connectives (dsl.nim:48-50 - `or` builds InstrAnd in the reference, restated as is), Eq on scalars, ToIndex
    (fptosi truncation towards zero, llvmgen.nim:229-236), Mod / IndexDiv on a computed index, pow with a tensor
    exponent, nested selects."""
    import oracle as o
    from oracle import layers as OL  # noqa: F401
    from exprgrad_b200 import frontend as F
    from exprgrad_b200.model import Program

    def net(d):
        a = d.input("a", [-1, 6]); b = d.input("b", [-1, 6]); v = d.input("v", [6])
        graphs = []
        r = d.Fun(); y, x = d.Iter("y"), d.Iter("x")
        r[y, x] += d.select((a[y, x] < 0.25).and_(b[y, x] < 0.5), a[y, x] * 2.0, d.select((a[y, x] < -0.5).or_(b[y, x] <= 0.0), b[y, x], -a[y, x]))
        graphs.append(r.target("bools", "cpu"))
        r = d.Fun(); y, x = d.Iter("y"), d.Iter("x")
        r[y, x] += d.select(a[y, x].eq(b[y, x]), d.lift(1.0), d.lift(0.0)) + d.to_scalar(d.to_index(a[y, x] * 3.7)) * v[x]
        graphs.append(r.target("toindex", "cpu"))
        r = d.Fun(); y, x = d.Iter("y"), d.Iter("x")
        r[y, x] += a[y, (x * 5 + 1) % 6] * v[(x + 3) // 2]
        r.copy_shape(a)          # x stands alone only in the write: nothing else would bound it
        graphs.append(r.target("modidx", "cpu"))
        r = d.Fun(); it = d.Iter("it")
        r.raw[it] += d.pow_(a.raw[it] * a.raw[it] + 1.0, b.raw[it])
        r.copy_shape(a)
        loss = d.Fun(); it = d.Iter("it"); loss[0] += r.raw[it]
        graphs += [r.target("pow", "cpu"), loss.backwards().grad(b).target("dpow_db", "cpu"),
                   loss.backwards().grad(a).target("dpow_da", "cpu")]
        return graphs
    rng = np.random.default_rng(7)
    a = rng.uniform(-1, 1, (5, 6)).astype(np.float32); b = rng.uniform(-1, 1, (5, 6)).astype(np.float32)
    a[0, 0] = b[0, 0]; a[1, 2] = b[1, 2]
    v = rng.uniform(-1, 1, 6).astype(np.float32)
    om = o.compile(*net(o), seed=0, openmp=False)
    prog = Program.from_graphs(net(F)).compile()
    full = {"a": a, "b": b, "v": v}
    for target in ("bools", "toindex", "modidx", "pow", "dpow_db", "dpow_da"):
        used = {k: full[k] for k in full if om.program.inputs.get(k) in om.program.targets[target].tensors}
        ref = np.array(om.call(target, used))
        for strict in (True, False):
            got = run_target(prog, target, used, {}, strict=strict)
            if target in ("bools", "toindex", "modidx"):
                assert np.array_equal(got, ref), f"{target} strict={strict}"
            else:
                assert norm_err(got, ref) <= 5e-6, f"{target} strict={strict}: {norm_err(got, ref):.2e}"


def test_empty_batch_through_the_host_side():
    """zero rows: shape inference gives [0, n] results, every lowered program has zero points (or a zero-trip reduction),
    the loss is 0 and a train step leaves the parameters alone - like the oracle"""
    import oracle as o
    from oracle import layers as OL
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    sizes = (6, 5, 4, 3)
    om = o.compile(*G.dense_net(o, OL, sizes, ct="cpu"), seed=1, openmp=False)
    x = np.zeros((0, 6), np.float32); y = np.zeros((0, 3), np.float32)
    prog = Program.from_graphs(G.dense_net(F, PL, sizes)).compile()
    state = {tid: np.array(v) for tid, v in om.params.items()}
    got = run_target(prog, "predict", {"x": x}, dict(state))
    assert got.shape == om.call("predict", {"x": x}).shape == (0, 3)
    assert np.array_equal(run_target(prog, "loss", {"x": x, "y": y}, dict(state)), om.call("loss", {"x": x, "y": y}))
    before = {k: v.copy() for k, v in state.items()}
    run_target(prog, "train", {"x": x, "y": y}, state)
    om.apply("train", {"x": x, "y": y})
    for tid in state:
        assert np.array_equal(state[tid], before[tid]) and np.array_equal(om.params[tid], before[tid])
