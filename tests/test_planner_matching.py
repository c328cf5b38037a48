"""CPU-only checks of the planner's structural matchers (csrc/pattern.cpp) through egb_program_classify: the
fixed forms of exprgrad/layers/base.nim and dnn.nim and of their derive()d adjoints are recognised from the
instruction DAG and the index pattern - not from text - so operand order of commutative operations and the
order of reads do not matter, while anything else stays on the generic loop-nest kernel."""
import graphs as G


def _classify(graphs, target, shapes):
    from exprgrad_b200 import frontend as F  # noqa: F401
    from exprgrad_b200.model import Program
    return Program.from_graphs(graphs).compile().classify(target, shapes)


def test_dense_net_train_kernels():
    from exprgrad_b200 import frontend as F, layers as PL
    got = _classify(G.dense_net(F, PL), "train", {"x": [1024, 784], "y": [1024, 10]})
    assert got[0] == "contraction 1024 512 784 NN"
    assert got[1] == "eltwise bias-row-add n=524288 row=512"
    assert got[2] == "eltwise relu n=524288"
    assert got[16] == "contraction 1024 512 10 NT" and got[17] == "contraction 512 10 1024 TN"
    assert got[18] == "eltwise relu-adjoint n=524288"
    assert got[25:] == [f"eltwise sgd-axpy n={n}" for n in (401408, 512, 262144, 512, 5120, 10)]
    assert all(g == "generic" for g in got[8:16])      # softmax / crossEntropy rows: fused_rows.cu or the row chain


def test_adam_conv_and_pool_layers():
    from exprgrad_b200 import frontend as F, layers as PL
    got = _classify(G.fashion_net(F, PL), "train", {"x": [64, 12, 12, 1], "y": [64, 10]})
    assert got[0] == "conv2 forward" and got[19] == "conv2 d_filters"
    assert got[1] == "eltwise leakyRelu n=25600" and got[18] == "eltwise leakyRelu-adjoint n=25600"
    assert got[2] == "generic"                          # maxpool2: four strided reads
    assert got[20:23] == ["eltwise adam-m n=36", "eltwise adam-v n=36", "eltwise adam-step n=36"]


def _unary(expr_of):
    from exprgrad_b200 import frontend as F
    x = F.input("x", [-1, 64])
    r = F.Fun(); it = F.Iter("it")
    r.raw[it] += expr_of(F, x.raw[it])
    r.copy_shape(x)
    return [r.target("y", "gpu")]


def test_operand_order_does_not_matter():
    shapes = {"x": [32, 64]}
    # leakyRelu written as x * select(...) instead of select(...) * x (dnn.nim:29-30)
    got = _classify(_unary(lambda F, x: x * F.select(x >= 0.0, 1.0, 0.2)), "y", shapes)
    assert got == ["eltwise leakyRelu n=2048"]
    # sigmoid with the addition swapped: 1 / (exp(-x) + 1)
    got = _classify(_unary(lambda F, x: 1.0 / (F.exp(-x) + 1.0)), "y", shapes)
    assert got == ["eltwise sigmoid n=2048"]
    # scale with the literal first
    got = _classify(_unary(lambda F, x: 3.0 * x), "y", shapes)
    assert got == ["eltwise scale n=2048"]


def test_other_expressions_stay_generic():
    shapes = {"x": [32, 64]}
    # relu6-like clamp, a shifted relu and a different comparison are NOT the reference's relu
    assert _classify(_unary(lambda F, x: F.select(x >= 0.0, F.select(x <= 6.0, x, 6.0), 0.0)), "y", shapes) == ["generic"]
    assert _classify(_unary(lambda F, x: F.select(x >= 1.0, x, 0.0)), "y", shapes) == ["generic"]
    assert _classify(_unary(lambda F, x: F.select(x > 0.0, x, 0.0)), "y", shapes) == ["generic"]


def test_transposed_and_strided_accesses_are_not_maps():
    from exprgrad_b200 import frontend as F, layers as PL
    m = F.input("m", [-1, -1])
    assert _classify([PL.transpose(m).target("t", "gpu")], "t", {"m": [8, 16]}) == ["generic"]
    img = F.input("img", [-1, 8, 8, 2])
    assert _classify([PL.avgpool2(img).target("p", "gpu")], "p", {"img": [2, 8, 8, 2]}) == ["generic"]


def _with_seed(expr_of, index=0):
    from exprgrad_b200 import frontend as F
    x = F.input("x", [-1, 64]); s = F.input("s", [4])
    r = F.Fun(); it = F.Iter("it")
    r.raw[it] += expr_of(x.raw[it], s[index])
    r.copy_shape(x)
    return [r.target("y", "gpu")]


def test_fixed_element_operand_of_the_square_adjoint():
    """derive() of sq(x) under a scalar loss multiplies the seed dL[0] into both products (passes.nim:399-403): the
    seed is ONE element, whatever the order of the products' operands; every other form with a fixed-element
    operand stays on the generic kernel."""
    shapes = {"x": [32, 64], "s": [4]}
    for f in (lambda x, s: s * x + s * x, lambda x, s: x * s + x * s, lambda x, s: s * x + x * s):
        assert _classify(_with_seed(f), "y", shapes) == ["eltwise square-adjoint n=2048"]
    assert _classify(_with_seed(lambda x, s: s * x + s * x, index=3), "y", shapes) == ["eltwise square-adjoint n=2048"]
    assert _classify(_with_seed(lambda x, s: x * s), "y", shapes) == ["generic"]
    assert _classify(_with_seed(lambda x, s: x + s), "y", shapes) == ["generic"]
    assert _classify(_with_seed(lambda x, s: s * x + s * s), "y", shapes) == ["generic"]


def _classify_dropping_unused(prog, target, shapes):
    """inputs a random graph did not end up using are not inputs of the model (model.nim:395-396)"""
    shapes = dict(shapes)
    while True:
        try:
            return prog.classify(target, shapes)
        except Exception as e:
            if "is not an input to the model" not in str(e):
                raise
            del shapes[str(e).split()[0]]


def test_every_kernel_of_random_graphs_gets_a_class():
    """The matchers (contraction, conv2, map forms) look at arbitrary kernels - random expression trees, shifted and
    strided reads, customGrad adjoints, several writers of one tensor - and must answer, never throw."""
    import fuzz_graphs as FG
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    seen = set()
    for seed in range(60):
        graphs, what = FG.random_net(F, PL, seed)
        prog = Program.from_graphs(graphs).compile()
        for target in ("out", "loss", "da"):
            got = _classify_dropping_unused(prog, target, {"a": [7, FG.COLS], "b": [7, FG.COLS], "v": [FG.COLS]})
            assert got and all(isinstance(g, str) and g for g in got), (seed, what, target)
            seen.update(g.split(" ")[0] for g in got)
    for seed in range(30):
        graphs, what, sh = FG.random_cnn(F, PL, seed)
        prog = Program.from_graphs(graphs).compile()
        for target in ("predict", "loss", "train"):
            got = prog.classify(target, {"x": sh["x"], "y": [sh["x"][0], sh["outs"]]})
            assert got and all(isinstance(g, str) and g for g in got), (seed, what, target)
            seen.update(g.split(" ")[0] for g in got)
    assert {"generic", "eltwise", "contraction", "conv2"} <= seen
