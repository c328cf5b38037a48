"""CPU-only differential test of the native passes (csrc/passes.cpp: autodiff `generate` / `derive`, dead-code and
dead-kernel elimination, shape-constraint construction and sorting, run-time shape inference) against the oracle's
restatement of exprgrad/passes.nim on seeded random graphs (tests/fuzz_graphs.py): the compiled programs must be
token-identical - kernel order, register numbering of every adjoint, constraints - and the inferred shapes of
every tensor of every target bit-exact."""
import pytest

import fuzz_graphs as FG
from test_passes_parity import _tokens

SEEDS = list(range(120))


def _build(seed):
    import oracle as o
    from oracle import layers as OL
    from oracle.passes import compile_program
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    ographs, what = FG.random_net(o, OL, seed)
    oprog = o.ir.to_program(ographs)
    compile_program(oprog)
    pgraphs, what2 = FG.random_net(F, PL, seed)
    assert what == what2
    prog = Program.from_graphs(pgraphs).compile()
    return o, oprog, F, prog, what


@pytest.mark.parametrize("seed", SEEDS)
def test_random_graph_compiles_to_the_same_program(seed):
    o, oprog, F, prog, what = _build(seed)
    want = _tokens(F.serialize(oprog, compiled=True))
    got = _tokens(prog.serialize())
    if "grads" in got:      # the library's gradient table (bookkeeping behind the reference's Program, see test_passes_parity)
        i = len(got) - 1 - got[::-1].index("grads")
        got = got[:i] + got[-1:]
    assert len(want) == len(got), what
    for i, (x, y) in enumerate(zip(want, got)):
        assert x == y, f"{what}: token {i}: oracle {want[max(0, i - 8):i + 4]} vs library {got[max(0, i - 8):i + 4]}"


@pytest.mark.parametrize("seed", SEEDS[::3])
@pytest.mark.parametrize("rows", [1, 7])
def test_random_graph_shapes_bit_exact(seed, rows):
    from oracle.passes import infer_shapes
    o, oprog, F, prog, what = _build(seed)
    inputs = {"a": [rows, FG.COLS], "b": [rows, FG.COLS], "v": [FG.COLS]}
    for target in oprog.targets:
        used = {k: v for k, v in inputs.items() if oprog.inputs.get(k) in oprog.targets[target].tensors}
        want = infer_shapes(oprog, target, {oprog.inputs[k]: v for k, v in used.items()})
        for tid in sorted(want):
            got = prog.infer_shapes(target, used, tensor_id=tid)
            assert got == list(want[tid]), f"{what} / {target}: tensor{tid - 1}: oracle {want[tid]} vs library {got}"


CNN_SEEDS = list(range(60))


def _build_cnn(seed):
    import oracle as o
    from oracle import layers as OL
    from oracle.passes import compile_program
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    ographs, what, shapes = FG.random_cnn(o, OL, seed)
    oprog = o.ir.to_program(ographs)
    compile_program(oprog)
    pgraphs, what2, _ = FG.random_cnn(F, PL, seed)
    assert what == what2
    prog = Program.from_graphs(pgraphs).compile()
    return o, oprog, F, prog, what, shapes


@pytest.mark.parametrize("seed", CNN_SEEDS)
def test_random_layer_stack_compiles_to_the_same_program_and_shapes(seed):
    """Random stacks of the reference's layers (customGrad pooling, `withShape` upsampling, dropout's random tensor,
    the reshape generator, adam's caches and epoch()): token-identical compiled programs, bit-exact shapes."""
    from oracle.passes import infer_shapes
    o, oprog, F, prog, what, shapes = _build_cnn(seed)
    want = _tokens(F.serialize(oprog, compiled=True))
    got = _tokens(prog.serialize())
    if "grads" in got:
        i = len(got) - 1 - got[::-1].index("grads")
        got = got[:i] + got[-1:]
    assert len(want) == len(got), what
    for i, (x, y) in enumerate(zip(want, got)):
        assert x == y, f"{what}: token {i}: oracle {want[max(0, i - 8):i + 4]} vs library {got[max(0, i - 8):i + 4]}"
    inputs = {"x": shapes["x"], "y": [shapes["x"][0], shapes["outs"]]}
    for target in oprog.targets:
        used = {k: v for k, v in inputs.items() if oprog.inputs.get(k) in oprog.targets[target].tensors}
        want_shapes = infer_shapes(oprog, target, {oprog.inputs[k]: v for k, v in used.items()})
        for tid in sorted(want_shapes):
            got_shape = prog.infer_shapes(target, used, tensor_id=tid)
            assert got_shape == list(want_shapes[tid]), f"{what} / {target}: tensor{tid - 1}: {want_shapes[tid]} vs {got_shape}"


@pytest.mark.parametrize("seed", list(range(60)))
def test_random_index_graph_compiles_to_the_same_program_and_shapes(seed):
    """Graphs whose indices do the work (explicit loop bounds, strided / divided / modular / wrapped accesses, array
    literals, scatter writes, `withShape`): token-identical programs, bit-exact shapes at two batch sizes - or the same
    error class where the reference's `derive` has no rule (ArrayRead)."""
    import oracle as o
    from oracle import layers as OL
    from oracle.passes import compile_program, infer_shapes
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    ographs, what = FG.random_index_net(o, OL, seed)
    oprog = o.ir.to_program(ographs)
    pgraphs, what2 = FG.random_index_net(F, PL, seed)
    assert what == what2
    try:
        compile_program(oprog)
    except (o.GradientError, o.ShapeError) as e:     # e.g. a scatter write leaves its tensor's shape underconstrained
        with pytest.raises(eg.GradientError if isinstance(e, o.GradientError) else eg.ShapeError):
            Program.from_graphs(pgraphs).compile()
        return
    prog = Program.from_graphs(pgraphs).compile()
    want = _tokens(F.serialize(oprog, compiled=True))
    got = _tokens(prog.serialize())
    if "grads" in got:
        i = len(got) - 1 - got[::-1].index("grads")
        got = got[:i] + got[-1:]
    assert len(want) == len(got), what
    for i, (x, y) in enumerate(zip(want, got)):
        assert x == y, f"{what}: token {i}: oracle {want[max(0, i - 8):i + 4]} vs library {got[max(0, i - 8):i + 4]}"
    for rows in (1, 5):
        inputs = {"a": [rows, 12], "v": [12]}
        for target in oprog.targets:
            used = {k: v for k, v in inputs.items() if oprog.inputs.get(k) in oprog.targets[target].tensors}
            want_shapes = infer_shapes(oprog, target, {oprog.inputs[k]: v for k, v in used.items()})
            for tid in sorted(want_shapes):
                got_shape = prog.infer_shapes(target, used, tensor_id=tid)
                assert got_shape == list(want_shapes[tid]), f"{what} / {target}: tensor{tid - 1}: {want_shapes[tid]} vs {got_shape}"
