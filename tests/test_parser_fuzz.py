"""CPU-only: damaged program text is an error, never a crash. The text form of `Program` crosses a process boundary
(the Nim front-end writes it, checkpoints store it - io/serialize.nim:323-379 on the reference side), so the parser
(csrc/program.cpp) checks counts, 32-bit fields and the referential integrity of what it read (tensor ids, registers,
enumerations) before the passes and the planner index with it. Token-level mutations of valid source and compiled
programs go through parse / compile / serialize / describe / classify / lower_dump / infer_shapes in a child process."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mutated_program_text_never_crashes(seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "parse_fuzz_worker.py"), str(seed), "250"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, f"worker died with {r.returncode}\n{r.stderr[-1500:]}"
    last = r.stdout.strip().splitlines()[-1].split()
    assert last[0] == "done" and int(last[1]) + int(last[2]) == 250
    assert int(last[2]) > 100          # most mutations are rejected ...
    assert int(last[1]) > 0            # ... and the harmless ones (a changed literal, a renamed tensor) still parse


def test_reference_errors_for_dangling_ids():
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as PL
    from exprgrad_b200.model import Program
    import graphs as G
    text = Program.from_graphs(G.matmul(F, PL)).serialize()
    toks = text.split()
    i = toks.index("R")
    toks[i + 1] = "99"                        # a read of tensor 99 of 3
    with pytest.raises(eg.ParserError):
        Program(" ".join(toks))
    toks = text.split()
    toks[toks.index("tensors") + 1] = "2000000000"   # an element count beyond the text
    with pytest.raises(eg.ParserError):
        Program(" ".join(toks))
    # a Fun is bound to the first program it is compiled into (parser.nim:261-262 `if fun.tensor == TensorId(0)`): reusing
    # it in a second program leaves a dangling id - an error here, not an out-of-bounds access
    x = F.input("x", [-1, 4]); y = F.input("y", [-1, 4])
    Program.from_graphs([PL.add(x, y).target("t")])
    with pytest.raises(eg.ParserError):
        Program.from_graphs([PL.sub(x, y).target("t")])
