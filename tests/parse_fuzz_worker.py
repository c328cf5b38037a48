"""Worker of tests/test_parser_fuzz.py (run in a child process: a crash must fail the test, not kill pytest):
mutates valid program texts token-wise and drives every host-only entry point with the result. Prints 'done <ok> <err>'."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import graphs as G  # noqa: E402
from exprgrad_b200 import frontend as F, layers as PL  # noqa: E402
from exprgrad_b200.model import Program  # noqa: E402

seed0, n = int(sys.argv[1]), int(sys.argv[2])
texts = []
for name in ("dense_net", "conv2_net", "fashion_net", "xor_net", "matmul"):
    p = Program.from_graphs(G.ALL[name](F, PL))
    texts.append(p.serialize())           # source program (stage 0)
    texts.append(p.compile().serialize())  # compiled program (stage 1)
SHAPES = {"x": [4, 784], "y": [4, 10], "a": [4, 4], "b": [4, 4], "img": [1, 6, 6, 3]}
WORDS = ["K", "L", "R", "W", "T", "end", "target", "S", "I", "C", "LI", "copy", "dims", "linear", "rank", "grads", "G"]
BIG = ["99999999999", "2147483647", "-2147483648", "65536", "1000000", "-1", "0"]
ok = err = 0
for i in range(n):
    rng = random.Random(seed0 * 100000 + i)
    toks = rng.choice(texts).split()
    for _ in range(rng.randint(1, 4)):
        j = rng.randrange(len(toks))
        op = rng.choice(["del", "dup", "num", "swap", "trunc", "neg", "big", "word"])
        if op == "del": del toks[j]
        elif op == "dup": toks.insert(j, toks[j])
        elif op == "num": toks[j] = str(rng.randint(-3, 300))
        elif op == "swap" and j + 1 < len(toks): toks[j], toks[j + 1] = toks[j + 1], toks[j]
        elif op == "trunc": toks = toks[:j + 1]
        elif op == "neg": toks[j] = "-" + toks[j]
        elif op == "big": toks[j] = rng.choice(BIG)
        elif op == "word": toks[j] = rng.choice(WORDS)
        if not toks:
            toks = ["x"]
    try:
        p = Program(" ".join(toks))
        p.compile()
        p.serialize()
        for t in ("train", "c", "conv", "predict"):
            try:
                p.describe(t)
                p.classify(t, SHAPES)
                p.lower_dump(t, SHAPES)
                for tid in range(1, min(p.tensor_count(), 6) + 1):
                    p.infer_shapes(t, SHAPES, tensor_id=tid)
            except Exception:
                pass
        ok += 1
    except Exception:
        err += 1
print("done", ok, err)
