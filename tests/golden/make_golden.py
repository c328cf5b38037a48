"""Regenerates the golden vectors under tests/golden/ from the oracle (the CPU restatement of the reference's path).

The reference itself cannot produce them here: it needs Nim >= 1.6 and LLVM 13, neither of which is in this image
(DESIGN.md 2). The oracle is pinned on the reference's own known-answer tests (tests/test_reference_vectors.py), on
closed-form numpy (tests/test_oracle_numpy.py) and on finite differences (tests/test_oracle_fuzz.py); these files
freeze its results for the reduced-size configurations of BASELINE.json so that
  * the oracle itself is checked for drift between hosts (another gcc / glibc / CPU: tests/test_golden.py, CPU tier),
  * the device path is compared with committed numbers as well as with the live oracle (tests/test_zgolden_gpu.py).

usage: python tests/golden/make_golden.py        (writes *.npz + shapes.json next to this file)"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import graphs as G  # noqa: E402

DENSE_SIZES = (64, 48, 32, 10)


def cases(o, OL):
    """name -> dict of arrays (inputs, injected parameters, results)"""
    out = {}
    # C2 shape class: a ragged contraction on mixed-sign data
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, (37, 53)).astype(np.float32); b = rng.uniform(-1, 1, (53, 29)).astype(np.float32)
    m = o.compile(*G.matmul(o, OL, ct="cpu"))
    out["matmul"] = {"a": a, "b": b, "c": m.call("c", {"a": a, "b": b})}
    # C3 at reduced size: predict, loss, parameters after two SGD steps
    m = o.compile(*G.dense_net(o, OL, DENSE_SIZES, ct="cpu"), seed=1)
    x, y, params = G.dense_inputs(96, DENSE_SIZES)
    d = {"x": x, "y": y}
    for i, tid in enumerate(sorted(m.params)):
        m.params[tid][...] = params[i]
        d[f"param{i}_before"] = params[i]
    d["predict"] = m.call("predict", {"x": x}); d["loss"] = m.call("loss", {"x": x, "y": y})
    for _ in range(2):
        m.apply("train", {"x": x, "y": y})
    for i, tid in enumerate(sorted(m.params)):
        d[f"param{i}_after"] = np.array(m.params[tid])
    out["dense_step"] = d
    # C4 at reduced size: forward, d_filters, d_images
    m = o.compile(*G.conv2_net(o, OL, ct="cpu"), seed=1)
    w = np.random.default_rng(1).uniform(-2, 2, (4, 3, 3, 3)).astype(np.float32)
    m.params[sorted(m.params)[0]][...] = w
    img = np.random.default_rng(0).uniform(0, 1, (2, 9, 8, 3)).astype(np.float32)
    out["conv2"] = {"img": img, "filters": w, **{t: m.call(t, {"img": img}) for t in ("conv", "loss", "dw", "dimg")}}
    # C1: xor trajectory (loss after every 10th of 100 steps) and the final parameters
    m = o.compile(*G.xor_net(o, OL, rate=0.1, ct="cpu"), seed=1)
    X = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float32); Y = np.array([[0], [1], [1], [0]], np.float32)
    d = {f"param{i}_before": np.array(m.params[tid]) for i, tid in enumerate(sorted(m.params))}
    losses = []
    for step in range(100):
        m.apply("train", {"x": X, "y": Y})
        if step % 10 == 9:
            losses.append(float(m.call("loss", {"x": X, "y": Y})[0]))
    d["losses"] = np.array(losses, np.float32)
    for i, tid in enumerate(sorted(m.params)):
        d[f"param{i}_after"] = np.array(m.params[tid])
    out["xor"] = d
    # next row f1/f2: conv + leakyRelu + maxpool + dense, adam (caches, epoch)
    m = o.compile(*G.fashion_net(o, OL, ct="cpu"), seed=1)
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (8, 12, 12, 1)).astype(np.float32)
    y = np.zeros((8, 10), np.float32); y[np.arange(8), rng.integers(0, 10, 8)] = 1
    d = {"x": x, "y": y}
    for i, tid in enumerate(sorted(m.params)):
        d[f"param{i}_before"] = np.array(m.params[tid])
    for _ in range(2):
        m.epoch += 1
        m.apply("train", {"x": x, "y": y})
    for i, tid in enumerate(sorted(m.params)):
        d[f"param{i}_after"] = np.array(m.params[tid])
    for i, tid in enumerate(sorted(m.caches)):
        d[f"cache{i}_after"] = np.array(m.caches[tid])
    out["fashion_adam"] = d
    return out


def shape_table(o, OL):
    """integer work: inferred shapes of every tensor for a few (graph, target, input shapes)"""
    from oracle.passes import compile_program, infer_shapes
    table = []
    for name, target, inputs in [("matmul", "c", {"a": [123, 100], "b": [100, 77]}),
                                 ("dense_net", "train", {"x": [8192, 784], "y": [8192, 10]}),
                                 ("conv2_net", "dimg", {"img": [256, 224, 224, 3]}),
                                 ("fashion_net", "train", {"x": [32, 12, 12, 1], "y": [32, 10]}),
                                 ("xor_net", "train", {"x": [4, 2], "y": [4, 1]})]:
        prog = o.ir.to_program(G.ALL[name](o, OL))
        compile_program(prog)
        shapes = infer_shapes(prog, target, {prog.inputs[k]: v for k, v in inputs.items()})
        table.append({"graph": name, "target": target, "inputs": inputs,
                      "shapes": {str(t): [int(v) for v in s] for t, s in sorted(shapes.items())}})
    return table


def main():
    import oracle as o
    from oracle import layers as OL
    for name, arrays in cases(o, OL).items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})
        print(name, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})
    with open(os.path.join(HERE, "shapes.json"), "w") as f:
        json.dump(shape_table(o, OL), f, indent=1)
    print("shapes.json")


if __name__ == "__main__":
    main()
