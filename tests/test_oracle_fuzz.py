"""CPU-only: the oracle's reverse-mode autodiff (oracle/passes.py derive / generate, restating passes.nim:383-698)
checked against central finite differences of its own forward pass in float64 on the seeded random graphs of
tests/fuzz_graphs.py - an independent pin of the adjoints the device kernels are compared with (the reference
holds derivative identities for single operations only, test_model.nim:265-359)."""
import numpy as np
import pytest

import fuzz_graphs as FG

SEEDS = list(range(0, 48, 2))


@pytest.mark.parametrize("seed", SEEDS)
def test_gradient_matches_finite_differences(seed):
    import oracle as o
    from oracle import layers as OL
    graphs, what = FG.random_net(o, OL, seed, ct="cpu")
    om = o.compile(*graphs, scalar="float64", seed=0, openmp=False)
    inputs = {k: v.astype(np.float64) for k, v in FG.random_inputs(np, seed, rows=3).items() if k in om.program.inputs}
    names = {om.program.tdef(t).name: t for t in om.params}
    for k, v in FG.random_params(np, seed).items():
        if k in names:
            om.params[names[k]][...] = v
    da = np.array(om.call("da", inputs))
    a = inputs["a"]
    assert da.shape == a.shape
    h = 1e-6
    fd = np.zeros_like(a)
    for idx in np.ndindex(*a.shape):
        up = dict(inputs); dn = dict(inputs)
        up["a"] = a.copy(); up["a"][idx] += h
        dn["a"] = a.copy(); dn["a"][idx] -= h
        fd[idx] = (float(om.call("loss", up)[0]) - float(om.call("loss", dn)[0])) / (2 * h)
    scale = max(np.abs(fd).max(), np.abs(da).max(), 1e-12)
    err = np.abs(da - fd).max() / scale
    assert err < 1e-5, f"seed {seed} ({what}): |autodiff - finite differences| / max = {err:.2e}"
