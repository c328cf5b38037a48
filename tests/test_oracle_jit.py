"""CPU tier: the LLVM-JIT form of the oracle's contraction kernel (oracle/jit.py, shaped like exprgrad/llvmgen.nim) against
the gcc-compiled restatement: same loop order, same un-contracted fp32 arithmetic, therefore the same bits."""
import numpy as np
import pytest

from parity_cases import oracle_matmul


@pytest.mark.parametrize("M,N,K,threads", [(1, 1, 1, 1), (2, 2, 3, 1), (64, 48, 100, 4), (123, 77, 129, 3)])
def test_jit_matmul_is_bit_identical_to_the_compiled_oracle(M, N, K, threads):
    jit = pytest.importorskip("oracle.jit")
    if not jit.available():
        pytest.skip("llvmlite is not importable")
    rng = np.random.default_rng(M + N + K)
    a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    b = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    assert np.array_equal(jit.matmul(a, b, threads=threads), oracle_matmul(a, b))


def test_jit_ir_has_the_reference_loop_shape():
    """cond / body / end / incr blocks per loop, icmp eq exit test, in-bounds GEPs, 4-byte aligned accesses, no
    fast-math flags (llvmgen.nim:277-301, 320-360; llvm.nim:486-491)."""
    jit = pytest.importorskip("oracle.jit")
    if not jit.available():
        pytest.skip("llvmlite is not importable")
    ir = jit.matmul_ir()
    for loop in ("y", "it", "x"):
        for part in ("cond", "body", "end", "incr"):
            assert f"{loop}_{part}" in ir
    assert ir.count("icmp eq") == 3 and ir.count("getelementptr inbounds") == 3
    assert "align 4" in ir and " fast " not in ir and "contract" not in ir
    # reference known answer (tests/test_model.nim:37-44)
    c = jit.matmul(np.array([[1, 2, 3], [4, 5, 6]], np.float32), np.array([[1, 2], [3, 4], [5, 6]], np.float32), threads=1)
    assert np.array_equal(c, np.array([[22, 28], [49, 64]], np.float32))
