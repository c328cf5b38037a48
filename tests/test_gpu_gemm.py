"""GPU parity of the tcgen05 contraction kernel (C-ABI egb_gemm_f32) against the oracle."""
import numpy as np
import pytest

from parity_cases import assert_close, gemm_f32, oracle_matmul

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import exprgrad_b200 as eg
    c = eg.new_gpu_context()
    yield c
    c.destroy()


def test_reference_known_answer(ctx):
    """tests/test_model.nim:37-44: [[1,2,3],[4,5,6]] x [[1,2],[3,4],[5,6]] == [[22,28],[49,64]] (exact)."""
    a = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    b = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    assert np.array_equal(gemm_f32(ctx, a, b), np.array([[22, 28], [49, 64]], np.float32))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (2, 2, 3), (64, 64, 64), (123, 77, 100), (128, 256, 64),
                                   (300, 200, 129), (1024, 512, 784), (1024, 10, 512), (784, 512, 1024)])
@pytest.mark.parametrize("lo,hi", [(0.0, 1.0), (-1.0, 1.0)])
def test_nn_matches_oracle(ctx, M, N, K, lo, hi):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.uniform(lo, hi, (M, K)).astype(np.float32)
    b = rng.uniform(lo, hi, (K, N)).astype(np.float32)
    assert_close(gemm_f32(ctx, a, b), oracle_matmul(a, b), what=f"NN {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(2, 2, 3), (123, 77, 100), (512, 10, 1024), (784, 512, 1024)])
def test_adjoint_layouts_match_oracle(ctx, M, N, K):
    """dA = dC.B^T (NT) and dB = A^T.dC (TN) layouts of passes.nim:519-549."""
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    b = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    ref = oracle_matmul(a, b)
    assert_close(gemm_f32(ctx, a, np.ascontiguousarray(b.T), tb=1), ref, what="NT")
    assert_close(gemm_f32(ctx, np.ascontiguousarray(a.T), b, ta=1), ref, what="TN")


def test_accumulate_bias_relu_epilogue(ctx):
    rng = np.random.default_rng(9)
    a = rng.uniform(-1, 1, (300, 100)).astype(np.float32)
    b = rng.uniform(-1, 1, (100, 200)).astype(np.float32)
    c0 = rng.uniform(-1, 1, (300, 200)).astype(np.float32)
    bias = rng.uniform(-1, 1, (200,)).astype(np.float32)
    ref = oracle_matmul(a, b)
    assert_close(gemm_f32(ctx, a, b, flags=1, c0=c0), ref + c0, what="accumulate")
    assert_close(gemm_f32(ctx, a, b, flags=2, bias=bias), ref + bias, what="bias")
    assert_close(gemm_f32(ctx, a, b, flags=6, bias=bias), np.maximum(ref + bias, 0), what="bias+relu")


def test_full_size_checksum_properties(ctx):
    """benchmarks/matmul at 4096^3 (BASELINE config 2): the oracle would take minutes, so check
    size-independent properties in fp64: row sums (C.1 == A.(B.1)), column sums, and linearity."""
    n = 4096
    for seed, lo in ((0, 0.0), (1, -1.0)):
        rng = np.random.default_rng(seed)
        a = rng.uniform(lo, 1, (n, n)).astype(np.float32)
        b = rng.uniform(lo, 1, (n, n)).astype(np.float32)
        c = gemm_f32(ctx, a, b).astype(np.float64)
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        rows = a64 @ b64.sum(1)
        cols = a64.sum(0) @ b64
        # same 1e-4 normalised bar as element-wise parity (the tensor cores' truncating fp32
        # accumulation leaves a ~2e-5 low bias on all-positive data at K=4096)
        assert np.abs(c.sum(1) - rows).max() / np.abs(rows).max() < 1e-4
        assert np.abs(c.sum(0) - cols).max() / np.abs(cols).max() < 1e-4
        # a 64x64 corner against the oracle itself
        ref = oracle_matmul(a[:64], b[:, :64])
        assert_close(c[:64, :64], ref, what="corner")
        if seed == 0:
            c2 = gemm_f32(ctx, a * np.float32(2), b).astype(np.float64)  # exact scaling by 2
            assert np.array_equal(c2, 2 * c)


@pytest.mark.parametrize("seed,lo", [(0, 0.0), (1, -1.0)])
def test_full_size_matches_oracle_elementwise(ctx, seed, lo):
    """BASELINE config 2 at full size, element by element: the whole 4096^3 product through the public model
    API (benchmarks/matmul/matmul_gpu.nim:28-36, inputs as :69-70 plus the mixed-sign variant) against the
    oracle's row-split loop nest on all host cores (about a second per product)."""
    import oracle as o
    from oracle import layers as OL
    import exprgrad_b200 as eg
    from exprgrad_b200 import frontend as F, layers as PL
    import graphs as G
    n = 4096
    rng = np.random.default_rng(seed)
    a = rng.uniform(lo, 1, (n, n)).astype(np.float32)
    b = rng.uniform(lo, 1, (n, n)).astype(np.float32)
    ref = o.compile(*G.matmul(o, OL, ct="threads")).call("c", {"a": a, "b": b})
    pm = eg.compile(*G.matmul(F, PL), gpu=ctx)
    got = pm.call("c", {"a": a, "b": b})
    e = assert_close(got, ref, what=f"4096^3 U({lo},1) full product")
    print(f"4096^3 U({lo},1): normalised max error vs oracle {e:.2e}")
    # the streamed host-buffer form of the same call (what bench.py's e2e leg times)
    out = np.empty((n, n), np.float32)
    pm.call("c", {"a": a, "b": b}, out=out)
    assert_close(out, ref, what="4096^3 streamed call")
    pm.free()


def _bf16_planes(x):
    """hi = bf16(x), mid = bf16(x - hi) with round-to-nearest-even (what split.cu computes), as uint16."""
    def rn(v):
        bits = v.astype(np.float32).view(np.uint32).astype(np.uint64)
        return (((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)
    hi = rn(x)
    hi_f = (hi.astype(np.uint32) << 16).view(np.float32)
    return hi, rn(x - hi_f)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(1024, 512, 784, 0, 0), (1024, 512, 784, 0, 1), (512, 10, 1024, 1, 1),
                                             (300, 200, 520, 0, 0), (784, 512, 1024, 1, 1), (128, 64, 128, 0, 0)])
@pytest.mark.parametrize("bn,ck", [(64, 2), (128, 4), (256, 8), (64, 8)])
def test_cluster_split_k_configurations(ctx, M, N, K, a_mn, b_mn, bn, ck):
    _planes_case(ctx, M, N, K, a_mn, b_mn, bn, ck, 0)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(1024, 512, 784, 0, 0), (1024, 512, 512, 0, 1), (300, 200, 520, 0, 0),
                                             (784, 512, 1024, 1, 1), (128, 64, 128, 0, 0), (37, 48, 64, 0, 0),
                                             (129, 36, 200, 1, 0)])
@pytest.mark.parametrize("bn,ck", [(32, 1), (64, 1), (128, 1), (256, 1), (32, 2), (64, 2), (32, 4), (64, 4)])
@pytest.mark.parametrize("flags", [0, 1 | 2, 2 | 4])
def test_latency_kernel_configurations(ctx, M, N, K, a_mn, b_mn, bn, ck, flags):
    """gemm_lat.cu (TMA-staged epilogue in the TMEM-native layout, push-based cluster split-K): every tile width and
    split factor it accepts, ragged edges (TMA clips), `+=` on the old contents, bias, relu in place."""
    if b_mn and bn % 64:
        pytest.skip("an MN-major B tile is made of 64-column groups")
    _planes_case(ctx, M, N, K, a_mn, b_mn, bn, ck, flags)


def _planes_case(ctx, M, N, K, a_mn, b_mn, bn, ck, xflags):
    """Forced tile width x cluster split-K factor (egb_gemm_planes): each CTA of a cluster reduces a k
    range of one tile and the partial tiles meet through distributed shared memory."""
    import ctypes
    from exprgrad_b200._ffi import check, lib
    rng = np.random.default_rng(M + N + K + bn + ck)
    a = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    b = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    pad8 = lambda v: (v + 7) // 8 * 8
    a_st = np.ascontiguousarray(a.T) if a_mn else a          # stored [K, M] (MN-major) or [M, K]
    b_st = b if b_mn else np.ascontiguousarray(b.T)          # stored [K, N] (MN-major) or [N, K]
    bufs = []
    def upload(mat):
        ld = pad8(mat.shape[1])
        padded = np.zeros((mat.shape[0], ld), np.float32); padded[:, :mat.shape[1]] = mat
        out = []
        for plane in _bf16_planes(padded):
            buf = ctx.alloc_buffer(plane.nbytes); buf.write(plane); bufs.append(buf); out.append(buf.device_ptr)
        return out, ld
    (a_hi, a_mid), lda = upload(a_st)
    (b_hi, b_mid), ldb = upload(b_st)
    c = ctx.alloc_buffer(M * N * 4); bufs.append(c)
    c0 = rng.uniform(-1, 1, (M, N)).astype(np.float32) if xflags & 1 else np.zeros((M, N), np.float32)
    c.write(c0)
    bias = rng.uniform(-1, 1, (N,)).astype(np.float32)
    dbias = ctx.alloc_buffer(N * 4); bufs.append(dbias); dbias.write(bias)
    flags = (16 if a_mn else 0) | (32 if b_mn else 0) | xflags
    check(lib.egb_gemm_planes(ctx.handle, M, N, K, a_hi, a_mid, lda, b_hi, b_mid, ldb, c.device_ptr, N, flags,
                              dbias.device_ptr if xflags & 2 else None, ctypes.c_float(1.0), bn | (ck << 16)))
    got = c.read().reshape(M, N)
    ref = a.astype(np.float64) @ b.astype(np.float64) + c0
    if xflags & 2:
        ref = ref + bias
    if xflags & 4:
        ref = np.maximum(ref, 0)
    assert_close(got, ref, what=f"bn{bn} ck{ck} flags{xflags}")
    for x in bufs:
        x.dealloc()
