"""Times the contraction kernel alone at the dense-net shapes (planes resident, back-to-back launches)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import exprgrad_b200 as eg
from exprgrad_b200 import gpu as G
from exprgrad_b200._ffi import check, lib
ctx = eg.new_gpu_context()
P = ctypes.c_void_p
def buf(nbytes):
    b = ctx.alloc_buffer(nbytes); return b
def run(M, N, K, a_mn, b_mn, bn, iters=300):
    a = buf(max(M, K) * ((max(M, K) + 7) // 8 * 8) * 2 * 2 + 4096); b = buf(max(N, K) * ((max(N, K) + 7) // 8 * 8) * 2 * 2 + 4096)
    for x in (a, b): x.fill(0.0)
    c = buf(M * N * 4)
    lda = ((M if a_mn else K) + 7) // 8 * 8
    ldb = ((N if b_mn else K) + 7) // 8 * 8
    flags = (16 if a_mn else 0) | (32 if b_mn else 0)
    half_a = (K if a_mn else M) * lda * 2
    half_b = (K if b_mn else N) * ldb * 2
    half_a = (half_a + 255) // 256 * 256; half_b = (half_b + 255) // 256 * 256
    def go():
        check(lib.egb_gemm_planes(ctx.handle, M, N, K, a.device_ptr, a.device_ptr + half_a, lda, b.device_ptr, b.device_ptr + half_b, ldb,
                                  c.device_ptr, N, flags, None, ctypes.c_float(1.0), bn))
    for _ in range(10): go()
    ctx.synchronize()
    e0, e1 = G.GpuEvent(ctx), G.GpuEvent(ctx); e0.record()
    for _ in range(iters): go()
    e1.record(); us = e0.elapsed_ms(e1) / iters * 1e3
    for x in (a, b, c): x.dealloc()
    return us
for (M, N, K, a_mn, b_mn, name) in [(1024, 512, 784, 0, 1, "fwd1 h=x.W1 (B MN-major)"), (1024, 512, 784, 0, 0, "fwd1 with K-major B"),
                                     (1024, 512, 512, 0, 0, "da1 = dh.W^T (K-major B)"), (512, 512, 1024, 1, 1, "dW2 = a^T.dh (both MN-major)"),
                                     (784, 512, 1024, 1, 1, "dW1"), (1024, 10, 512, 0, 1, "fwd3 N=10"), (512, 10, 1024, 1, 1, "dW3 N=10"), (1024, 512, 10, 0, 0, "da2 K=10")]:
    res = []
    res.append(f"auto:{run(M, N, K, a_mn, b_mn, 0):.1f}")
    for bn in ([32, 64, 128, 256] if not b_mn else [64, 128, 256]):
        if bn > max(64, N) * 2: continue
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
        for ck in (1, 2, 4, 8):
            if ck > 1 and (tiles * ck > 148 or ck * 2 > (K + 63) // 64): continue
            res.append(f"bn{bn}x{ck}:{run(M, N, K, a_mn, b_mn, bn | (ck << 16)):.1f}")
    print(f"{name:34s} M={M} N={N} K={K}  " + "  ".join(res), flush=True)
