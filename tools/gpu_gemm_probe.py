"""Development probe (not a parity test): runs the tcgen05 GEMM on a B200 and compares with fp64."""
import ctypes, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.path.join(ROOT, "exprgrad_b200", "libegb200.so"))
L.egb_last_error.restype = ctypes.c_char_p
L.egb_buffer_device_ptr.restype = ctypes.c_void_p
L.egb_context_stream.restype = ctypes.c_void_p
P = ctypes.c_void_p
i64 = ctypes.c_int64
def ck(st):
    if st != 0:
        raise RuntimeError(L.egb_last_error().decode())
ctx = P()
ck(L.egb_context_create(0, ctypes.byref(ctx)))
def dev(arr=None, nbytes=None):
    b = P()
    n = arr.nbytes if arr is not None else nbytes
    ck(L.egb_alloc_buffer(ctx, ctypes.c_size_t(n), ctypes.byref(b)))
    if arr is not None:
        ck(L.egb_buffer_write(b, arr.ctypes.data_as(P), ctypes.c_size_t(n)))
    return b
def ptr(b): return P(L.egb_buffer_device_ptr(b))
def read(b, shape, dt=np.float32):
    out = np.empty(shape, dt)
    ck(L.egb_buffer_read_into(b, out.ctypes.data_as(P), ctypes.c_size_t(out.nbytes)))
    return out
import torch
def ref64(a, b):
    return (torch.from_numpy(a).cuda().double() @ torch.from_numpy(b).cuda().double()).cpu().numpy()

def run(M, N, K, ta=0, tb=0, flags=0, seed=0, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    A = rng.uniform(lo, hi, (K, M) if ta else (M, K)).astype(np.float32)
    B = rng.uniform(lo, hi, (N, K) if tb else (K, N)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (M, N)).astype(np.float32)
    bias = rng.uniform(-1, 1, (N,)).astype(np.float32)
    dA, dB, dC, dbias = dev(A), dev(B), dev(C0), dev(bias)
    ck(L.egb_gemm_f32(ctx, ta, tb, i64(M), i64(N), i64(K), ptr(dA), i64(A.shape[1]), ptr(dB), i64(B.shape[1]),
                      ptr(dC), i64(N), flags, ptr(dbias), ctypes.c_float(1.0)))
    ck(L.egb_context_synchronize(ctx))
    C = read(dC, (M, N))
    R = ref64(A.T if ta else A, B.T if tb else B)
    if flags & 2: R = R + bias[None, :]
    if flags & 1: R = R + C0
    if flags & 4: R = np.maximum(R, 0)
    err = np.abs(C - R).max() / max(np.abs(R).max(), 1e-30)
    print(f"M={M} N={N} K={K} ta={ta} tb={tb} flags={flags} range=({lo},{hi}) normalised max err = {err:.3e}", flush=True)
    for b in (dA, dB, dC, dbias): L.egb_buffer_free(b)
    return err

bad = 0
for (M, N, K) in [(128, 32, 64), (128, 256, 64), (128, 256, 128), (256, 512, 256), (123, 77, 100), (2, 2, 3),
                  (1024, 512, 784), (1024, 10, 512), (784, 512, 1024), (512, 512, 1024)]:
    for (ta, tb) in [(0, 0), (0, 1), (1, 0)]:
        e = run(M, N, K, ta, tb, flags=0, lo=-1.0, hi=1.0)
        bad += e > 1e-4
e = run(1024, 512, 784, flags=2 | 4, lo=-1, hi=1); bad += e > 1e-4
e = run(300, 200, 100, flags=1, lo=-1, hi=1); bad += e > 1e-4
e = run(4096, 4096, 4096, lo=0, hi=1); bad += e > 1e-4
e = run(4096, 4096, 4096, lo=-1, hi=1, seed=1); bad += e > 1e-4
print("BAD =", bad, flush=True)

# timing 4096^3
M = N = K = 4096
rng = np.random.default_rng(0)
A = rng.uniform(0, 1, (M, K)).astype(np.float32); B = rng.uniform(0, 1, (K, N)).astype(np.float32)
dA, dB, dC = dev(A), dev(B), dev(nbytes=M * N * 4)
st = torch.cuda.ExternalStream(L.egb_context_stream(ctx))
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    ck(L.egb_context_synchronize(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): fn()
    e1.record(st)
    e1.synchronize()
    return e0.elapsed_time(e1) / iters
t = timeit(lambda: ck(L.egb_gemm_f32(ctx, 0, 0, i64(M), i64(N), i64(K), ptr(dA), i64(K), ptr(dB), i64(N), ptr(dC), i64(N), 0, None, ctypes.c_float(1.0))))
print(f"gemm_f32 4096^3 (split+gemm): {t:.3f} ms  {2*M*N*K/t/1e9:.1f} TFLOP/s", flush=True)
pl = [dev(nbytes=M * K * 2) for _ in range(4)]
ck(L.egb_split_bf16(ctx, ptr(dA), i64(M), i64(K), i64(K), 0, ptr(pl[0]), ptr(pl[1]), i64(K), 0))
ck(L.egb_split_bf16(ctx, ptr(dB), i64(K), i64(N), i64(N), 1, ptr(pl[2]), ptr(pl[3]), i64(K), 0))
t = timeit(lambda: ck(L.egb_split_bf16(ctx, ptr(dA), i64(M), i64(K), i64(K), 0, ptr(pl[0]), ptr(pl[1]), i64(K), 0)))
print(f"split rows 4096^2: {t*1000:.1f} us  {(M*K*8)/t/1e6:.0f} GB/s", flush=True)
t = timeit(lambda: ck(L.egb_split_bf16(ctx, ptr(dB), i64(K), i64(N), i64(N), 1, ptr(pl[2]), ptr(pl[3]), i64(K), 0)))
print(f"split transpose 4096^2: {t*1000:.1f} us  {(M*K*8)/t/1e6:.0f} GB/s", flush=True)
for bn in (256, 224, 192, 128, 64):
    t = timeit(lambda: ck(L.egb_gemm_planes(ctx, i64(M), i64(N), i64(K), ptr(pl[0]), ptr(pl[1]), i64(K), ptr(pl[2]), ptr(pl[3]), i64(K), ptr(dC), i64(N), 0, None, ctypes.c_float(1.0), bn)))
    print(f"gemm_planes BN={bn}: {t:.3f} ms  {2*M*N*K/t/1e9:.1f} TFLOP/s fp32-equivalent, {6*M*N*K/t/1e9:.1f} TFLOP/s bf16 tensor", flush=True)
C = read(dC, (M, N))
print("checksum", float(C.sum()))
