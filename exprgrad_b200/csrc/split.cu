// fp32 -> bf16 (hi, mid) operand planes for the bf16x3 tensor-core contraction, optionally
// transposing so that the reduction dimension becomes contiguous (K-major) and optionally applying
// the reference relu (exprgrad/layers/dnn.nim:26-27) on the fly. HBM-bound streaming kernels:
// 128-bit loads, 64/128-bit stores, grid-stride over a grid sized to the SM count.
#include "egb_internal.hpp"

namespace egb {
namespace {

__device__ __forceinline__ float apply_act(float x, int act) {
  return act == 1 ? ((0.0f <= x) ? x : 0.0f) : x;
}

__device__ __forceinline__ void split2(float x, __nv_bfloat16& h, __nv_bfloat16& m) {
  h = __float2bfloat16_rn(x);
  m = __float2bfloat16_rn(x - __bfloat162float(h));
}

// rows x cols, no transpose. One thread handles 8 consecutive columns (two float4 loads,
// one 16-byte store per plane) when the row is 8-aligned, otherwise falls back to scalars.
__global__ void split_rows_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid, int dst_ld,
                                  int act) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = (cols + 7) >> 3;
  const long total = (long)rows * chunks;
  const bool vec = ((ld & 3) == 0) && ((dst_ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / chunks);
    const int c0 = (int)(i % chunks) << 3;
    const float* s = src + (size_t)r * ld + c0;
    __nv_bfloat16* h = hi + (size_t)r * dst_ld + c0;
    __nv_bfloat16* m = mid + (size_t)r * dst_ld + c0;
    if (vec && c0 + 8 <= cols) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(s));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(s) + 1);
      const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      __align__(16) __nv_bfloat16 hv[8];
      __align__(16) __nv_bfloat16 mv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2(apply_act(xs[j], act), hv[j], mv[j]);
      *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(hv);
      *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(mv);
    } else {
      for (int j = 0; j < 8 && c0 + j < cols; ++j) split2(apply_act(s[j], act), h[j], m[j]);
    }
  }
}

// Transposing variant: src [rows, cols] -> planes [cols, rows]. 64x64 tiles through shared memory so
// both the fp32 reads (along cols) and the bf16 writes (along rows) are coalesced.
__global__ void split_transpose_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                                       int dst_ld, int act) {
  __shared__ float tile[64][65];
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_c = (cols + 63) >> 6;
  const int tiles_r = (rows + 63) >> 6;
  const int tx = threadIdx.x & 63;  // 256 threads: 64 x 4
  const int ty = threadIdx.x >> 6;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) << 6;
    const int c0 = (t % tiles_c) << 6;
    __syncthreads();
#pragma unroll 4
    for (int j = ty; j < 64; j += 4) {
      const int r = r0 + j, c = c0 + tx;
      tile[j][tx] = (r < rows && c < cols) ? __ldg(src + (size_t)r * ld + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = ty; j < 64; j += 4) {
      const int c = c0 + j, r = r0 + tx;  // output row = source column
      if (c < cols && r < rows) {
        __nv_bfloat16 h, m;
        split2(apply_act(tile[tx][j], act), h, m);
        hi[(size_t)c * dst_ld + r] = h;
        mid[(size_t)c * dst_ld + r] = m;
      }
    }
  }
}

}  // namespace

void launch_split_bf16(Context& ctx, const float* src, int rows, int cols, int ld, bool transpose,
                       __nv_bfloat16* hi, __nv_bfloat16* mid, int dst_ld, int act, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  if (!transpose) {
    const long total = (long)rows * ((cols + 7) >> 3);
    long blocks = (total + 255) / 256;
    const long cap = (long)ctx.sm_count * 8;
    if (blocks > cap) blocks = cap;
    Launch l(ctx, KC_SPLIT, st);
    launch_kernel(ctx, split_rows_kernel, dim3((int)blocks), dim3(256), 0, st, src, rows, cols, ld, hi, mid, dst_ld, act);
  } else {
    const long tiles = (long)((rows + 63) >> 6) * ((cols + 63) >> 6);
    long blocks = tiles;
    const long cap = (long)ctx.sm_count * 8;
    if (blocks > cap) blocks = cap;
    Launch l(ctx, KC_SPLIT, st);
    launch_kernel(ctx, split_transpose_kernel, dim3((int)blocks), dim3(256), 0, st, src, rows, cols, ld, hi, mid, dst_ld,
                  act);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
