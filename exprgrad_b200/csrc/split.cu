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

// Several row splits in one launch (the operand planes every step needs before its first contraction:
// the input batch and the parameters). Work items are 8-column chunks, numbered job after job.
__global__ void split_rows_batch_kernel(const SplitBatch b) {
  pdl_launch_dependents();
  pdl_wait();
  // the plan's zero-initialised results (a few KB; the gradient bucket under data parallelism) are cleared
  // here instead of by a memset node in front of this kernel on the critical path
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.zero_vec16; i += (size_t)gridDim.x * blockDim.x)
    b.zero_ptr[i] = make_uint4(0, 0, 0, 0);
  const long total = b.first_chunk[b.n];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int j = 0;
#pragma unroll
    for (int q = 1; q < SplitBatch::MAX_JOBS; ++q)
      if (q < b.n && i >= b.first_chunk[q]) j = q;
    const SplitJob& job = b.jobs[j];
    const long li = i - b.first_chunk[j];
    const int chunks = (job.cols + 7) >> 3;
    const int r = (int)(li / chunks);
    const int c0 = (int)(li % chunks) << 3;
    const bool vec = ((job.ld & 3) == 0) && ((job.dst_ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(job.src) & 15) == 0);
    const float* s = job.src + (size_t)r * job.ld + c0;
    __nv_bfloat16* h = job.hi + (size_t)r * job.dst_ld + c0;
    __nv_bfloat16* m = job.mid + (size_t)r * job.dst_ld + c0;
    if (vec && c0 + 8 <= job.cols) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(s));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(s) + 1);
      const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      __align__(16) __nv_bfloat16 hv[8];
      __align__(16) __nv_bfloat16 mv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) split2(apply_act(xs[q], job.act), hv[q], mv[q]);
      *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(hv);
      *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(mv);
    } else {
      for (int q = 0; q < 8 && c0 + q < job.cols; ++q) split2(apply_act(s[q], job.act), h[q], m[q]);
    }
  }
}

// Transposing variant: src [rows, cols] -> planes [cols, rows]. 64x64 tiles through shared memory:
// 128-bit loads along the source rows, then every thread converts 8 consecutive source rows of one
// source column and writes them as one 16-byte store per plane (the output row is contiguous in r).
__global__ void split_transpose_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                                       int dst_ld, int act) {
  __shared__ float tile[64][65];
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_c = (cols + 63) >> 6;
  const int tiles_r = (rows + 63) >> 6;
  const bool vec_in = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const bool vec_out = ((dst_ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(hi) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(mid) & 15) == 0);
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) << 6;
    const int c0 = (t % tiles_c) << 6;
    __syncthreads();
    // load: 64 rows x 16 float4 = 1024 vector loads, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int rr = idx >> 4, c4 = (idx & 15) << 2;
      const int r = r0 + rr, c = c0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        const float* p = src + (size_t)r * ld + c;
        if (vec_in && c + 4 <= cols) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (c + 0 < cols) v.x = __ldg(p + 0);
          if (c + 1 < cols) v.y = __ldg(p + 1);
          if (c + 2 < cols) v.z = __ldg(p + 2);
          if (c + 3 < cols) v.w = __ldg(p + 3);
        }
      }
      tile[rr][c4 + 0] = v.x; tile[rr][c4 + 1] = v.y; tile[rr][c4 + 2] = v.z; tile[rr][c4 + 3] = v.w;
    }
    __syncthreads();
    // store: 64 output rows (source columns) x 8 groups of 8 source rows = 512 items, 2 per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int g = idx & 7, cc = idx >> 3;    // 8 neighbouring lanes write one contiguous 128-byte output row segment
      const int c = c0 + cc, r = r0 + g * 8;
      if (c >= cols || r >= rows) continue;
      __align__(16) __nv_bfloat16 hv[8];
      __align__(16) __nv_bfloat16 mv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split2(apply_act(tile[g * 8 + e][cc], act), hv[e], mv[e]);
      __nv_bfloat16* h = hi + (size_t)c * dst_ld + r;
      __nv_bfloat16* m = mid + (size_t)c * dst_ld + r;
      if (vec_out && r + 8 <= rows) {
        *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(hv);
        *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(mv);
      } else {
        for (int e = 0; e < 8 && r + e < rows; ++e) {
          h[e] = hv[e];
          m[e] = mv[e];
        }
      }
    }
  }
}

}  // namespace

void launch_split_batch(Context& ctx, const SplitJob* jobs, int n, cudaStream_t st, void* zero_ptr, size_t zero_bytes) {
  if (n <= 0) return;
  if ((reinterpret_cast<uintptr_t>(zero_ptr) & 15) || (zero_bytes & 15)) fail(EGB_ERR_GPU, "split batch: unaligned zero region");
  if (n > SplitBatch::MAX_JOBS) fail(EGB_ERR_GPU, "split batch of %d jobs exceeds %d", n, SplitBatch::MAX_JOBS);
  SplitBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  long total = 0;
  for (int j = 0; j < n; ++j) {
    b.jobs[j] = jobs[j];
    b.first_chunk[j] = total;
    total += (long)jobs[j].rows * ((jobs[j].cols + 7) >> 3);
  }
  b.first_chunk[n] = total;
  b.zero_ptr = (uint4*)zero_ptr;
  b.zero_vec16 = zero_bytes / 16;
  if (total == 0 && zero_bytes == 0) return;
  long blocks = (total + 255) / 256;
  const long cap = (long)ctx.sm_count * 8;
  if (blocks > cap) blocks = cap;
  Launch l(ctx, KC_SPLIT, st);
  launch_kernel(ctx, split_rows_batch_kernel, dim3((int)blocks), dim3(256), 0, st, b);
  EGB_CUDA(cudaGetLastError());
}

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
