// Hand-fused row kernel for the reference's classification head: softmax (exprgrad/layers/dnn.nim:90-94,
// no max-subtraction) followed by crossEntropy (exprgrad/layers/base.nim:66-67) and the adjoint kernels
// `derive` generates for the pair (exprgrad/passes.nim:383-549):
//
//   s[y]     = sum_x exp(h[y,x])                               p[y,x]  = exp(h[y,x]) / s[y]
//   dp[y,x]  = ((0 - dL/N) * labels[y,x]) / p[y,x]             dh[y,x] = (dp[y,x] / s[y]) * exp(h[y,x])
//   ds[y]    = sum_x (0 - exp(h[y,x])) * (dp[y,x] / (s[y]*s[y]))   dh[y,x] += ds[y] * exp(h[y,x])
//
// The planner substitutes it for the interpreted row chain when the six kernels appear with exactly this
// wiring; every intermediate tensor is still written, so nothing downstream can tell the difference.
// One warp per row, exp evaluated once per element, both row sums by warp shuffles.
#include "egb_internal.hpp"

namespace egb {
namespace {

constexpr int MAX_PER_LANE = 4;  // rows of up to 128 columns

__global__ void __launch_bounds__(256) softmax_xent_rows_kernel(const float* __restrict__ H, const float* __restrict__ Y,
                                                                const float* __restrict__ DL, float* __restrict__ S,
                                                                float* __restrict__ P, float* __restrict__ DP,
                                                                float* __restrict__ DH, float* __restrict__ DS, int rows,
                                                                int cols, float* __restrict__ colsum,
                                                                __nv_bfloat16* __restrict__ out_hi,
                                                                __nv_bfloat16* __restrict__ out_mid, int ld_out) {
  __shared__ float colpart[8][32 * MAX_PER_LANE];
  pdl_launch_dependents();
  pdl_wait();
  float colacc[MAX_PER_LANE];
#pragma unroll
  for (int i = 0; i < MAX_PER_LANE; ++i) colacc[i] = 0.0f;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float scale = 0.0f - __fdiv_rn(DL[0], (float)rows);   // 0 - dL / toScalar(shape[0])
  for (int r = warp; r < rows; r += nwarps) {
    float e[MAX_PER_LANE], t[MAX_PER_LANE];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < MAX_PER_LANE; ++i) {
      const int x = lane + 32 * i;
      e[i] = x < cols ? expf(H[(size_t)r * cols + x]) : 0.0f;
      sum += e[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float s = sum;
    float dsum = 0.0f;
#pragma unroll
    for (int i = 0; i < MAX_PER_LANE; ++i) {
      const int x = lane + 32 * i;
      t[i] = 0.0f;
      if (x < cols) {
        const size_t idx = (size_t)r * cols + x;
        const float p = __fdiv_rn(e[i], s);
        const float dp = __fdiv_rn(__fmul_rn(scale, Y[idx]), p);
        P[idx] = p;
        DP[idx] = dp;
        t[i] = __fmul_rn(__fdiv_rn(dp, s), e[i]);
        dsum += __fmul_rn(0.0f - e[i], __fdiv_rn(dp, __fmul_rn(s, s)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) {
      S[r] = s;
      DS[r] = dsum;
    }
#pragma unroll
    for (int i = 0; i < MAX_PER_LANE; ++i) {
      const int x = lane + 32 * i;
      if (x < cols) {
        const float dh = __fadd_rn(t[i], __fmul_rn(dsum, e[i]));
        DH[(size_t)r * cols + x] = dh;
        colacc[i] += dh;
        if (out_hi) {  // bf16 operand planes of dh for the adjoint contractions (split.cu's arithmetic)
          const __nv_bfloat16 hi = __float2bfloat16_rn(dh);
          out_hi[(size_t)r * ld_out + x] = hi;
          out_mid[(size_t)r * ld_out + x] = __float2bfloat16_rn(dh - __bfloat162float(hi));
        }
      }
    }
  }
  if (colsum) {
    // bias gradient db[x] += sum_y dh[y,x] (the reference's column-sum kernel behind the head, dnn.nim:22-24
    // adjoint): per-warp partial sums meet in shared memory, one atomic per column and block
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MAX_PER_LANE; ++i) colpart[w][lane + 32 * i] = colacc[i];
    __syncthreads();
    for (int x = threadIdx.x; x < cols; x += blockDim.x) {
      float sum = 0.0f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) sum += colpart[ww][x];
      atomicAdd(colsum + x, sum);
    }
  }
}

}  // namespace

bool softmax_xent_supported(int64_t cols) { return cols >= 1 && cols <= 32 * MAX_PER_LANE; }

void launch_softmax_xent_rows(Context& ctx, const float* H, const float* Y, const float* DL, float* S, float* P, float* DP,
                              float* DH, float* DS, int rows, int cols, float* colsum, __nv_bfloat16* out_hi,
                              __nv_bfloat16* out_mid, int ld_out, cudaStream_t st) {
  if (rows <= 0) return;
  const int blocks = (rows + 7) / 8;
  const int cap = ctx.sm_count * 8;
  {
    Launch l(ctx, KC_REDUCE, st);
    launch_kernel(ctx, softmax_xent_rows_kernel, dim3(blocks < cap ? blocks : cap), dim3(256), 0, st, H, Y, DL, S, P, DP, DH, DS,
                  rows, cols, colsum, out_hi, out_mid, ld_out);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
