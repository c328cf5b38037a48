// The classification head of a dense net as ONE row kernel: the skinny contraction in front of the reference's
// softmax + crossEntropy pair, the pair itself with its derive()d adjoints, and the skinny adjoint contraction behind it
//
//   z[y, j]   = sum_k a[y, k] * w[k, j] + bias[j]                      (dense, dnn.nim:19-24; j < 32 classes)
//   s, p, dp, dz, ds                                                   (softmax + crossEntropy: see fused_rows.cu)
//   g[y, c]   = sum_j dz[y, j] * w'[j, c]  -> bias / relu mask / ...   (derive() of the dense layer, passes.nim:519-549)
//
// In the dense train step these were three launches on the critical path - a [1024 x 10 x 512] contraction, the row
// kernel, a [1024 x 512 x 10] contraction - of ~6-10 us each although they hold 20 MFLOP together: a tensor-core tile
// with 10 of its 64 columns (or 10 of its 64 k-steps) in use, launch gaps, two epilogues. Here one warp owns a batch
// row end to end: both weight operands sit in shared memory as fp32 (hi + mid of the bf16 planes the contractions
// would have read - the same values the tensor cores see), the forward dot products are reduced by shuffles, the
// softmax / crossEntropy chain runs one class per lane exactly as in fused_rows.cu, and the adjoint contraction is
// 8 output columns per lane with the usual fused stages (bias, relu / leakyRelu, their adjoint masks, operand planes,
// column sums). Every tensor the three launches wrote is still written (unless the planner marked it dead).
#include "egb_internal.hpp"
#include "ptx.cuh"
#include "runtime.hpp"

namespace egb {
namespace {

constexpr int WPR = 4;                        // warps per batch row
constexpr int ROWS = 8;                       // rows a CTA works on at a time
constexpr int HEAD_THREADS = WPR * ROWS * 32;
constexpr int CP = 16;                        // classes, padded: the forward table is [k][CP]
constexpr int COL_TILE = 128 * WPR;           // the adjoint table is padded to whole tiles of 4 columns per lane and warp
constexpr int MAX_CHUNKS = 2;                 // up to 2 x COL_TILE output columns

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void row_group_sync(int row) {   // the WPR warps of one row (named barriers 1 ..)
  asm volatile("bar.sync %0, %1;" ::"r"(row + 1), "r"(WPR * 32) : "memory");
}

// fp32 tables of both weight operands, built once per step from the bf16 planes the absorbed contractions would
// have read (hi + mid: the values the tensor cores see):  fwd[k][CP] (classes >= cols are zero) and
// bwd[j][nout_p] (columns >= Nout are zero). Element (k, j) of the forward operand: win_mn ? [k][ld] : [j][ld];
// element (c, j) of the adjoint operand (n = c, k = j): wout_mn ? [j][ld] : [c][ld].
__global__ void __launch_bounds__(256) head_tables_kernel(const HeadParams p, int nout_p) {
  pdl_launch_dependents();
  pdl_wait();
  const int nf = p.Kin * CP, nb = p.cols * nout_p;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nf + nb; idx += gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (idx < nf) {
      const int k = idx / CP, j = idx % CP;
      if (j < p.cols) {
        const size_t off = p.win_mn ? (size_t)k * p.win_ld + j : (size_t)j * p.win_ld + k;
        v = __bfloat162float(p.win_hi[off]) + __bfloat162float(p.win_mid[off]);
      }
    } else {
      const int j = (idx - nf) / nout_p, c = (idx - nf) % nout_p;
      if (c < p.Nout) {
        const size_t off = p.wout_mn ? (size_t)j * p.wout_ld + c : (size_t)c * p.wout_ld + j;
        v = __bfloat162float(p.wout_hi[off]) + __bfloat162float(p.wout_mid[off]);
      }
    }
    p.tables[idx] = v;
  }
}

// One CTA = ROWS batch rows at a time, WPR warps per row. The kernel is bound by instruction issue and dependent-
// instruction latency, not by memory (ncu on the one-warp-per-row version: 3 450 instructions per warp at one issue
// per 6.5 cycles, 2 warps per scheduler; 42 % of them built the weight tables in every CTA): the tables now arrive by
// one bulk copy, and a row's dot products are split over four warps so that every scheduler has 8 warps to pick from.
template <int kEpi>
__global__ void __launch_bounds__(HEAD_THREADS, 1) head_rows_kernel(const HeadParams p) {
  extern __shared__ __align__(128) float sm[];
  const int nout_p = (p.Nout + COL_TILE - 1) / COL_TILE * COL_TILE;
  float* const w_fwd = sm;                                   // [Kin][CP]
  float* const w_bwd = w_fwd + p.Kin * CP;                   // [cols][nout_p]
  float* const arow = w_bwd + p.cols * nout_p;               // [ROWS][Kin]  fp32 operand rows
  float* const zpart = arow + ROWS * p.Kin;                  // [ROWS][WPR][CP]
  float* const colpart = zpart + ROWS * WPR * CP;            // [ROWS][nout_p]
  float* const colpart_dz = colpart + ROWS * nout_p;         // [ROWS][CP]
  float* const dzs = colpart_dz + ROWS * CP;                 // [ROWS][CP]  dL/dlogits of each row, for its four warps
  uint64_t* const bar = reinterpret_cast<uint64_t*>(dzs + ROWS * CP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & (WPR - 1), rl = warp / WPR;          // warp of its row group, row slot of the CTA
  const int chunks = nout_p / COL_TILE;
  unsigned long long stamps[TRACE_SLOT_WORDS];
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tracing) {
#pragma unroll
    for (int w = 0; w < TRACE_SLOT_WORDS; ++w) stamps[w] = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[0]));
    stamps[1] = clock64();
    stamps[9] = p.rows; stamps[10] = p.cols; stamps[11] = p.Kin; stamps[12] = p.Nout; stamps[13] = 0; stamps[14] = gridDim.x;
  }
#define HEAD_STAMP(w) do { if (tracing) stamps[w] = clock64(); } while (0)
  const uint32_t table_bytes = (uint32_t)((p.Kin * CP + p.cols * nout_p) * 4);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    // The tables were written by a node that is complete before this kernel can start whenever the planner sets
    // w_early (see head_may_read_weights_early): then they are fetched while the predecessor is still running.
    if (p.w_early) {
      ptx::mbar_arrive_expect_tx(bar, table_bytes);
      ptx::bulk_load(w_fwd, p.tables, table_bytes, bar);
    }
  }
  HEAD_STAMP(2);
  pdl_wait();
  HEAD_STAMP(3);
  // The next kernel of the stream (normally the next adjoint contraction) cannot share an SM with this one (shared
  // memory): launched now, its CTAs take the idle SMs and then follow this kernel's CTAs as they leave - its launch
  // latency and prologue are hidden (measured: 2.7 us between the two kernels when triggered at the end).
  pdl_launch_dependents();
  if (threadIdx.x == 0 && !p.w_early) {
    ptx::mbar_arrive_expect_tx(bar, table_bytes);
    ptx::bulk_load(w_fwd, p.tables, table_bytes, bar);
  }
  float colacc[MAX_CHUNKS][4];
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) colacc[i][e] = 0.0f;
  float colacc_dz = 0.0f;
  const float scale = 0.0f - __fdiv_rn(p.DL[0], (float)p.rows);   // 0 - dL / toScalar(shape[0])
  __syncthreads();   // the barrier is initialised
  bool first = true;
  float* const myrow = arow + rl * p.Kin;
  const int tg = wr * 32 + lane;                             // thread of the row group

  for (int r0 = blockIdx.x * ROWS; r0 < p.rows; r0 += gridDim.x * ROWS) {
    const int r = r0 + rl;
    const bool live = r < p.rows;                            // (warp-uniform; barriers are still taken)
    // ---- requests of this row: operand planes (8 elements per thread), label, bias, mask source
    for (int k0 = 8 * tg; k0 < p.Kin && live; k0 += 8 * WPR * 32) {
      const uint4 h = *reinterpret_cast<const uint4*>(p.a_hi + (size_t)r * p.lda + k0);
      const uint4 m = *reinterpret_cast<const uint4*>(p.a_mid + (size_t)r * p.lda + k0);
      *reinterpret_cast<float4*>(myrow + k0) = make_float4(bf_lo(h.x) + bf_lo(m.x), bf_hi(h.x) + bf_hi(m.x), bf_lo(h.y) + bf_lo(m.y), bf_hi(h.y) + bf_hi(m.y));
      *reinterpret_cast<float4*>(myrow + k0 + 4) = make_float4(bf_lo(h.z) + bf_lo(m.z), bf_hi(h.z) + bf_hi(m.z), bf_lo(h.w) + bf_lo(m.w), bf_hi(h.w) + bf_hi(m.w));
    }
    const bool on = live && lane < p.cols;     // lane j owns class j in the softmax part (every warp of the row redundantly)
    const size_t idx = (size_t)r * p.cols + lane;
    const float yv = on ? p.Y[idx] : 0.0f;
    const float bv = (on && p.bias_in) ? p.bias_in[lane] : 0.0f;
    float4 aux[MAX_CHUNKS];
    if constexpr (kEpi == EPI_MASK_RELU || kEpi == EPI_MASK_LEAKY) {
#pragma unroll
      for (int i = 0; i < MAX_CHUNKS; ++i) {
        const int c0 = i * COL_TILE + wr * 128 + 4 * lane;
        if (i < chunks && c0 < p.Nout && live) aux[i] = *reinterpret_cast<const float4*>(p.Hm + (size_t)r * p.ldc + c0);
      }
    }
    if (first) {
      ptx::mbar_wait(bar, 0, 9);   // the tables have landed
      first = false;
    }
    row_group_sync(rl);            // the operand row is complete
    HEAD_STAMP(4);
    // ---- forward: lane (kg, jq) = (lane / 4, lane % 4) of warp wr sums classes 4 jq .. + 3 over k = kg + 8 (wr + WPR t)
    const int kg = lane >> 2, jq = lane & 3;
    float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int k = kg + 8 * wr; k < p.Kin; k += 8 * WPR) {
      const float a = myrow[k];
      const float4 w = *reinterpret_cast<const float4*>(w_fwd + k * CP + 4 * jq);
      z4[0] = fmaf(a, w.x, z4[0]); z4[1] = fmaf(a, w.y, z4[1]); z4[2] = fmaf(a, w.z, z4[2]); z4[3] = fmaf(a, w.w, z4[3]);
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) z4[u] += __shfl_xor_sync(0xffffffffu, z4[u], o);
    }
    if (lane < 4) *reinterpret_cast<float4*>(zpart + (rl * WPR + wr) * CP + 4 * lane) = make_float4(z4[0], z4[1], z4[2], z4[3]);
    row_group_sync(rl);
    // ---- softmax + crossEntropy and their adjoints: by ONE warp of the row (the arithmetic of fused_rows.cu, one
    //      class per lane; at most 16 classes, lanes 16.. hold zeros: four butterfly steps); the others wait for dz.
    //      (All four warps doing it redundantly was simpler - one barrier less - but 32 warps x ~250 instructions is
    //      2 000 issue cycles per scheduler for nothing.)
    if (wr == 0) {
      float z = 0.0f;
      if (lane < CP) {
#pragma unroll
        for (int w = 0; w < WPR; ++w) z += zpart[(rl * WPR + w) * CP + lane];   // fixed order
      }
      HEAD_STAMP(5);
      if (on && p.bias_in) z = __fadd_rn(z, bv);
      const float e = on ? expf(z) : 0.0f;
      float s = e;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      float t = 0.0f, dsum = 0.0f, dp = 0.0f, pr = 0.0f;
      if (on) {
        pr = __fdiv_rn(e, s);
        dp = __fdiv_rn(__fmul_rn(scale, yv), pr);
        t = __fmul_rn(__fdiv_rn(dp, s), e);
        dsum = __fmul_rn(0.0f - e, __fdiv_rn(dp, __fmul_rn(s, s)));
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      const float dz = on ? __fadd_rn(t, __fmul_rn(dsum, e)) : 0.0f;
      if (lane < CP) dzs[rl * CP + lane] = dz;
      if (on) {
        p.Z[idx] = z;
        p.P[idx] = pr;
        p.DP[idx] = dp;
        p.DH[idx] = dz;
        if (p.dh_hi) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(dz);
          p.dh_hi[(size_t)r * p.dh_ld + lane] = hi;
          p.dh_mid[(size_t)r * p.dh_ld + lane] = __float2bfloat16_rn(dz - __bfloat162float(hi));
        }
      }
      if (lane == 0 && live) {
        p.S[r] = s;
        p.DS[r] = dsum;
      }
      colacc_dz += dz;
    }
    row_group_sync(rl);   // dz of this row is in shared memory
    HEAD_STAMP(15);
    // ---- adjoint contraction: g[c] = sum_j dz[j] * w'[j, c], 4 columns per lane: warp wr owns columns
    //      [i COL_TILE + wr 128, + 128) of every chunk i
    float gacc[MAX_CHUNKS][4];
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS; ++i)
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) gacc[i][e2] = 0.0f;
#pragma unroll 2
    for (int jj = 0; jj < p.cols; ++jj) {
      const float d = dzs[rl * CP + jj];
      const float* wrow = w_bwd + jj * nout_p + wr * 128 + 4 * lane;
#pragma unroll
      for (int i = 0; i < MAX_CHUNKS; ++i) {
        if (i < chunks) {
          const float4 w0 = *reinterpret_cast<const float4*>(wrow + i * COL_TILE);
          gacc[i][0] = fmaf(d, w0.x, gacc[i][0]); gacc[i][1] = fmaf(d, w0.y, gacc[i][1]);
          gacc[i][2] = fmaf(d, w0.z, gacc[i][2]); gacc[i][3] = fmaf(d, w0.w, gacc[i][3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS; ++i) {
      const int c0 = i * COL_TILE + wr * 128 + 4 * lane;
      if (i >= chunks || c0 >= p.Nout || !live) continue;
      float (&gv)[4] = gacc[i];
      if (p.bias_out) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias_out + c0);
        gv[0] = __fadd_rn(gv[0], b.x); gv[1] = __fadd_rn(gv[1], b.y); gv[2] = __fadd_rn(gv[2], b.z); gv[3] = __fadd_rn(gv[3], b.w);
      }
      if (!(p.flags & GEMM_SKIP_C)) *reinterpret_cast<float4*>(p.C + (size_t)r * p.ldc + c0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      if constexpr (kEpi != EPI_NONE) {
        const float h4[4] = {aux[i].x, aux[i].y, aux[i].z, aux[i].w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          if constexpr (kEpi == EPI_RELU) gv[e2] = (0.0f <= gv[e2]) ? gv[e2] : 0.0f;
          else if constexpr (kEpi == EPI_LEAKY) gv[e2] = __fmul_rn((0.0f <= gv[e2]) ? 1.0f : p.epi_param, gv[e2]);
          else if constexpr (kEpi == EPI_MASK_RELU) gv[e2] = (0.0f <= h4[e2]) ? gv[e2] : 0.0f;
          else if constexpr (kEpi == EPI_MASK_LEAKY) gv[e2] = __fmul_rn(gv[e2], (0.0f <= h4[e2]) ? 1.0f : p.epi_param);
        }
        if (!(p.flags & GEMM_SKIP_D)) *reinterpret_cast<float4*>(p.D + (size_t)r * p.ldc + c0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
      if (p.flags & GEMM_SPLIT_OUT) {
        const uint32_t h0 = pack_bf16x2(gv[0], gv[1]), h1 = pack_bf16x2(gv[2], gv[3]);
        const uint32_t m0 = pack_bf16x2(gv[0] - bf_lo(h0), gv[1] - bf_hi(h0)), m1 = pack_bf16x2(gv[2] - bf_lo(h1), gv[3] - bf_hi(h1));
        *reinterpret_cast<uint2*>(p.out_hi + (size_t)r * p.ld_out + c0) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(p.out_mid + (size_t)r * p.ld_out + c0) = make_uint2(m0, m1);
      }
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) colacc[i][e2] += gv[e2];
    }
    row_group_sync(rl);   // the row buffer and the partial logits are rewritten by the next row
    HEAD_STAMP(16);
  }
  // ---- column sums (bias gradients of both layers): the rows of the CTA meet in shared memory, one atomic per
  //      column and block
  if (p.colsum_out) {
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS; ++i)
      if (i < chunks)
        *reinterpret_cast<float4*>(colpart + rl * nout_p + i * COL_TILE + wr * 128 + 4 * lane) = make_float4(colacc[i][0], colacc[i][1], colacc[i][2], colacc[i][3]);
  }
  if (wr == 0 && lane < CP) colpart_dz[rl * CP + lane] = colacc_dz;
  __syncthreads();
  if (p.colsum_out) {
    for (int c = threadIdx.x; c < p.Nout; c += HEAD_THREADS) {
      float sum = 0.0f;
#pragma unroll
      for (int w = 0; w < ROWS; ++w) sum += colpart[w * nout_p + c];
      atomicAdd(p.colsum_out + c, sum);
    }
  }
  if (p.colsum_dz && threadIdx.x < p.cols) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < ROWS; ++w) sum += colpart_dz[w * CP + threadIdx.x];
    atomicAdd(p.colsum_dz + threadIdx.x, sum);
  }
  if (tracing) {
    stamps[6] = clock64(); stamps[7] = stamps[6];
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamps[8]));
    const unsigned long long i = (unsigned long long)p.trace_index;
    if (p.trace[0] < i) p.trace[0] = i;
#pragma unroll
    for (int w = 0; w < TRACE_SLOT_WORDS; ++w) p.trace[i * TRACE_SLOT_WORDS + w] = stamps[w];
  }
}

int head_nout_p(int nout) { return (nout + COL_TILE - 1) / COL_TILE * COL_TILE; }

size_t head_smem_bytes(const HeadParams& p) {
  const size_t nout_p = (size_t)head_nout_p(p.Nout);
  return 4 * ((size_t)p.Kin * CP + (size_t)p.cols * nout_p + (size_t)ROWS * p.Kin + ROWS * WPR * CP + ROWS * nout_p + 2 * ROWS * CP) + 16;
}

}  // namespace

size_t head_table_floats(const HeadParams& p) { return (size_t)p.Kin * CP + (size_t)p.cols * head_nout_p(p.Nout); }

bool head_rows_supported(const HeadParams& p) {
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (p.cols < 1 || p.cols > CP || p.rows < 1) return false;
  if (p.Kin < 8 || (p.Kin & 7) || (p.lda & 7) || !a16(p.a_hi) || !a16(p.a_mid)) return false;
  if (p.Nout < 4 || (p.Nout & 3) || p.Nout > COL_TILE * MAX_CHUNKS) return false;
  if ((p.ldc & 3) || !a16(p.C) || (p.epi != EPI_NONE && !a16(p.D))) return false;
  if ((p.epi == EPI_MASK_RELU || p.epi == EPI_MASK_LEAKY) && !a16(p.Hm)) return false;
  if ((p.flags & GEMM_BIAS) && !a16(p.bias_out)) return false;
  if ((p.flags & GEMM_SPLIT_OUT) && ((p.ld_out & 3) || (reinterpret_cast<uintptr_t>(p.out_hi) & 7) || (reinterpret_cast<uintptr_t>(p.out_mid) & 7))) return false;
  if (p.epi != EPI_NONE && p.epi != EPI_RELU && p.epi != EPI_LEAKY && p.epi != EPI_MASK_RELU && p.epi != EPI_MASK_LEAKY) return false;
  return head_smem_bytes(p) <= 200 * 1024;
}

void launch_head_tables(Context& ctx, const HeadParams& p, cudaStream_t st) {
  const int total = (int)head_table_floats(p);
  {
    Launch l(ctx, KC_SPLIT, st);
    launch_kernel(ctx, head_tables_kernel, dim3((total + 255) / 256), dim3(256), 0, st, p, head_nout_p(p.Nout));
  }
  EGB_CUDA(cudaGetLastError());
}

void launch_head_rows(Context& ctx, const HeadParams& p, cudaStream_t st) {
  if (!head_rows_supported(p) || !p.tables) fail(EGB_ERR_GPU, "head kernel: unsupported configuration");
  typedef void (*KernelFn)(HeadParams);
  KernelFn fn = nullptr;
  switch (p.epi) {
    case EPI_NONE: fn = head_rows_kernel<EPI_NONE>; break;
    case EPI_RELU: fn = head_rows_kernel<EPI_RELU>; break;
    case EPI_LEAKY: fn = head_rows_kernel<EPI_LEAKY>; break;
    case EPI_MASK_RELU: fn = head_rows_kernel<EPI_MASK_RELU>; break;
    default: fn = head_rows_kernel<EPI_MASK_LEAKY>; break;
  }
  HeadParams q = p;
  q.trace = ctx.trace;
  q.trace_index = ctx.trace ? 1 + (int)(ctx.trace_next++ % (TRACE_SLOTS - 1)) : 0;
  const size_t smem = head_smem_bytes(p);
  EGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int blocks = (p.rows + ROWS - 1) / ROWS;
  {
    Launch l(ctx, KC_REDUCE, st);
    launch_kernel(ctx, fn, dim3(blocks < ctx.sm_count ? blocks : ctx.sm_count), dim3(HEAD_THREADS), smem, st, q);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
