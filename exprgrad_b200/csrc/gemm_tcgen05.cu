// fp32-accurate contraction on the 5th-gen tensor cores (sm_100a):
//
//   C[M,N] (+)= alpha * sum_k A[m,k] * B[n,k]        (both operands K-major)
//
// This is the device kernel behind the reference's contraction kernels
//   result[y, x] ++= a[y, it] * b[it, x] | (y, x, it)      (exprgrad/layers/base.nim:27-28,
//   exprgrad/layers/dnn.nim:19-21, benchmarks/matmul/matmul_gpu.nim:32)
// and their autodiff adjoints (exprgrad/passes.nim:399-403, 519-549).
//
// The tensor cores have no fp32 MMA. Each fp32 operand x is pre-split into two bf16 planes
// hi = bf16(x), mid = bf16(x - hi); the kernel accumulates the three products
//   hi*hi + hi*mid + mid*hi
// in the fp32 TMEM accumulator ("bf16x3"), which keeps the result within ~4e-6 (normalised) of
// the reference's sequential fp32 loop - far inside the 1e-4 parity bar.
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles of the four planes
//   warp 1      MMA issuer:  one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x 12 per k-block
//   warp 2      TMEM allocator (512 columns = two 128 x 256 fp32 accumulators, double buffered)
//   warps 4-7   epilogue: tcgen05.ld -> registers -> alpha / bias / relu / accumulate -> global
// Pipelines: smem full/empty mbarriers between TMA and MMA; TMEM full/empty between MMA and epilogue,
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>

#include "egb_internal.hpp"
#include "ptx.cuh"

namespace egb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                      // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_PLANE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int MN_GROUP_BYTES = BK * 128;    // MN-major: one group of 64 MN-elements x BK rows
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int NUM_THREADS = 256;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

struct KParams {
  float* C;
  const float* bias;
  float* D;            // second output: activation / masked gradient / updated parameter (same layout as C)
  const float* H;      // EPI_MASK_*: pre-activation the mask is computed from (same layout as C)
  float* colsum;       // += column sums of the final value (atomic)
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_mid;
  int ldc, ld_out;
  int epi;             // EpiMode
  float epi_param;     // leak (leaky relu) or rate (SGD)
  int M, N, K;
  int BN, stages;
  int tiles_m, tiles_n;
  int flags;
  float alpha;
  int a_mn, b_mn;  // operand is MN-major (stored [K, MN] row-major)
  // split-K (kFused kernel only): every output tile is computed by `splits` CTAs over disjoint
  // k-block ranges; partial tiles are added into C with RED.ADD and the CTA that arrives last at the
  // tile's counter re-reads the sum and runs the fused epilogue stages.
  int splits, kb_per_split;
  int* counters;
};

__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* mid, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  *hi = h;
  *mid = __float2bfloat16_rn(v - __bfloat162float(h));
}

// kFused = false: plain contraction epilogue (C (+)= alpha * acc), the 4096^3 benchmark path;
// kFused = true: bias / second stage / operand planes / column sums fused behind the contraction.
template <bool kFused>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_mid,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_mid,
                   const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_plane_bytes = p.BN * BK * 2;
  const int stage_bytes = 2 * A_PLANE_BYTES + 2 * b_plane_bytes;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_units = num_tiles * p.splits;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a_hi);
    ptx::prefetch_tensormap(&tm_a_mid);
    ptx::prefetch_tensormap(&tm_b_hi);
    ptx::prefetch_tensormap(&tm_b_mid);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_slot, TMEM_COLS);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch_dependents();
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the
      // previous kernel; operand planes written by it may only be read from here on. Every other role
      // is ordered behind these loads through the mbarrier pipeline.
      pdl_wait();
      uint32_t it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit / p.splits;
        const int kb0 = (unit % p.splits) * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int m0 = (tile % p.tiles_m) * BM;
        const int n0 = (tile / p.tiles_m) * p.BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1, 1);
          uint8_t* st = smem + s * stage_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            ptx::tma_load_2d(st, &tm_a_hi, &full_bar[s], k0, m0);
            ptx::tma_load_2d(st + A_PLANE_BYTES, &tm_a_mid, &full_bar[s], k0, m0);
          } else {
            // MN-major: one {64 (M), BK (K rows)} box per group of 64 M-elements
            for (int g = 0; g < BM / 64; ++g) {
              ptx::tma_load_2d(st + g * MN_GROUP_BYTES, &tm_a_hi, &full_bar[s], m0 + g * 64, k0);
              ptx::tma_load_2d(st + A_PLANE_BYTES + g * MN_GROUP_BYTES, &tm_a_mid, &full_bar[s], m0 + g * 64, k0);
            }
          }
          uint8_t* sb = st + 2 * A_PLANE_BYTES;
          if (!p.b_mn) {
            ptx::tma_load_2d(sb, &tm_b_hi, &full_bar[s], k0, n0);
            ptx::tma_load_2d(sb + b_plane_bytes, &tm_b_mid, &full_bar[s], k0, n0);
          } else {
            for (int g = 0; g < p.BN / 64; ++g) {
              ptx::tma_load_2d(sb + g * MN_GROUP_BYTES, &tm_b_hi, &full_bar[s], n0 + g * 64, k0);
              ptx::tma_load_2d(sb + b_plane_bytes + g * MN_GROUP_BYTES, &tm_b_mid, &full_bar[s], n0 + g * 64, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(BM, p.BN, p.a_mn != 0, p.b_mn != 0);
    uint32_t it = 0;
    uint32_t local_tile = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local_tile) {
      const uint32_t acc = local_tile & 1;
      const uint32_t use = local_tile >> 1;
      const int kb0 = (unit % p.splits) * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      ptx::mbar_wait(&tmem_empty[acc], (use & 1) ^ 1, 2);  // epilogue drained this accumulator
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        ptx::mbar_wait(&full_bar[s], ph, 3);  // TMA bytes have landed
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t st = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t sb = st + 2 * A_PLANE_BYTES;
          const uint64_t a_hi = p.a_mn ? ptx::make_mnmajor_sw128_desc(st, MN_GROUP_BYTES) : ptx::make_kmajor_sw128_desc(st);
          const uint64_t a_mid = p.a_mn ? ptx::make_mnmajor_sw128_desc(st + A_PLANE_BYTES, MN_GROUP_BYTES)
                                        : ptx::make_kmajor_sw128_desc(st + A_PLANE_BYTES);
          const uint64_t b_hi = p.b_mn ? ptx::make_mnmajor_sw128_desc(sb, MN_GROUP_BYTES) : ptx::make_kmajor_sw128_desc(sb);
          const uint64_t b_mid = p.b_mn ? ptx::make_mnmajor_sw128_desc(sb + b_plane_bytes, MN_GROUP_BYTES)
                                        : ptx::make_kmajor_sw128_desc(sb + b_plane_bytes);
          // advancing by UMMA_K = 16 along K: K-major: 32 bytes inside the 128-byte swizzle row;
          // MN-major: 16 rows of 128 bytes (address field is in 16-byte units)
          const uint64_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
          const uint64_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t aa = a_step * k, ba = b_step * k;
            // small cross terms first, dominant hi*hi last
            ptx::umma_f16<1>(d_tmem, a_mid + aa, b_hi + ba, idesc, (kb != kb0) || (k != 0));
            ptx::umma_f16<1>(d_tmem, a_hi + aa, b_mid + ba, idesc, 1);
            ptx::umma_f16<1>(d_tmem, a_hi + aa, b_hi + ba, idesc, 1);
          }
          ptx::umma_commit(&empty_bar[s]);                          // smem slot free when these MMAs retire
          if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    // Thread t of warp q owns accumulator row (q*32 + t); it walks the tile 32 columns at a time:
    //   v = acc [+ bias[c]] [+ C_old]          -> C   (the contraction's own output tensor)
    //   w = second stage (activation / gradient mask / SGD update) of v   -> D
    //   [bf16 hi/mid planes of w]  [column sums of w -> colsum]
    // Every stage restates one reference kernel that would otherwise run as a separate launch
    // (bias: dnn.nim:22-24, relu/leakyRelu: dnn.nim:26-30 and their derive()d adjoints,
    // gradientDescent: base.nim:37-38); arithmetic is kept un-contracted (__fmul_rn/__fadd_rn).
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t local_tile = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.H) & 15) == 0);
    // one 32-column chunk of one accumulator row through the fused stages; `v` holds alpha * acc, or
    // (split-K, last CTA of the tile) the complete sum read back from C
    auto finish_chunk = [&](float (&v)[32], const int row, const int col0, const bool from_memory) {
      const int ncols = min(32, p.N - col0);
      const bool row_ok = row < p.M;
      const bool full = vec_ok && ncols == 32;
      const size_t off = (size_t)row * p.ldc + col0;
        if (kFused && (p.flags & GEMM_BIAS)) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) v[j] = __fadd_rn(v[j], __ldg(p.bias + col0 + j));
      }
      if (row_ok) {
        if (!from_memory && (p.flags & GEMM_ACCUMULATE)) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o = *reinterpret_cast<const float4*>(p.C + off + j);
              v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) v[j] += p.C[off + j];
          }
        }
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(p.C + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) p.C[off + j] = v[j];
        }
        // ---- second stage
        if (!kFused) {
        } else if (p.epi == EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (0.0f <= v[j]) ? v[j] : 0.0f;
        } else if (p.epi == EPI_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fmul_rn((0.0f <= v[j]) ? 1.0f : p.epi_param, v[j]);
        } else if (p.epi == EPI_MASK_RELU || p.epi == EPI_MASK_LEAKY) {
          float h[32];
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o = *reinterpret_cast<const float4*>(p.H + off + j);
              h[j] = o.x; h[j + 1] = o.y; h[j + 2] = o.z; h[j + 3] = o.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = (j < ncols) ? p.H[off + j] : 0.0f;
          }
          if (p.epi == EPI_MASK_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (0.0f <= h[j]) ? v[j] : 0.0f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(v[j], (0.0f <= h[j]) ? 1.0f : p.epi_param);
          }
        } else if (p.epi == EPI_SGD) {
          // P += (0 - g) * rate   (base.nim:37-38; negate is `0 - x`, llvm.nim:333-336)
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o = *reinterpret_cast<const float4*>(p.D + off + j);
              v[j] = __fadd_rn(o.x, __fmul_rn(0.0f - v[j], p.epi_param));
              v[j + 1] = __fadd_rn(o.y, __fmul_rn(0.0f - v[j + 1], p.epi_param));
              v[j + 2] = __fadd_rn(o.z, __fmul_rn(0.0f - v[j + 2], p.epi_param));
              v[j + 3] = __fadd_rn(o.w, __fmul_rn(0.0f - v[j + 3], p.epi_param));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) v[j] = __fadd_rn(p.D[off + j], __fmul_rn(0.0f - v[j], p.epi_param));
          }
        }
        if (kFused && p.epi != EPI_NONE) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(p.D + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) p.D[off + j] = v[j];
          }
        }
        if (kFused && (p.flags & GEMM_SPLIT_OUT)) {
          __nv_bfloat16* hrow = p.out_hi + (size_t)row * p.ld_out + col0;
          __nv_bfloat16* mrow = p.out_mid + (size_t)row * p.ld_out + col0;
          if (ncols == 32 && (p.ld_out & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __align__(16) __nv_bfloat16 hv[8];
              __align__(16) __nv_bfloat16 mv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) split_store(&hv[e], &mv[e], v[j + e]);
              *reinterpret_cast<uint4*>(hrow + j) = *reinterpret_cast<const uint4*>(hv);
              *reinterpret_cast<uint4*>(mrow + j) = *reinterpret_cast<const uint4*>(mv);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) split_store(hrow + j, mrow + j, v[j]);
          }
        }
      }
      if (kFused && p.colsum) {
        // column sums over the 32 rows this warp holds: butterfly transpose-reduce (31 shuffles);
        // afterwards lane j holds the sum of column j. Rows outside the matrix contribute zero.
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!row_ok || j >= ncols) v[j] = 0.0f;
#pragma unroll
        for (int offw = 16; offw >= 1; offw >>= 1) {
          const bool upper = (lane & offw) != 0;
#pragma unroll
          for (int j = 0; j < offw; ++j) {
            const float send = upper ? v[j] : v[j + offw];
            const float keep = upper ? v[j + offw] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, offw);
          }
        }
        if (lane < ncols) atomicAdd(p.colsum + col0 + lane, v[0]);
      }
    };
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local_tile) {
      const int tile = unit / p.splits;
      const uint32_t acc = local_tile & 1;
      const uint32_t use = local_tile >> 1;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      ptx::mbar_wait(&tmem_full[acc], use & 1, 4);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      const bool split = kFused && p.splits > 1;
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c, r);
        ptx::tmem_ld_wait();
        const int col0 = n0 + c;
        if (col0 >= p.N) break;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (!split) {
          finish_chunk(v, row, col0, false);
        } else if (row < p.M) {
          // partial tile: add into C (zero-filled or holding the value to accumulate onto)
          float* crow = p.C + (size_t)row * p.ldc + col0;
          const int ncols = min(32, p.N - col0);
          if (vec_ok && ncols == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)   // 16-byte vector reductions: 4x fewer L2 atomic operations
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) atomicAdd(crow + j, v[j]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (split) {
        // the CTA that increments the tile counter last owns the complete sum
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 128) {
          const int prev = atomicAdd(p.counters + tile, 1);
          const bool last = prev == p.splits - 1;
          if (last) p.counters[tile] = 0;  // self-resetting for the next launch
          tmem_slot[1] = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool last = tmem_slot[1] != 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // flag may be rewritten by the next unit
        if (last) {
          __threadfence();
          for (int c = 0; c < p.BN; c += 32) {
            const int col0 = n0 + c;
            if (col0 >= p.N) break;
            float v[32];
            const int ncols = min(32, p.N - col0);
            const float* crow = p.C + (size_t)row * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (row < p.M && j < ncols) ? __ldcg(crow + j) : 0.0f;
            finish_chunk(v, row, col0, true);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

// K-major: the plane is [rows = MN extent, K] -> box {BK, box_rows}. MN-major: the plane is
// [K, cols = MN extent] -> box {64, BK}.
void encode_plane(Context& ctx, CUtensorMap* tm, const __nv_bfloat16* base, int mn, int K, int ld, int box_rows,
                  bool mn_major) {
  if ((ld & 7) != 0) fail(EGB_ERR_GPU, "bf16 plane leading dimension %d is not a multiple of 8", ld);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) fail(EGB_ERR_GPU, "bf16 plane is not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)mn};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  if (mn_major) {
    dims[0] = (cuuint64_t)mn;
    dims[1] = (cuuint64_t)K;
    box[0] = 64;
    box[1] = BK;
  }
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx.encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

int choose_bn(int M, int N, int sm_count, bool b_mn) {
  // Smallest tile count that still fills the machine wins; prefer wide tiles (more operand reuse).
  const int tiles_m = (M + BM - 1) / BM;
  const int step = b_mn ? 64 : 32;  // an MN-major B tile is made of 64-column groups
  int best = step;
  double best_cost = 1e30;
  for (int bn = 256; bn >= step; bn -= step) {
    const int tiles_n = (N + bn - 1) / bn;
    const long tiles = (long)tiles_m * tiles_n;
    const long waves = (tiles + sm_count - 1) / sm_count;
    // time ~ waves * bn (MMA work per tile) with a small bonus for wider tiles (less smem traffic / flop)
    double cost = (double)waves * bn * (1.0 + 16.0 / bn);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

void gemm_choose_config(int M, int N, int K, bool b_mn, int sm_count, int* bn, int* splits, int* tiles) {
  *bn = choose_bn(M, N, sm_count, b_mn);
  *tiles = ((M + BM - 1) / BM) * ((N + *bn - 1) / *bn);
  const int num_kb = (K + BK - 1) / BK;
  int s = 1;
  // few tiles and a long reduction: one SM per tile would stream its whole K extent at the per-SM TMA
  // rate (~100 GB/s) while the rest of the machine idles - split the reduction instead
  // (measured on the dense step: splitting pays only when very few SMs would be busy - the zero fill,
  // L2 reductions and re-read of C cost more than a 2x shorter k-loop saves on mid-sized grids)
  if (*tiles * 8 <= sm_count && num_kb >= 4) {
    s = sm_count / *tiles;
    if (s > num_kb / 2) s = num_kb / 2;
    if (s > 8) s = 8;
    if (s < 1) s = 1;
  }
  *splits = s;
}

void launch_gemm_bf16x3(Context& ctx, const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return;
  if (a.K <= 0) fail(EGB_ERR_GPU, "gemm: K must be positive");
  if (!ctx.encode_tiled) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled entry point not available");
  if (a.bn == 0 && gemm_2cta_eligible(a)) {
    launch_gemm_bf16x3_2cta(ctx, a, st);
    return;
  }
  KParams p;
  p.C = a.C; p.bias = a.bias; p.out_hi = a.out_hi; p.out_mid = a.out_mid;
  p.D = a.D; p.H = a.H; p.colsum = a.colsum;
  p.epi = a.epi; p.epi_param = a.epi_param;
  if (a.flags & GEMM_RELU) {  // legacy flag of the raw entry point: relu written in place of C
    fail(EGB_ERR_GPU, "gemm: use epi = EPI_RELU with a second output instead of GEMM_RELU");
  }
  if (a.epi != EPI_NONE && !a.D) fail(EGB_ERR_GPU, "gemm: fused second stage needs an output tensor");
  if ((a.epi == EPI_MASK_RELU || a.epi == EPI_MASK_LEAKY) && !a.H) fail(EGB_ERR_GPU, "gemm: mask stage needs H");
  p.ldc = a.ldc; p.ld_out = a.ld_out;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.flags = a.flags; p.alpha = a.alpha;
  p.a_mn = a.a_mn ? 1 : 0;
  p.b_mn = a.b_mn ? 1 : 0;
  p.counters = a.counters;
  p.BN = a.bn > 0 ? a.bn : choose_bn(a.M, a.N, ctx.sm_count, a.b_mn);
  if (p.BN % 32 != 0 || p.BN < 32 || p.BN > 256) fail(EGB_ERR_GPU, "gemm: invalid BN %d", p.BN);
  if (a.b_mn && p.BN % 64 != 0) fail(EGB_ERR_GPU, "gemm: BN must be a multiple of 64 for an MN-major B operand");
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + p.BN - 1) / p.BN;
  const int stage_bytes = 2 * A_PLANE_BYTES + 2 * p.BN * BK * 2;
  const int bar_bytes = (2 * MAX_STAGES + 4) * 8 + 16;
  int stages = (SMEM_LIMIT - 1024 - bar_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) fail(EGB_ERR_GPU, "gemm: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + bar_bytes;

  CUtensorMap tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid;
  encode_plane(ctx, &tm_a_hi, a.a_hi, a.M, a.K, a.lda, BM, a.a_mn);
  encode_plane(ctx, &tm_a_mid, a.a_mid, a.M, a.K, a.lda, BM, a.a_mn);
  encode_plane(ctx, &tm_b_hi, a.b_hi, a.N, a.K, a.ldb, p.BN, a.b_mn);
  encode_plane(ctx, &tm_b_mid, a.b_mid, a.N, a.K, a.ldb, p.BN, a.b_mn);

  static bool attr_set = false;
  if (!attr_set) {
    EGB_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    EGB_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  static const bool force_fused = getenv("EGB_GEMM_FUSED_ALWAYS") != nullptr;
  const int tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (a.K + BK - 1) / BK;
  int splits = (a.splits > 1 && a.counters) ? a.splits : 1;
  if (splits > num_kb) splits = num_kb;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  const int units = tiles * p.splits;
  const bool fused = force_fused || (a.flags & (GEMM_BIAS | GEMM_SPLIT_OUT)) || a.epi != EPI_NONE || a.colsum || p.splits > 1;
  const int grid = units < ctx.sm_count ? units : ctx.sm_count;
  {
    Launch l(ctx, KC_GEMM, st);
    if (fused)
      launch_kernel(ctx, gemm_bf16x3_kernel<true>, dim3(grid), dim3(NUM_THREADS), smem, st, tm_a_hi, tm_a_mid, tm_b_hi,
                    tm_b_mid, p);
    else
      launch_kernel(ctx, gemm_bf16x3_kernel<false>, dim3(grid), dim3(NUM_THREADS), smem, st, tm_a_hi, tm_a_mid, tm_b_hi,
                    tm_b_mid, p);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
