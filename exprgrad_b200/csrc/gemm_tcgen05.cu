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

#include <set>

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
constexpr int NUM_THREADS = 384;  // TMA, MMA, TMEM-allocator, idle warp + 8 epilogue warps
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
  // cluster split-K: the `ck` CTAs of a thread-block cluster each accumulate a disjoint k-block range of
  // ONE output tile in their own TMEM, stage the partial tile in shared memory and reduce it through
  // distributed shared memory; CTA r then runs the fused epilogue on rows [r*128/ck, (r+1)*128/ck).
  // No atomics, no zero fill, fixed summation order. (splits == ck > 1: one unit per CTA.)
  int ck;
  int splits, kb_per_split;
  unsigned long long* trace;  // debug timeline (Context::trace) or null
  int trace_index;            // slot of this launch
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// slot words: 0 gtime entry, 1 clk entry, 2 clk setup done, 3 clk pdl_wait passed, 4 clk first operands landed,
// 5 clk accumulator complete, 6 clk epilogue done, 7 clk exit, 8 gtime exit, 9 M, 10 N, 11 K, 12 BN, 13 ck, 14 grid
// Stamps are collected in shared memory and copied out once at kernel exit: a store to the host-mapped
// trace buffer in the middle of a phase stalls the warp's later stores behind a PCIe write (observer effect).
#define EGB_TRACE(word)                                                        \
  do {                                                                         \
    if (trace_slot) trace_sm[word] = (unsigned long long)clock64();            \
  } while (0)

constexpr int STAGING_PAD = 4;  // floats; keeps 16-byte row-strided accesses conflict-free
constexpr int UNIT_COLS = 16;                             // an epilogue work unit is 32 rows x 16 columns
constexpr int WSTAGE_LD = UNIT_COLS + STAGING_PAD;        // row stride of a warp's transposition buffer
constexpr int EPI_WARPS = 8;
constexpr int WSTAGE_BYTES = EPI_WARPS * 32 * WSTAGE_LD * 4;
constexpr int BAR_BYTES = (2 * MAX_STAGES + 4) * 8 + 32 + TRACE_SLOT_WORDS * 8;  // mbarriers, TMEM slot, trace pointer + stamps

__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t cta_rank) {
  float4 v;
  asm volatile(
      "{\n"
      ".reg .b32 remote;\n"
      "mapa.shared::cluster.u32 remote, %4, %5;\n"
      "ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [remote];\n"
      "}\n"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "r"(local_addr), "r"(cta_rank)
      : "memory");
  return v;
}

__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* mid, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  *hi = h;
  *mid = __float2bfloat16_rn(v - __bfloat162float(h));
}

// kFused = false: plain contraction epilogue (C (+)= alpha * acc), the 4096^3 benchmark path;
// kFused = true: bias / second stage / operand planes / column sums fused behind the contraction.
// kEpi: the second stage (EpiMode) is a COMPILE-TIME parameter. With all stages behind run-time branches the
// kernel was 16 296 SASS instructions (260 KB); its epilogue runs once per launch with a cold instruction cache,
// and ncu showed the epilogue warps stalled on instruction fetch (stall_no_instruction 14 per issue,
// profiles/r02d_dense_gemm_ncu.txt): 2.3 us for the ~600 instructions that finish one 32 x 16 unit.
template <bool kFused, int kEpi>
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
  unsigned long long** trace_slot_shp = reinterpret_cast<unsigned long long**>(tmem_slot + 4);
  unsigned long long* trace_sm = reinterpret_cast<unsigned long long*>(tmem_slot + 8);  // TRACE_SLOT_WORDS stamps
  if (threadIdx.x == 0) {
    unsigned long long* slot = nullptr;
    if (p.trace && blockIdx.x == 0) {
      const unsigned long long i = (unsigned long long)p.trace_index;  // fixed per launch site: graph replays overwrite
      if (p.trace[0] < i) p.trace[0] = i;  // (only CTA 0 of each launch writes: per-CTA atomics on the host-mapped
                                           // buffer delay kernel completion by > 100 us - observer effect)
      {
        slot = p.trace + i * TRACE_SLOT_WORDS;
        for (int w = 0; w < TRACE_SLOT_WORDS; ++w) trace_sm[w] = 0;
        trace_sm[0] = globaltimer_ns(); trace_sm[1] = (unsigned long long)clock64();
        trace_sm[9] = p.M; trace_sm[10] = p.N; trace_sm[11] = p.K; trace_sm[12] = p.BN; trace_sm[13] = p.ck; trace_sm[14] = gridDim.x;
      }
    }
    *trace_slot_shp = slot;
  }

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
      ptx::mbar_init(&tmem_empty[a], 8);  // one arrive per epilogue warp
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
  unsigned long long* const trace_slot = *trace_slot_shp;
  if (threadIdx.x == 0) EGB_TRACE(2);

  if (warp == 0) {
    // ===================================================== TMA producer
    // The whole warp walks the loop convergently and `elect.sync` guards the issue: with `if (lane == 0)`
    // around the loop nvcc feeds every uniform-register operand of UTMALDG / UTCHMMA through an
    // ELECT / R2UR / BRA.U.ANY waterfall (~100 cycles per instruction: small contractions were
    // issue-bound at 0.62 us per 64-deep k-block); under elect.sync it emits them back to back.
    // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the
    // previous kernel; operand planes written by it may only be read from here on. Every other role
    // is ordered behind these loads through the mbarrier pipeline.
    pdl_wait();
    if (lane == 0) EGB_TRACE(3);
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
        const int k0 = kb * BK;
        if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        if (!p.a_mn) {
          ptx::tma_load_2d(st, &tm_a_hi, &full_bar[s], k0, m0);
          ptx::tma_load_2d(st + A_PLANE_BYTES, &tm_a_mid, &full_bar[s], k0, m0);
        } else {
          // MN-major: one {64 (M), BK (K rows)} box per group of 64 M-elements
#pragma unroll
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
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (convergent warp, lane 0 issues)
    const uint32_t idesc = ptx::make_idesc_bf16_f32(BM, p.BN, p.a_mn != 0, p.b_mn != 0);
    // advancing by UMMA_K = 16 along K: K-major: 32 bytes inside the 128-byte swizzle row;
    // MN-major: 16 rows of 128 bytes (address field is in 16-byte units)
    const uint64_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    const uint64_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    // descriptors of stage 0; a stage / plane offset only moves the 14-bit start-address field
    const uint32_t smem0 = ptx::smem_u32(smem);
    const uint64_t a_desc0 = p.a_mn ? ptx::make_mnmajor_sw128_desc(smem0, MN_GROUP_BYTES) : ptx::make_kmajor_sw128_desc(smem0);
    const uint64_t b_desc0 = p.b_mn ? ptx::make_mnmajor_sw128_desc(smem0, MN_GROUP_BYTES) : ptx::make_kmajor_sw128_desc(smem0);
    const uint64_t a_mid_off = (uint64_t)(A_PLANE_BYTES >> 4);
    const uint64_t b_off = (uint64_t)((2 * A_PLANE_BYTES) >> 4);
    const uint64_t b_mid_off = b_off + (uint64_t)(b_plane_bytes >> 4);
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
        if (it == 0 && lane == 0) EGB_TRACE(4);
        const uint64_t so = (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        const uint64_t a_hi = a_desc0 + so, a_mid = a_hi + a_mid_off;
        const uint64_t b_hi = b_desc0 + so + b_off, b_mid = b_desc0 + so + b_mid_off;
        if (ptx::elect_one()) {
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
    // ===================================================== epilogue (8 warps)
    // Warps q and q + 4 of this group share TMEM lane quarter q (accumulator rows q*32 .. q*32+31) and
    // alternate over its 16-column units. tcgen05.ld hands thread t row t with consecutive columns - a
    // layout in which every global access of a warp touches 32 different rows (32 LSU wavefronts per
    // instruction). Each 32 x 16 unit is therefore transposed through a warp-private padded buffer into
    // the "coalesced layout": thread t holds rows (t>>2) + 8i, i = 0..3, columns 4*(t&3) .. +3, so one
    // warp instruction covers 8 rows x 64 contiguous bytes (whole 32-byte sectors). All fused stages run
    // in that layout:
    //   v = alpha * acc [+ bias[c]] [+ C_old]   -> C   (the contraction's own output tensor)
    //   w = second stage (activation / gradient mask / SGD update) of v   -> D
    //   [bf16 hi/mid planes of w]  [column sums of w -> colsum]
    // Every stage restates one reference kernel that would otherwise run as a separate launch
    // (bias: dnn.nim:22-24, relu/leakyRelu: dnn.nim:26-30 and their derive()d adjoints,
    // gradientDescent: base.nim:37-38); arithmetic is kept un-contracted (__fmul_rn/__fadd_rn).
    // Eight warps rather than four: the epilogue is a long dependent instruction stream, and a single
    // warp per scheduler ran it at ~5 cycles per instruction (timeline: 3 us per 32 x 32 block).
    const int q = warp & 3;
    const int ew = warp - 4;          // 0..7
    const int eh = ew >> 2;           // which of the two warps of the lane quarter
    float* wstage = reinterpret_cast<float*>(smem + p.stages * stage_bytes + BAR_BYTES) + ew * (32 * WSTAGE_LD);
    const int lr = lane >> 2;         // row of this thread inside a group of 8 rows
    const int c4 = (lane & 3) * 4;    // first of its 4 columns inside the 16-column unit
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.H) & 15) == 0);
    const bool planes_vec = ((p.ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 7) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out_mid) & 7) == 0);
    const bool need_c = (p.flags & GEMM_ACCUMULATE) != 0;
    constexpr bool need_h = kFused && (kEpi == EPI_MASK_RELU || kEpi == EPI_MASK_LEAKY);
    constexpr bool need_d = kFused && kEpi == EPI_SGD;
    // One 32-row x 16-column unit in the coalesced layout: v[i] = alpha * acc of row row0 + lr + 8i.
    // Phase 1 of a unit: every global read (bias, old C, mask source / parameter) is issued before the first
    // store - the pointers may alias as far as the compiler knows, so loads interleaved with stores would
    // serialise into one L2 round trip per row group. For the first unit of a warp these loads are issued
    // even before the accumulator is complete (they do not depend on it).
    struct UnitLoads {
      float b[4];
      float4 oldc[4], aux[4];  // aux: H (mask source) or the parameter the SGD stage updates
      uint32_t valid;
    };
    auto unit_loads = [&](UnitLoads& L, const int row0, const int nrows, const int col0) {
      const int ncols = min(UNIT_COLS, p.N - col0);
      const int cnt = max(0, min(4, ncols - c4));  // valid columns of this thread
      const bool full = vec_ok && cnt == 4;
      float (&b)[4] = L.b;
      float4 (&oldc)[4] = L.oldc;
      float4 (&aux)[4] = L.aux;
#pragma unroll
      for (int e = 0; e < 4; ++e) b[e] = 0.0f;
      if (kFused && (p.flags & GEMM_BIAS)) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < cnt) b[e] = __ldg(p.bias + col0 + c4 + e);
      }
      uint32_t valid = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lr + 8 * i;
        const int row = row0 + r;
        oldc[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        aux[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (r >= nrows || row >= p.M || cnt == 0) continue;
        valid |= 1u << i;
        const size_t off = (size_t)row * p.ldc + col0 + c4;
        const float* auxp = need_h ? p.H : p.D;
        if (full) {
          if (need_c) oldc[i] = *reinterpret_cast<const float4*>(p.C + off);
          if (need_h || need_d) aux[i] = *reinterpret_cast<const float4*>(auxp + off);
        } else {
          float o[4] = {0.0f, 0.0f, 0.0f, 0.0f}, a2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < cnt) {
              if (need_c) o[e] = p.C[off + e];
              if (need_h || need_d) a2[e] = auxp[off + e];
            }
          oldc[i] = make_float4(o[0], o[1], o[2], o[3]);
          aux[i] = make_float4(a2[0], a2[1], a2[2], a2[3]);
        }
      }
      L.valid = valid;
    };
    // Phase 2 of a unit in the coalesced layout: v[i] = alpha * acc of row row0 + lr + 8i.
    auto finish_unit = [&](float4 (&v)[4], const UnitLoads& L, const int row0, const int col0) {
      const int ncols = min(UNIT_COLS, p.N - col0);
      const int cnt = max(0, min(4, ncols - c4));  // valid columns of this thread
      const bool full = vec_ok && cnt == 4;
      const float (&b)[4] = L.b;
      const float4 (&oldc)[4] = L.oldc;
      const float4 (&aux)[4] = L.aux;
      const uint32_t valid = L.valid;
      float cs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!((valid >> i) & 1u)) continue;
        const int row = row0 + lr + 8 * i;
        const size_t off = (size_t)row * p.ldc + col0 + c4;
        float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        if (kFused && (p.flags & GEMM_BIAS)) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = __fadd_rn(x[e], b[e]);
        }
        if (need_c) {
          x[0] += oldc[i].x; x[1] += oldc[i].y; x[2] += oldc[i].z; x[3] += oldc[i].w;
        }
        if (p.flags & GEMM_SKIP_C) {
          // dead store: nothing reads the fp32 C after this epilogue
        } else if (full) {
          *reinterpret_cast<float4*>(p.C + off) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < cnt) p.C[off + e] = x[e];
        }
        if (kFused) {
          // ---- second stage
          const float a2[4] = {aux[i].x, aux[i].y, aux[i].z, aux[i].w};
          if constexpr (kEpi == EPI_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (0.0f <= x[e]) ? x[e] : 0.0f;
          } else if constexpr (kEpi == EPI_LEAKY) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fmul_rn((0.0f <= x[e]) ? 1.0f : p.epi_param, x[e]);
          } else if constexpr (kEpi == EPI_MASK_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (0.0f <= a2[e]) ? x[e] : 0.0f;
          } else if constexpr (kEpi == EPI_MASK_LEAKY) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fmul_rn(x[e], (0.0f <= a2[e]) ? 1.0f : p.epi_param);
          } else if constexpr (kEpi == EPI_SIGMOID) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(0.0f, x[e]))));
          } else if constexpr (kEpi == EPI_TANH) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ep = expf(x[e]), en = expf(__fsub_rn(0.0f, x[e]));
              x[e] = __fdiv_rn(__fsub_rn(ep, en), __fadd_rn(ep, en));
            }
          } else if constexpr (kEpi == EPI_SGD) {
            // P += (0 - g) * rate   (base.nim:37-38; negate is `0 - x`, llvm.nim:333-336)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fadd_rn(a2[e], __fmul_rn(0.0f - x[e], p.epi_param));
          }
          if (kEpi != EPI_NONE && !(p.flags & GEMM_SKIP_D)) {
            if (full) {
              *reinterpret_cast<float4*>(p.D + off) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < cnt) p.D[off + e] = x[e];
            }
          }
          if (p.flags & GEMM_SPLIT_OUT) {
            __nv_bfloat16* hrow = p.out_hi + (size_t)row * p.ld_out + col0 + c4;
            __nv_bfloat16* mrow = p.out_mid + (size_t)row * p.ld_out + col0 + c4;
            __align__(8) __nv_bfloat16 hv[4];
            __align__(8) __nv_bfloat16 mv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_store(&hv[e], &mv[e], x[e]);
            if (planes_vec && cnt == 4) {
              *reinterpret_cast<uint2*>(hrow) = *reinterpret_cast<const uint2*>(hv);
              *reinterpret_cast<uint2*>(mrow) = *reinterpret_cast<const uint2*>(mv);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < cnt) { hrow[e] = hv[e]; mrow[e] = mv[e]; }
            }
          }
          if (p.colsum) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < cnt) cs[e] += x[e];
          }
        }
      }
      if (kFused && p.colsum) {
        // the 8 lanes that share a column quad (lane ^ 4, ^ 8, ^ 16) hold the other rows
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 4);
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
        }
        if (lane < 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < cnt) atomicAdd(p.colsum + col0 + c4 + e, cs[e]);
        }
      }
    };
    // Loads of this warp's first unit, issued while the main loop is still running. They read tensors that
    // earlier kernels produced, so the warp first joins the programmatic-launch dependency.
    UnitLoads pre;
    bool have_pre = false;
    pdl_wait();
    if ((int)blockIdx.x < num_units) {
      const int tile = blockIdx.x / p.splits;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      if (p.ck > 1) {
        const int rows_per = BM / p.ck, nunits = p.BN / UNIT_COLS;
        const int groups = (rows_per + 31) / 32;
        if (ew < groups * nunits) {
          const int g = ew / nunits, c = (ew % nunits) * UNIT_COLS;
          if (n0 + c < p.N) {
            unit_loads(pre, m0 + (int)ptx::cluster_ctarank() * rows_per + g * 32, min(32, rows_per - g * 32), n0 + c);
            have_pre = true;
          }
        }
      } else if (eh * UNIT_COLS < p.BN && n0 + eh * UNIT_COLS < p.N) {
        unit_loads(pre, m0 + q * 32, 32, n0 + eh * UNIT_COLS);
        have_pre = true;
      }
    }
    uint32_t local_tile = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local_tile) {
      const int tile = unit / p.splits;
      const uint32_t acc = local_tile & 1;
      const uint32_t use = local_tile >> 1;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      ptx::mbar_wait(&tmem_full[acc], use & 1, 4);
      ptx::tc_fence_after();
      // Programmatic dependent launch, triggered late: the dependent grid may be scheduled once every CTA
      // has reached its LAST epilogue, so that its launch latency (~4.5 us as a full graph edge) overlaps
      // this epilogue. Triggering at kernel start instead parks the dependents' CTAs - one whole SM each -
      // in griddepcontrol.wait and starves the contractions of the parallel branches (measured).
      if (unit + (int)gridDim.x >= num_units) pdl_launch_dependents();
      if (local_tile == 0 && threadIdx.x == 128) EGB_TRACE(5);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      if (p.ck > 1) {
        // cluster split-K: park the raw partial accumulator in shared memory (the operand stages are idle:
        // tmem_full means every MMA of this CTA has retired and this CTA loads nothing else)
        float* stage_row = reinterpret_cast<float*>(smem) + (size_t)(q * 32 + lane) * (p.BN + STAGING_PAD);
        for (int c = eh * UNIT_COLS; c < p.BN; c += 2 * UNIT_COLS) {
          if (n0 + c >= p.N) break;  // warp-uniform
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(t_row + c, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(stage_row + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        ptx::tc_fence_before();
        break;  // exactly one unit per CTA
      }
      for (int c = eh * UNIT_COLS; c < p.BN; c += 2 * UNIT_COLS) {
        const int col0 = n0 + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(t_row + c, r);
        ptx::tmem_ld_wait();
        if (local_tile == 0 && c == 0 && threadIdx.x == 128) EGB_TRACE(15);
        float* mine = wstage + lane * WSTAGE_LD;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(mine + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = *reinterpret_cast<const float4*>(wstage + (lr + 8 * i) * WSTAGE_LD + c4);
          v[i].x *= p.alpha; v[i].y *= p.alpha; v[i].z *= p.alpha; v[i].w *= p.alpha;
        }
        __syncwarp();  // the staging block is rewritten by the next unit
        if (local_tile == 0 && c == 0 && threadIdx.x == 128) EGB_TRACE(16);
        if (!(have_pre && local_tile == 0 && c == eh * UNIT_COLS)) unit_loads(pre, m0 + q * 32, 32, col0);
        finish_unit(v, pre, m0 + q * 32, col0);
        if (local_tile == 0 && c == 0 && threadIdx.x == 128) EGB_TRACE(17);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
    if (p.ck > 1) {
      __syncwarp();
      ptx::cluster_sync_all();  // every CTA of the cluster has staged its partial tile
      if (threadIdx.x == 128) EGB_TRACE(15);
      const int tile = blockIdx.x / p.ck;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      const uint32_t crank = ptx::cluster_ctarank();
      const int rows_per = BM / p.ck;                 // rows of the tile this CTA finishes
      const int groups = (rows_per + 31) / 32;
      const int nunits = p.BN / UNIT_COLS;
      const uint32_t stage_base = ptx::smem_u32(smem);
      // This warp's units: all partial sums are read first; the cluster is then told that this warp is done
      // with its peers' staging buffers (arrive) before the fused stages run, so no CTA waits for another
      // one's epilogue - only for its reads.
      constexpr int MAX_U = 4;  // (BN / 16 units) * (rows_per / 32 groups) / 8 warps <= 4 for ck >= 2
      float4 v[MAX_U][4];
      const int total_u = groups * nunits;
#pragma unroll
      for (int k = 0; k < MAX_U; ++k) {
        const int u = ew + k * EPI_WARPS;
#pragma unroll
        for (int i = 0; i < 4; ++i) v[k][i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (u >= total_u) continue;
        const int g = u / nunits, c = (u % nunits) * UNIT_COLS;
        if (n0 + c >= p.N) continue;  // warp-uniform
        const int nrows = min(32, rows_per - g * 32);
        const int trow0 = (int)crank * rows_per + g * 32;
        for (int peer = 0; peer < p.ck; peer += 2) {   // ascending k ranges: fixed summation order
          // the loads of two peers are in flight before the first add (a DSMEM round trip each otherwise)
          float4 t[2][4];
#pragma unroll
          for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = min(lr + 8 * i, nrows - 1);  // rows past the block repeat the last one (not used)
              const uint32_t addr = stage_base + (uint32_t)(((trow0 + r) * (p.BN + STAGING_PAD) + c + c4) * 4);
              t[pp][i] = ld_dsmem_f4(addr, (uint32_t)(peer + pp));   // ck is even
            }
#pragma unroll
          for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[k][i].x += t[pp][i].x; v[k][i].y += t[pp][i].y; v[k][i].z += t[pp][i].z; v[k][i].w += t[pp][i].w;
            }
        }
      }
      __syncwarp();
      ptx::cluster_arrive_relaxed();   // "done reading the peers' staging buffers": nothing to publish
      if (threadIdx.x == 128) EGB_TRACE(16);
#pragma unroll
      for (int k = 0; k < MAX_U; ++k) {
        const int u = ew + k * EPI_WARPS;
        if (u >= total_u) continue;
        const int g = u / nunits, c = (u % nunits) * UNIT_COLS;
        const int col0 = n0 + c;
        if (col0 >= p.N) continue;  // warp-uniform
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[k][i].x *= p.alpha; v[k][i].y *= p.alpha; v[k][i].z *= p.alpha; v[k][i].w *= p.alpha;
        }
        if (!(have_pre && k == 0)) unit_loads(pre, m0 + (int)crank * rows_per + g * 32, min(32, rows_per - g * 32), col0);
        finish_unit(v[k], pre, m0 + (int)crank * rows_per + g * 32, col0);
      }
      __syncwarp();
      if (threadIdx.x == 128) EGB_TRACE(17);
      ptx::cluster_wait();  // peers may still be reading this CTA's staging buffer until here
    }
  }

  if (threadIdx.x == 128) EGB_TRACE(6);
  if (p.ck > 1 && warp < 4) {
    // non-epilogue warps take part in the two cluster barriers of the split-K reduction
    __syncwarp();
    ptx::cluster_sync_all();
    ptx::cluster_arrive_relaxed();
    ptx::cluster_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
  if (trace_slot && threadIdx.x == 0) {
    trace_sm[7] = (unsigned long long)clock64();
    trace_sm[8] = globaltimer_ns();
    for (int w = 0; w < 18; ++w) trace_slot[w] = trace_sm[w];
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

// Tile width and cluster split-K factor for problems that cannot fill the machine with 128 x 256 tiles.
// Cost model fitted to the device timeline (tools/gemm_trace.py, EGB_GEMM_TRACE), in microseconds:
//   * ~2.1 fixed: prologue, first TMA round trip, exit;
//   * main loop: shared-memory bandwidth sets the pace at these tile widths - per 64-deep k-block TMA
//     writes (128 + bn) * 256 B and the 12 SS-mode MMAs read 3 * (128 + bn) * 128 B, at 128 B/clk
//     (measured 0.45 / 0.49 / 0.62 us at bn = 32 / 64 / 128; one product instead of three: 0.32 us);
//   * epilogue: ~1.5 per 32 x 16 unit a warp finishes (8 warps; one L2 round trip + the fused stages);
//   * cluster split-K: k-blocks and epilogue rows are divided by ck, plus staging + two cluster barriers
//     (~0.7 + up to 2 us of skew on a full machine) and the DSMEM reduction, which moves only ~20 B/clk per SM - wide tiles
//     with deep splits (bn = 256, ck = 8: 13 us epilogues) lose to narrow tiles with ck = 2.
void choose_small_config(int M, int N, int K, int sm_count, bool b_mn, int max_ck, int* bn_out, int* ck_out) {
  const int tiles_m = (M + BM - 1) / BM;
  const int num_kb = (K + BK - 1) / BK;
  const int step = b_mn ? 64 : 32;
  double best = 1e30;
  *bn_out = choose_bn(M, N, sm_count, b_mn);
  *ck_out = 1;
  for (int bn = 256; bn >= step; bn -= step) {
    if (bn > step && bn - step >= N) continue;  // a narrower tile already covers N
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    const double t_kb = (128.0 + bn) * 640.0 / 128.0 / 1900.0;   // bytes through shared memory / (128 B/clk)
    const int units = ((N < bn ? N : bn) + UNIT_COLS - 1) / UNIT_COLS;  // 16-column units with data
    for (int ck = 1; ck <= max_ck; ck *= 2) {
      if (ck > 1 && tiles * ck > sm_count) break;   // clusters must be co-resident in one wave
      const int kbps = (num_kb + ck - 1) / ck;
      if (ck > 1 && (ck - 1) * kbps >= num_kb) break;  // an empty k range
      const double waves = (double)((tiles * ck + sm_count - 1) / sm_count);
      double cost = 2.1 + waves * kbps * t_kb;
      if (ck == 1) {
        cost += waves * 1.5 * ((units + 1) / 2);        // two warps share a lane quarter
      } else {
        const int groups = (BM / ck + 31) / 32;
        const int upw = (groups * units + EPI_WARPS - 1) / EPI_WARPS;
        if (upw > 4) continue;
        // DSMEM moves ~20 B/clk per SM: the (ck - 1) remote partial slices of this CTA's rows
        const double t_dsmem = 0.3 + (double)(BM / ck) * bn * 4.0 * (ck - 1) / 20.0 / 1900.0;
        // staging + barrier, reduction, fused stages, exit skew (grows with the share of the machine in use)
        cost += 0.7 + t_dsmem + upw * 1.5 + 2.0 * (double)(tiles * ck) / sm_count;
      }
      if (cost < best - 1e-9) {
        best = cost;
        *bn_out = bn;
        *ck_out = ck;
      }
    }
  }
}

}  // namespace

// What launch_gemm_bf16x3 does when the caller leaves the tile width open: `sms` SMs to plan for on a machine of
// `machine_sms`, cluster split-K factors up to max_ck.
void gemm_plan(int M, int N, int K, bool b_mn, int sms, int machine_sms, int max_ck, int* bn, int* ck) {
  *bn = choose_bn(M, N, sms, b_mn);
  *ck = 1;
  if (((M + BM - 1) / BM) * ((N + 255) / 256) * 2 <= machine_sms) choose_small_config(M, N, K, sms, b_mn, max_ck, bn, ck);
  const int num_kb = (K + BK - 1) / BK;
  while (*ck > 1 && (*ck - 1) * ((num_kb + *ck - 1) / *ck) >= num_kb) *ck /= 2;
}

void gemm_choose_config(int M, int N, int K, bool b_mn, int sm_count, int* bn, int* tiles) {
  (void)K;
  *bn = choose_bn(M, N, sm_count, b_mn);
  *tiles = ((M + BM - 1) / BM) * ((N + *bn - 1) / *bn);
}

// CTAs the launch of `a` will occupy when it may plan for `sm_budget` SMs (0 = all) of a machine of `machine_sms`:
// the planner's view of launch_gemm_bf16x3's dispatch (host only).
int gemm_planned_ctas(const GemmArgs& a, int sm_budget, int machine_sms) {
  if (a.M <= 0 || a.N <= 0) return 0;
  if (a.bn == 0 && gemm_2cta_eligible(a)) return machine_sms;
  const int sms = (sm_budget > 0 && sm_budget < machine_sms) ? sm_budget : machine_sms;
  const int tiles_m = (a.M + BM - 1) / BM;
  const bool small = tiles_m * ((a.N + 255) / 256) * 2 <= machine_sms;
  int bn = a.bn, ck = a.cluster_k > 0 ? a.cluster_k : 1;
  if (bn == 0) {
    int ck_auto = 1;
    if (small && gemm_lat_eligible(a)) {
      gemm_lat_plan(a.M, a.N, a.K, a.b_mn, sms, a.cluster_k == 1 ? 1 : (a.cluster_k > 1 ? a.cluster_k : 4), &bn, &ck_auto);
      if (a.cluster_k == 0) ck = ck_auto;
      if (ck > 1 && bn > 64) bn = 64;
    } else {
      gemm_plan(a.M, a.N, a.K, a.b_mn, sms, machine_sms, a.cluster_k == 1 ? 1 : 8, &bn, &ck_auto);
      if (a.cluster_k == 0) ck = ck_auto;
    }
  }
  const int units = tiles_m * ((a.N + bn - 1) / bn) * (ck > 1 ? ck : 1);
  return (ck > 1 || units < sms) ? units : sms;
}

void launch_gemm_bf16x3(Context& ctx, const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return;
  if (a.K <= 0) fail(EGB_ERR_GPU, "gemm: K must be positive");
  if (!ctx.encode_tiled) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled entry point not available");
  if (a.bn == 0 && gemm_2cta_eligible(a)) {
    launch_gemm_bf16x3_2cta(ctx, a, st);
    return;
  }
  {
    // problems that cannot fill the machine with 128 x 256 tiles are latency bound: gemm_lat.cu
    static const bool lat_always = getenv("EGB_GEMM_LAT_ALWAYS") != nullptr;
    const bool small = ((a.M + BM - 1) / BM) * ((a.N + 255) / 256) * 2 <= ctx.sm_count;
    if ((small || lat_always) && gemm_lat_eligible(a) && launch_gemm_lat(ctx, a, st)) return;
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
  p.trace = ctx.trace;
  p.trace_index = ctx.trace ? 1 + (int)(ctx.trace_next++ % (TRACE_SLOTS - 1)) : 0;
  const int sms = (a.sm_budget > 0 && a.sm_budget < ctx.sm_count) ? a.sm_budget : ctx.sm_count;
  p.BN = a.bn;
  int ck = a.cluster_k > 0 ? a.cluster_k : 1;
  static const bool no_cluster = getenv("EGB_GEMM_NO_CLUSTER_SPLITK") != nullptr;
  if (a.bn == 0) {
    int ck_auto = 1;
    gemm_plan(a.M, a.N, a.K, a.b_mn, sms, ctx.sm_count, a.cluster_k == 1 || no_cluster ? 1 : 8, &p.BN, &ck_auto);
    if (a.cluster_k == 0) ck = ck_auto;
  }
  if (p.BN % 32 != 0 || p.BN < 32 || p.BN > 256) fail(EGB_ERR_GPU, "gemm: invalid BN %d", p.BN);
  if (a.b_mn && p.BN % 64 != 0) fail(EGB_ERR_GPU, "gemm: BN must be a multiple of 64 for an MN-major B operand");
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + p.BN - 1) / p.BN;
  const int stage_bytes = 2 * A_PLANE_BYTES + 2 * p.BN * BK * 2;
  const int bar_bytes = BAR_BYTES + WSTAGE_BYTES;  // barriers + the epilogue warps' transposition buffers
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

  static const bool force_fused = getenv("EGB_GEMM_FUSED_ALWAYS") != nullptr;
  const int tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (a.K + BK - 1) / BK;
  int splits = 1;
  if (ck > 1) {
    while (ck > 1 && (ck - 1) * ((num_kb + ck - 1) / ck) >= num_kb) ck /= 2;  // every CTA needs k-blocks
    if ((size_t)BM * (p.BN + STAGING_PAD) * 4 > (size_t)stages * stage_bytes) ck = 1;
    if (ck != 1 && ck != 2 && ck != 4 && ck != 8) fail(EGB_ERR_GPU, "gemm: invalid cluster split-K factor %d", ck);
  }
  if (ck > 1) splits = ck;
  if (splits > num_kb) splits = num_kb;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.ck = ck > 1 ? p.splits : 1;
  const int units = tiles * p.splits;
  const bool fused = force_fused || (a.flags & (GEMM_BIAS | GEMM_SPLIT_OUT)) || a.epi != EPI_NONE || a.colsum || p.splits > 1;
  const int grid = (p.ck > 1 || units < sms) ? units : sms;
  // one instantiation per second-stage mode (compile-time epilogue, see the kernel's comment)
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, KParams);
  KernelFn fn = nullptr;
  if (!fused) {
    fn = gemm_bf16x3_kernel<false, EPI_NONE>;
  } else {
    switch (a.epi) {
      case EPI_NONE: fn = gemm_bf16x3_kernel<true, EPI_NONE>; break;
      case EPI_RELU: fn = gemm_bf16x3_kernel<true, EPI_RELU>; break;
      case EPI_LEAKY: fn = gemm_bf16x3_kernel<true, EPI_LEAKY>; break;
      case EPI_MASK_RELU: fn = gemm_bf16x3_kernel<true, EPI_MASK_RELU>; break;
      case EPI_MASK_LEAKY: fn = gemm_bf16x3_kernel<true, EPI_MASK_LEAKY>; break;
      case EPI_SGD: fn = gemm_bf16x3_kernel<true, EPI_SGD>; break;
      case EPI_SIGMOID: fn = gemm_bf16x3_kernel<true, EPI_SIGMOID>; break;
      case EPI_TANH: fn = gemm_bf16x3_kernel<true, EPI_TANH>; break;
      default: fail(EGB_ERR_GPU, "gemm: unknown second-stage mode %d", a.epi);
    }
  }
  if (first_use_on_device(ctx, (const void*)fn))
    EGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.ck > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = (unsigned)p.ck;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (ctx.pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    Launch l(ctx, KC_GEMM, st);
    EGB_CUDA(cudaLaunchKernelEx(&cfg, fn, tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid, p));
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
