// Latency-optimised variant of the bf16x3 contraction for the small problems of a training step
// (dense layers and their adjoints: exprgrad/layers/dnn.nim:19-27, passes.nim:519-549).
//
// The main loop is the one of gemm_tcgen05.cu (TMA producer warp -> mbarrier ring -> tcgen05.mma issuer ->
// fp32 accumulator in TMEM, three bf16 products per fp32 product - issued here as TWO MMAs per k-step, see the
// MMA issuer). What differs is everything behind the accumulator, which dominated those launches (device timeline,
// profiles/r02c_dense_step_trace_before.txt: main loop 0.4-4.0 us, epilogue 3.7-5.5 us per contraction):
//
//   * The epilogue works in the TMEM-native layout - thread t of a warp owns accumulator row t - on units of
//     32 rows x 32 columns. Every tensor the fused stages read or write moves as ONE 4 KB (fp32) or 2 KB (bf16)
//     TMA box per unit between global memory and a 128-byte / 64-byte SWIZZLED shared-memory tile, in which
//     row-per-thread 16-byte accesses are bank-conflict free. No transposition pass, no per-thread global
//     loads or stores, no edge code: TMA zero-fills reads and clips writes outside the tensor.
//   * The tiles a unit READS (old C for `+=`, the relu mask source, the parameter of the fused SGD update)
//     are prefetched by TMA while the main loop is still running, and updated in place, so a tile is also the
//     source of the TMA store of the result.
//   * Cluster split-K is PUSH based: the ck CTAs of a cluster each reduce a k range of one tile in their own
//     TMEM; CTA r finishes accumulator rows [r*128/ck, ...). A warp whose rows belong to another CTA drains
//     them into shared memory and sends them with one bulk copy (cp.async.bulk shared::cta -> shared::cluster)
//     that completes on an mbarrier of the owner - a one-way trip by the copy engine instead of two cluster
//     barriers and a round trip of per-thread DSMEM loads. Partial sums are added in k order: deterministic.
//   * The second stage is a compile-time parameter and the unit code exists once: the kernel is a quarter of the
//     size of the general one (4 100 vs 16 300 instructions), which matters for code that runs once per launch.
//
// Restrictions (everything else stays on gemm_tcgen05.cu): fp32 tensors with a 16-byte aligned base and a
// leading dimension that is a multiple of 4 (TMA), cluster split-K factors 2 and 4 with tiles of at most 64
// columns.
#include <stdlib.h>

#include <map>

#include "egb_internal.hpp"
#include "ptx.cuh"

namespace egb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_PLANE_BYTES = BM * BK * 2;
constexpr int MN_GROUP_BYTES = BK * 128;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int NUM_THREADS = 384;  // TMA, MMA, TMEM-allocator, idle warp + 8 epilogue warps
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int EPI_WARPS = 8;
constexpr int UNIT = 32;                       // an epilogue unit is 32 rows x 32 columns
constexpr int F32_TILE = UNIT * UNIT * 4;      // 4 KB, rows of 128 bytes, SWIZZLE_128B
constexpr int BF16_TILE = UNIT * UNIT * 2;     // 2 KB, rows of 64 bytes, SWIZZLE_64B
constexpr int EPI_DYNAMIC = 100;   // template argument: second stage chosen at run time
constexpr int BAR_BYTES = (2 * MAX_STAGES + 4 + 2 * EPI_WARPS) * 8 + 32 + TRACE_SLOT_WORDS * 8;

struct LParams {
  const float* bias;
  float* colsum;
  int epi;
  float epi_param;
  int M, N, K;
  int BN, stages;
  int tiles_m, tiles_n;
  int flags;
  float alpha;
  int a_mn, b_mn;
  int ck, kb_per_split;
  // per-warp epilogue region (offsets from the region start; -1 = absent)
  int epw;          // bytes per warp
  int off_a;        // fp32 tile: old C in / C out; staging of a pushed partial unit
  int off_b;        // fp32 tile: mask source or parameter in / second-stage value out
  int off_hi, off_mid, off_bias;
  int recv_off;     // cluster split-K: (finisher, peer) slots of 4 KB, from the start of the epilogue area
  int aux_is_h;     // the tile read into `b` is H (mask source); otherwise it is D itself (SGD parameter)
  int dry_run;      // epilogue warps pre-walk their code while the main loop runs
  int two_mma;      // bf16x3 as two MMAs per k-step: a_hi x [b_hi ; b_mid] (N = 2 BN) and a_mid x b_hi (see the MMA issuer)
  unsigned long long* trace;
  int trace_index;
};

#define EGB_TRACE(word)                                                        \
  do {                                                                         \
    if (trace_slot) trace_sm[word] = (unsigned long long)clock64();            \
  } while (0)

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
// bulk copy of this CTA's shared memory into the shared memory of a CTA of the cluster; completes (bytes) on
// an mbarrier of the DESTINATION CTA
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(ptx::smem_u32(src)), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// byte offset of the 16-byte chunk j (4 fp32) of row r inside a SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ int sw128(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }
// ... of the 16-byte chunk j (8 bf16) of row r inside a SWIZZLE_64B tile of 64-byte rows
__device__ __forceinline__ int sw64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Bounded mbarrier wait of this kernel: a protocol bug traps (never hangs the GPU) but does not print - the
// printf of ptx::mbar_wait costs ~40 instructions at each of the kernel's waits, and a kernel that runs once per
// step pays for every instruction it carries (measured: +1 800 instructions in this kernel = +6 us per dense step).
// Set EGB_LAT_WAIT_PRINTF at build time to get the tagged message back.
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int tag) {
#ifdef EGB_LAT_WAIT_PRINTF
  ptx::mbar_wait(bar, parity, tag);
#else
  (void)tag;
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
#endif
}

template <int kEpi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_lat_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_mid,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_mid,
                const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_d,
                const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_ohi,
                const __grid_constant__ CUtensorMap tm_omid, const LParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_plane_bytes = p.BN * BK * 2;
  const int stage_bytes = 2 * A_PLANE_BYTES + 2 * b_plane_bytes;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_units = num_tiles * p.ck;
  const uint32_t crank = p.ck > 1 ? ptx::cluster_ctarank() : 0u;

  uint8_t* const epi_area = smem + p.stages * stage_bytes;                       // 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_area + EPI_WARPS * p.epw + (p.ck > 1 ? (p.ck - 1) * (BM / p.ck) * 64 * 4 : 0));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_bar = tmem_empty + 2;            // per epilogue warp: its prefetched / loaded input tiles
  uint64_t* recv_bar = aux_bar + EPI_WARPS;      // per epilogue warp: pushed partial units of its peers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_bar + EPI_WARPS);
  unsigned long long** trace_slot_shp = reinterpret_cast<unsigned long long**>(tmem_slot + 4);
  unsigned long long* trace_sm = reinterpret_cast<unsigned long long*>(tmem_slot + 8);
  if (threadIdx.x == 0) {
    unsigned long long* slot = nullptr;
    if (p.trace && blockIdx.x == 0) {
      const unsigned long long i = (unsigned long long)p.trace_index;
      if (p.trace[0] < i) p.trace[0] = i;
      slot = p.trace + i * TRACE_SLOT_WORDS;
      for (int w = 0; w < TRACE_SLOT_WORDS; ++w) trace_sm[w] = 0;
      trace_sm[0] = globaltimer_ns(); trace_sm[1] = (unsigned long long)clock64();
      trace_sm[9] = p.M; trace_sm[10] = p.N; trace_sm[11] = p.K; trace_sm[12] = p.BN; trace_sm[13] = p.ck; trace_sm[14] = gridDim.x;
    }
    *trace_slot_shp = slot;
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a_hi);
    ptx::prefetch_tensormap(&tm_a_mid);
    ptx::prefetch_tensormap(&tm_b_hi);
    ptx::prefetch_tensormap(&tm_b_mid);
  }
  if (warp == 3 && lane == 0) {
    if (p.off_a >= 0) ptx::prefetch_tensormap(&tm_c);
    if (p.off_b >= 0) { ptx::prefetch_tensormap(&tm_d); ptx::prefetch_tensormap(&tm_h); }
    if (p.off_hi >= 0) { ptx::prefetch_tensormap(&tm_ohi); ptx::prefetch_tensormap(&tm_omid); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    for (int w = 0; w < EPI_WARPS; ++w) {
      ptx::mbar_init(&aux_bar[w], 1);
      ptx::mbar_init(&recv_bar[w], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  // the receive barriers of every CTA of the cluster exist before any peer may push into them
  if (p.ck > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* const trace_slot = *trace_slot_shp;
  if (threadIdx.x == 0) EGB_TRACE(2);

  if (warp == 0) {
    // ===================================================== TMA producer (convergent warp, elected lane issues)
    pdl_wait();
    if (lane == 0) EGB_TRACE(3);
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit / p.ck;
      const int kb0 = (unit % p.ck) * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        wait_bar(&empty_bar[s], ph ^ 1, 1);
        uint8_t* st = smem + s * stage_bytes;
        const int k0 = kb * BK;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          if (!p.a_mn) {
            ptx::tma_load_2d(st, &tm_a_hi, &full_bar[s], k0, m0);
            ptx::tma_load_2d(st + A_PLANE_BYTES, &tm_a_mid, &full_bar[s], k0, m0);
          } else {
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
    // ===================================================== MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(BM, p.BN, p.a_mn != 0, p.b_mn != 0);
    // Two-MMA form of the three-product scheme: the b_hi and b_mid tiles of a stage are adjacent in shared memory, so
    // ONE descriptor with N = 2 BN describes [b_hi ; b_mid] and a_hi is read once for both of its products:
    //   D[:, 0:BN] += a_hi b_hi + a_mid b_hi,   D[:, BN:2BN] += a_hi b_mid   (the epilogue adds the two halves)
    // - 18 -> 14 KB of operand reads per 16-deep k-step at BN = 64, and the main loop of these tiles is bound by
    // shared-memory bandwidth (0.49 -> 0.43 us per 64-deep k-block).
    const uint32_t idesc2 = ptx::make_idesc_bf16_f32(BM, 2 * p.BN, p.a_mn != 0, p.b_mn != 0);
    const uint64_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    const uint64_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
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
      const int kb0 = (unit % p.ck) * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      wait_bar(&tmem_empty[acc], (use & 1) ^ 1, 2);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        wait_bar(&full_bar[s], ph, 3);
        ptx::tc_fence_after();
        if (it == 0 && lane == 0) EGB_TRACE(4);
        const uint64_t so = (uint64_t)((uint32_t)(s * stage_bytes) >> 4);
        const uint64_t a_hi = a_desc0 + so, a_mid = a_hi + a_mid_off;
        const uint64_t b_hi = b_desc0 + so + b_off, b_mid = b_desc0 + so + b_mid_off;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t aa = a_step * k, ba = b_step * k;
            if (p.two_mma) {
              ptx::umma_f16<1>(d_tmem, a_hi + aa, b_hi + ba, idesc2, (kb != kb0) || (k != 0));
              ptx::umma_f16<1>(d_tmem, a_mid + aa, b_hi + ba, idesc, 1);
            } else {
              ptx::umma_f16<1>(d_tmem, a_mid + aa, b_hi + ba, idesc, (kb != kb0) || (k != 0));
              ptx::umma_f16<1>(d_tmem, a_hi + aa, b_mid + ba, idesc, 1);
              ptx::umma_f16<1>(d_tmem, a_hi + aa, b_hi + ba, idesc, 1);
            }
          }
          ptx::umma_commit(&empty_bar[s]);
          if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (8 warps)
    // Warp (q, eh): TMEM lane quarter q (accumulator rows q*32 .. +31), units at columns eh*32, eh*32 + 64, ...
    const int q = warp & 3;
    const int ew = warp - 4;
    const int eh = ew >> 2;
    uint8_t* const region = epi_area + ew * p.epw;
    uint8_t* const buf_a = region + max(p.off_a, 0);
    uint8_t* const buf_b = region + max(p.off_b, 0);
    uint8_t* const buf_hi = region + max(p.off_hi, 0);
    uint8_t* const buf_mid = region + max(p.off_mid, 0);
    float* const bias_s = reinterpret_cast<float*>(region + max(p.off_bias, 0));
    const bool has_bias = (p.flags & GEMM_BIAS) != 0;
    const bool load_c = (p.flags & GEMM_ACCUMULATE) != 0;
    // kEpi == EPI_DYNAMIC: the stage is read from the parameters (one kernel image for every launch of a step)
    const int epi = kEpi == EPI_DYNAMIC ? p.epi : kEpi;
    const bool need_aux = epi == EPI_MASK_RELU || epi == EPI_MASK_LEAKY || epi == EPI_SGD;
    const bool store_c = !(p.flags & GEMM_SKIP_C);
    const bool store_d = epi != EPI_NONE && !(p.flags & GEMM_SKIP_D);
    const bool planes = (p.flags & GEMM_SPLIT_OUT) != 0;
    const uint32_t aux_bytes = (load_c ? F32_TILE : 0) + (need_aux ? F32_TILE : 0);
    // cluster split-K roles: the CTA that finishes this warp's accumulator rows
    const int rows_per = BM / p.ck;
    const uint32_t owner = p.ck > 1 ? (uint32_t)(q * 32 / rows_per) : 0u;
    const bool finisher = owner == crank;
    const int fl = (q - (int)owner * (rows_per / 32)) * 2 + eh;         // index of this warp's unit among them
    uint32_t aux_phase = 0;

    // input tiles of a unit: issued by lane 0, land on aux_bar[ew]
    auto issue_inputs = [&](const int row0, const int col0) {
      if (aux_bytes && lane == 0) {
        ptx::mbar_arrive_expect_tx(&aux_bar[ew], aux_bytes);
        if (load_c) ptx::tma_load_2d(buf_a, &tm_c, &aux_bar[ew], col0, row0);
        if (need_aux) ptx::tma_load_2d(buf_b, p.aux_is_h ? &tm_h : &tm_d, &aux_bar[ew], col0, row0);
      }
      if (has_bias) {
        const int c = col0 + lane;
        bias_s[lane] = c < p.N ? __ldg(p.bias + c) : 0.0f;
      }
    };
    // v[0..31]: the raw accumulator row of this thread (columns col0 .. col0+31 of row row0 + lane)
    auto finish_unit = [&](uint32_t (&r)[32], const int row0, const int col0, const bool live) {
      if (aux_bytes && live) {
        wait_bar(&aux_bar[ew], aux_phase, 5);
        aux_phase ^= 1;
      }
      __syncwarp();  // bias_s written by the other lanes
      if (live && threadIdx.x == 128) EGB_TRACE(18);
      const bool row_ok = row0 + lane < p.M;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t hw[4] = {0u, 0u, 0u, 0u}, mw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * jj + h;
          float x[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = __fmul_rn(__uint_as_float(r[4 * j + e]), p.alpha);
          if (has_bias) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * j);
            x[0] = __fadd_rn(x[0], b.x); x[1] = __fadd_rn(x[1], b.y); x[2] = __fadd_rn(x[2], b.z); x[3] = __fadd_rn(x[3], b.w);
          }
          float4* const pa = reinterpret_cast<float4*>(buf_a + sw128(lane, j));
          float4* const pb = reinterpret_cast<float4*>(buf_b + sw128(lane, j));
          if (load_c) {
            const float4 o = *pa;
            x[0] = __fadd_rn(x[0], o.x); x[1] = __fadd_rn(x[1], o.y); x[2] = __fadd_rn(x[2], o.z); x[3] = __fadd_rn(x[3], o.w);
          }
          if (!row_ok) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = 0.0f;   // rows past M: clipped by the stores, must not reach the column sums
          }
          if (live && p.off_a >= 0 && (store_c || (epi == EPI_NONE && p.colsum))) *pa = make_float4(x[0], x[1], x[2], x[3]);
          float a2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          if (need_aux) {
            const float4 t = *pb;
            a2[0] = t.x; a2[1] = t.y; a2[2] = t.z; a2[3] = t.w;
          }
          if (epi == EPI_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (0.0f <= x[e]) ? x[e] : 0.0f;
          } else if (epi == EPI_LEAKY) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fmul_rn((0.0f <= x[e]) ? 1.0f : p.epi_param, x[e]);
          } else if (epi == EPI_MASK_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (0.0f <= a2[e]) ? x[e] : 0.0f;
          } else if (epi == EPI_MASK_LEAKY) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fmul_rn(x[e], (0.0f <= a2[e]) ? 1.0f : p.epi_param);
          } else if (kEpi != EPI_DYNAMIC && epi == EPI_SIGMOID) {   // (the transcendental stages keep their own images: ~1 000 instructions each)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(0.0f, x[e]))));
          } else if (kEpi != EPI_DYNAMIC && epi == EPI_TANH) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ep = expf(x[e]), en = expf(__fsub_rn(0.0f, x[e]));
              x[e] = __fdiv_rn(__fsub_rn(ep, en), __fadd_rn(ep, en));
            }
          } else if (epi == EPI_SGD) {
            // P += (0 - g) * rate   (base.nim:37-38; negate is `0 - x`, llvm.nim:333-336)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fadd_rn(a2[e], __fmul_rn(0.0f - x[e], p.epi_param));
          }
          if (epi != EPI_NONE) {
            if (!row_ok) {
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = 0.0f;
            }
            if (live && p.off_b >= 0) *pb = make_float4(x[0], x[1], x[2], x[3]);
          }
          // bf16 planes of the final value: hi = bf16(x), mid = bf16(x - hi)
          if (!planes) continue;
          hw[2 * h] = pack_bf16x2(x[0], x[1]);
          hw[2 * h + 1] = pack_bf16x2(x[2], x[3]);
          const float d0 = x[0] - __uint_as_float(hw[2 * h] << 16), d1 = x[1] - __uint_as_float(hw[2 * h] & 0xffff0000u);
          const float d2 = x[2] - __uint_as_float(hw[2 * h + 1] << 16), d3 = x[3] - __uint_as_float(hw[2 * h + 1] & 0xffff0000u);
          mw[2 * h] = pack_bf16x2(d0, d1);
          mw[2 * h + 1] = pack_bf16x2(d2, d3);
        }
        if (planes && live) {
          *reinterpret_cast<uint4*>(buf_hi + sw64(lane, jj)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(buf_mid + sw64(lane, jj)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        }
      }
      if (live && threadIdx.x == 128) EGB_TRACE(19);
      ptx::fence_proxy_async();   // the tiles are read by the TMA stores below
      __syncwarp();
      if (live && threadIdx.x == 128) EGB_TRACE(20);
      if (lane == 0 && live) {
        if (store_c) ptx::tma_store_2d(&tm_c, buf_a, col0, row0);
        if (store_d) ptx::tma_store_2d(&tm_d, buf_b, col0, row0);
        if (planes) {
          ptx::tma_store_2d(&tm_ohi, buf_hi, col0, row0);
          ptx::tma_store_2d(&tm_omid, buf_mid, col0, row0);
        }
        ptx::bulk_commit_group();
      }
      if (live && threadIdx.x == 128) EGB_TRACE(21);
      if (p.colsum) {
        // column `lane` of the final tile, rows top to bottom (a row's 32 words sit in 32 different banks)
        const uint8_t* fin = epi != EPI_NONE ? buf_b : buf_a;
        float s = 0.0f;
#pragma unroll 8
        for (int rr = 0; rr < UNIT; ++rr) s += *reinterpret_cast<const float*>(fin + sw128(rr, lane >> 2) + (lane & 3) * 4);
        if (live && col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, s);
      }
    };

    pdl_wait();
    // inputs of this warp's first unit: in flight while the main loop runs
    bool have_pre = false;
    if ((int)blockIdx.x < num_units && (p.ck == 1 || finisher)) {
      const int tile = blockIdx.x / p.ck;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      if (eh * UNIT < p.BN && n0 + eh * UNIT < p.N) {
        issue_inputs(m0 + q * 32, n0 + eh * UNIT);
        have_pre = true;
      }
    }
    if (p.ck > 1 && finisher && lane == 0 && (int)blockIdx.x < num_units) {
      const int tile = blockIdx.x / p.ck;
      const int n0 = (tile / p.tiles_m) * p.BN;
      if (eh * UNIT < p.BN && n0 + eh * UNIT < p.N) ptx::mbar_arrive_expect_tx(&recv_bar[ew], (uint32_t)((p.ck - 1) * F32_TILE));
    }
    // Dry pass: while the main loop of its first tile is running the warp walks through the code of that tile's
    // epilogue once with every side effect predicated off (it reads the idle second accumulator). That code runs
    // exactly once per launch otherwise - with a cold instruction cache, which the device timeline showed as the
    // largest part of the epilogue.
    uint32_t local_tile = 0;
    bool dry = p.dry_run != 0;
    bool cl_arrived = false;   // this warp has told the cluster that its part of the exchange is over
    for (int unit = blockIdx.x; unit < num_units;) {
      const bool live = !dry;
      const int tile = unit / p.ck;
      const uint32_t acc = live ? (local_tile & 1) : 1u;
      const uint32_t use = local_tile >> 1;
      const int m0 = (tile % p.tiles_m) * BM;
      const int n0 = (tile / p.tiles_m) * p.BN;
      if (live) {
        wait_bar(&tmem_full[acc], use & 1, 4);
        ptx::tc_fence_after();
        if (unit + (int)gridDim.x >= num_units) pdl_launch_dependents();
        if (local_tile == 0 && threadIdx.x == 128) EGB_TRACE(5);
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      for (int c = eh * UNIT; c < p.BN; c += 2 * UNIT) {
        const int col0 = n0 + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c, r);
        if (p.two_mma) {   // the a_hi x b_mid products sit BN columns further (see the MMA issuer)
          uint32_t r2[32];
          ptx::tmem_ld_32x32b_x32(t_row + p.BN + c, r2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
        } else {
          ptx::tmem_ld_wait();
        }
        if (live && local_tile == 0 && threadIdx.x == 128) EGB_TRACE(15);
        if (p.ck > 1 && !finisher) {
          // ---- push this partial unit to the CTA that finishes these rows
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(buf_a + sw128(lane, j)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && live) {
            const int peer_idx = (int)crank < (int)owner ? (int)crank : (int)crank - 1;
            uint8_t* slot = epi_area + p.recv_off + (fl * (p.ck - 1) + peer_idx) * F32_TILE;
            // the warp with the same (q, eh) in the owner CTA finishes these rows: same barrier index
            bulk_push(mapa(ptx::smem_u32(slot), owner), buf_a, F32_TILE, mapa(ptx::smem_u32(&recv_bar[ew]), owner));
          }
          if (live) {
            __syncwarp();
            ptx::cluster_arrive_relaxed();
            cl_arrived = true;
          }
          continue;
        }
        if (p.ck > 1) {
          // ---- add the pushed partial units in k order (this CTA's own partial takes its place in that order)
          if (live) {
            wait_bar(&recv_bar[ew], 0, 6);
            // every partial unit for these rows has arrived: the peers that sent them may exit (see the end of the kernel)
            __syncwarp();
            ptx::cluster_arrive_relaxed();
            cl_arrived = true;
          }
          const uint8_t* slots = epi_area + p.recv_off + fl * (p.ck - 1) * F32_TILE;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int peer = 0; peer < p.ck; ++peer) {
              float4 t;
              if (peer == (int)crank) {
                t = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
              } else {
                const int peer_idx = peer < (int)crank ? peer : peer - 1;
                t = *reinterpret_cast<const float4*>(slots + peer_idx * F32_TILE + sw128(lane, j));
              }
              if (peer == 0) { s[0] = t.x; s[1] = t.y; s[2] = t.z; s[3] = t.w; }
              else { s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w; }
            }
            r[4 * j] = __float_as_uint(s[0]); r[4 * j + 1] = __float_as_uint(s[1]);
            r[4 * j + 2] = __float_as_uint(s[2]); r[4 * j + 3] = __float_as_uint(s[3]);
          }
        }
        if (live && local_tile == 0 && threadIdx.x == 128) EGB_TRACE(16);
        if (live && !(have_pre && local_tile == 0 && c == eh * UNIT)) {
          // the previous unit's stores must have finished reading the tiles before they are refilled
          if (lane == 0) ptx::bulk_wait_group_read0();
          __syncwarp();
          issue_inputs(m0 + q * 32, col0);
        }
        finish_unit(r, m0 + q * 32, col0, live);
        if (live && local_tile == 0 && threadIdx.x == 128) EGB_TRACE(17);
      }
      if (!live) {
        dry = false;
        continue;   // now the same tile for real
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (p.ck > 1) break;  // exactly one unit per CTA
      unit += gridDim.x;
      ++local_tile;
    }
    // the stores read shared memory and must be complete before the grid counts as finished
    if (lane == 0) ptx::bulk_wait_group0();
    __syncwarp();
    if (p.ck > 1 && !cl_arrived) ptx::cluster_arrive_relaxed();
  }
  if (p.ck > 1 && warp < 4) {   // producer, MMA issuer, allocator, idle warp: nothing of theirs crosses the cluster
    __syncwarp();
    ptx::cluster_arrive_relaxed();
  }

  if (threadIdx.x == 128) EGB_TRACE(6);
  if (p.ck > 1) {
    // No CTA of the cluster exits while a peer's copy engine may still be reading its shared memory: a pushed unit
    // has left its source when the receiving warp has seen it land, and that is when that warp arrives (split-phase
    // barrier: every other warp arrives as soon as its role is over, nobody waits for anybody's stores).
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
    for (int w = 0; w < TRACE_SLOT_WORDS; ++w) trace_slot[w] = trace_sm[w];
  }
}

void encode_plane(Context& ctx, CUtensorMap* tm, const __nv_bfloat16* base, int mn, int K, int ld, int box_rows, bool mn_major) {
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
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

// a row-major [rows, cols] tensor seen through 32 x 32 element boxes (the epilogue unit)
void encode_unit_map(Context& ctx, CUtensorMap* tm, const void* base, int rows, int cols, int ld, bool bf16) {
  const int es = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)UNIT, (cuuint32_t)UNIT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx.encode_tiled(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled (epilogue unit map) failed with CUresult %d", (int)r);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Tile width and cluster split-K factor. Cost model in microseconds, from the device timeline (tools/gemm_trace.py):
// ~2 fixed (prologue, first TMA round trip, exit), main loop bound by shared-memory bandwidth (see
// choose_small_config in gemm_tcgen05.cu), ~1 for the first epilogue unit of a warp and ~0.4 for every further
// one, ~0.8 for the push of the partial units of a cluster split.
void gemm_lat_plan(int M, int N, int K, bool b_mn, int sms, int max_ck, int* bn_out, int* ck_out) {
  const int tiles_m = (M + BM - 1) / BM;
  const int num_kb = (K + BK - 1) / BK;
  const int step = b_mn ? 64 : 32;
  double best = 1e30;
  *bn_out = step;
  *ck_out = 1;
  for (int bn = 256; bn >= step; bn -= step) {
    if (bn > step && bn - step >= N) continue;  // a narrower tile already covers N
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    // bytes through shared memory per 64-deep k-block / (128 B/clk): TMA writes (128 + bn) * 256, the MMAs read
    // 3 * (128 + bn) * 128, or (2 * 128 + 3 * bn) * 128 in the two-MMA form (bn <= 128)
    const double t_kb = (bn <= 128 ? 65536.0 + 640.0 * bn : (128.0 + bn) * 640.0) / 128.0 / 1900.0;
    const int units_per_warp = (((N < bn ? N : bn) + UNIT - 1) / UNIT + 1) / 2;
    for (int ck = 1; ck <= max_ck && ck <= 4; ck *= 2) {
      if (ck > 1 && (bn > 64 || tiles * ck > sms)) break;
      const int kbps = (num_kb + ck - 1) / ck;
      if (ck > 1 && (ck - 1) * kbps >= num_kb) break;
      const double waves = (double)((tiles * ck + sms - 1) / sms);
      double cost = 2.0 + waves * (kbps * t_kb + 1.0 + 0.4 * (units_per_warp - 1));
      static const double split_penalty = getenv("EGB_LAT_SPLIT_PENALTY") ? atof(getenv("EGB_LAT_SPLIT_PENALTY")) : 0.8;
      if (ck > 1) cost += split_penalty;
      // a full machine finishes later than a half empty one (L2 and launch skew)
      cost += 0.5 * (double)(tiles * ck < sms ? tiles * ck : sms) / sms;
      if (cost < best - 1e-9) {
        best = cost;
        *bn_out = bn;
        *ck_out = ck;
      }
    }
  }
}

bool gemm_lat_eligible(const GemmArgs& a) {
  static const bool off = getenv("EGB_GEMM_NO_LAT") != nullptr;
  if (off) return false;
  if ((a.ldc & 3) != 0 || !aligned16(a.C)) return false;
  if (a.epi != EPI_NONE && !aligned16(a.D)) return false;
  if ((a.epi == EPI_MASK_RELU || a.epi == EPI_MASK_LEAKY) && !aligned16(a.H)) return false;
  if ((a.flags & GEMM_SPLIT_OUT) && ((a.ld_out & 7) != 0 || !aligned16(a.out_hi) || !aligned16(a.out_mid))) return false;
  if (a.cluster_k == 8) return false;
  if (a.cluster_k > 1 && a.bn > 64) return false;
  return true;
}

bool launch_gemm_lat(Context& ctx, const GemmArgs& a, cudaStream_t st) {
  LParams p;
  memset(&p, 0, sizeof(p));
  p.bias = a.bias; p.colsum = a.colsum; p.epi = a.epi; p.epi_param = a.epi_param;
  if (a.epi != EPI_NONE && !a.D) fail(EGB_ERR_GPU, "gemm: fused second stage needs an output tensor");
  if ((a.epi == EPI_MASK_RELU || a.epi == EPI_MASK_LEAKY) && !a.H) fail(EGB_ERR_GPU, "gemm: mask stage needs H");
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.flags = a.flags; p.alpha = a.alpha;
  // second stage written in place of C (raw entry point): two TMA stores to one address would race
  if (a.epi != EPI_NONE && a.D == a.C) p.flags |= GEMM_SKIP_C;
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
    gemm_lat_plan(a.M, a.N, a.K, a.b_mn, sms, a.cluster_k == 1 || no_cluster ? 1 : (a.cluster_k > 1 ? a.cluster_k : 4), &p.BN, &ck_auto);
    if (a.cluster_k == 0) ck = ck_auto;
    if (ck > 1 && p.BN > 64) p.BN = 64;
  }
  if (p.BN % 32 != 0 || p.BN < 32 || p.BN > 256) fail(EGB_ERR_GPU, "gemm: invalid BN %d", p.BN);
  if (a.b_mn && p.BN % 64 != 0) fail(EGB_ERR_GPU, "gemm: BN must be a multiple of 64 for an MN-major B operand");
  p.tiles_m = (a.M + BM - 1) / BM;
  p.tiles_n = (a.N + p.BN - 1) / p.BN;
  const int num_kb = (a.K + BK - 1) / BK;
  while (ck > 1 && (ck - 1) * ((num_kb + ck - 1) / ck) >= num_kb) ck /= 2;  // every CTA needs k-blocks
  if (ck != 1 && ck != 2 && ck != 4) fail(EGB_ERR_GPU, "gemm: invalid cluster split-K factor %d", ck);
  p.kb_per_split = (num_kb + ck - 1) / ck;
  p.ck = ck;
  // Dry pass of the epilogue code (see the kernel): it removes 0.4-1.3 us of cold-code stalls from the epilogue but
  // its shared-memory and TMEM reads compete with the main loop, which is bound by exactly those (a 2 us main loop
  // became 3 us: 70.4 vs 67.6 us per step when every launch did it). It pays where the main loop is long and the
  // epilogue is not preceded by a split-K exchange: unsplit launches with >= 6 k-blocks. EGB_GEMM_LAT_DRY=0/1 forces it.
  static const char* dry_env = getenv("EGB_GEMM_LAT_DRY");
  p.dry_run = dry_env ? (atoi(dry_env) != 0) : (ck == 1 && num_kb >= 6);

  // per-warp epilogue region
  const bool need_aux = a.epi == EPI_MASK_RELU || a.epi == EPI_MASK_LEAKY || a.epi == EPI_SGD;
  const bool load_c = (a.flags & GEMM_ACCUMULATE) != 0;
  const bool store_c = !(p.flags & GEMM_SKIP_C);
  const bool store_d = a.epi != EPI_NONE && !(p.flags & GEMM_SKIP_D);
  const bool has_a = load_c || store_c || (a.epi == EPI_NONE && a.colsum) || ck > 1;
  const bool has_b = need_aux || store_d || (a.epi != EPI_NONE && a.colsum);
  int off = 0;
  p.off_a = p.off_b = p.off_hi = p.off_mid = p.off_bias = -1;
  if (has_a) { p.off_a = off; off += F32_TILE; }
  if (has_b) { p.off_b = off; off += F32_TILE; }
  if (a.flags & GEMM_SPLIT_OUT) { p.off_hi = off; off += BF16_TILE; p.off_mid = off; off += BF16_TILE; }
  if (a.flags & GEMM_BIAS) { p.off_bias = off; off += 128; }
  p.epw = (off + 1023) & ~1023;
  if (p.epw == 0) p.epw = 1024;
  p.aux_is_h = (a.epi == EPI_MASK_RELU || a.epi == EPI_MASK_LEAKY) ? 1 : 0;
  p.recv_off = EPI_WARPS * p.epw;
  const int recv_bytes = ck > 1 ? (ck - 1) * (BM / ck) * 64 * 4 : 0;   // (finisher, peer) slots for 64-column tiles
  const int stage_bytes = 2 * A_PLANE_BYTES + 2 * p.BN * BK * 2;
  const int fixed = EPI_WARPS * p.epw + recv_bytes + BAR_BYTES;
  int stages = (SMEM_LIMIT - 1024 - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > num_kb + 1) stages = num_kb + 1;
  if (stages < 2) return false;   // wide tile + every epilogue buffer: the general kernel takes it
  p.stages = stages;
  static const bool three_mma = getenv("EGB_GEMM_LAT_THREE_MMA") != nullptr;
  p.two_mma = (!three_mma && 2 * p.BN <= ACC_COLS) ? 1 : 0;   // both column halves must fit one accumulator
  const size_t smem = 1024 + (size_t)stages * stage_bytes + fixed;

  CUtensorMap tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid, tm_c, tm_d, tm_h, tm_ohi, tm_omid;
  encode_plane(ctx, &tm_a_hi, a.a_hi, a.M, a.K, a.lda, BM, a.a_mn);
  encode_plane(ctx, &tm_a_mid, a.a_mid, a.M, a.K, a.lda, BM, a.a_mn);
  encode_plane(ctx, &tm_b_hi, a.b_hi, a.N, a.K, a.ldb, p.BN, a.b_mn);
  encode_plane(ctx, &tm_b_mid, a.b_mid, a.N, a.K, a.ldb, p.BN, a.b_mn);
  encode_unit_map(ctx, &tm_c, a.C, a.M, a.N, a.ldc, false);
  tm_d = tm_c; tm_h = tm_c; tm_ohi = tm_c; tm_omid = tm_c;
  if (a.epi != EPI_NONE) encode_unit_map(ctx, &tm_d, a.D, a.M, a.N, a.ldc, false);
  if (p.aux_is_h) encode_unit_map(ctx, &tm_h, a.H, a.M, a.N, a.ldc, false);
  if (a.flags & GEMM_SPLIT_OUT) {
    encode_unit_map(ctx, &tm_ohi, a.out_hi, a.M, a.N, a.ld_out, true);
    encode_unit_map(ctx, &tm_omid, a.out_mid, a.M, a.N, a.ld_out, true);
  }

  const int units = p.tiles_m * p.tiles_n * ck;
  const int grid = (ck > 1 || units < sms) ? units : sms;
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                           CUtensorMap, LParams);
  KernelFn fn = nullptr;
  // One kernel image per second stage. (EGB_GEMM_LAT_ONE_IMAGE: one image for every launch, stage read from the
  // parameters - measured 62.0 vs 59.0 us per dense step: the run-time stage tests sit on the epilogue's critical
  // path, and the instruction caches do not carry an image from one launch to the next anyway.)
  static const bool one_image = getenv("EGB_GEMM_LAT_ONE_IMAGE") != nullptr;
  if (one_image && a.epi != EPI_SIGMOID && a.epi != EPI_TANH) fn = gemm_lat_kernel<EPI_DYNAMIC>;
  else
  switch (a.epi) {
    case EPI_NONE: fn = gemm_lat_kernel<EPI_NONE>; break;
    case EPI_RELU: fn = gemm_lat_kernel<EPI_RELU>; break;
    case EPI_LEAKY: fn = gemm_lat_kernel<EPI_LEAKY>; break;
    case EPI_MASK_RELU: fn = gemm_lat_kernel<EPI_MASK_RELU>; break;
    case EPI_MASK_LEAKY: fn = gemm_lat_kernel<EPI_MASK_LEAKY>; break;
    case EPI_SGD: fn = gemm_lat_kernel<EPI_SGD>; break;
    case EPI_SIGMOID: fn = gemm_lat_kernel<EPI_SIGMOID>; break;
    case EPI_TANH: fn = gemm_lat_kernel<EPI_TANH>; break;
    default: fail(EGB_ERR_GPU, "gemm: unknown second-stage mode %d", a.epi);
  }
  if (first_use_on_device(ctx, (const void*)fn))
    EGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (ck > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)ck;
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
  {
    Launch l(ctx, KC_GEMM, st);
    EGB_CUDA(cudaLaunchKernelEx(&cfg, fn, tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid, tm_c, tm_d, tm_h, tm_ohi, tm_omid, p));
  }
  EGB_CUDA(cudaGetLastError());
  return true;
}

}  // namespace egb
