// 2-CTA variant of the bf16x3 contraction for large problems (both operands K-major, plain epilogue).
//
// The single-CTA kernel (gemm_tcgen05.cu) is bound by L2 -> SM operand delivery on large problems:
// a 128 x 256 tile streams 96 KB of bf16 planes per k-block, which at the measured ~35 B/clk/SM
// (chip-wide L2 cap ~6300 B/clk) takes 1.4 us while its 12 MMAs need 0.83 us, so the tensor pipe idles
// ~40% of the time. Here two CTAs of a cluster form one `cta_group::2` MMA over a 256 x 256 tile: every
// CTA loads its own 128 rows of A and only HALF of the B tile (128 of the 256 columns); the tensor
// cores of both SMs read B from both shared memories. Per CTA and k-block that is 64 KB instead of
// 96 KB for the same MMA work.
//
// Tail wave ("stream-K" over the last wave only): 4096^3 has 256 pair tiles for 74 CTA pairs = 3.46 waves, so the
// fourth wave ran 46 % full at the length of a whole wave. When the remainder tiles fit the machine twice or more,
// each of them is cut along K into `split` units that run side by side in the last wave: the units with part > 0
// park their raw fp32 accumulators in a global workspace (coalesced, L2-resident: 256 KB per unit) and bump a
// flag; the part-0 unit of the tile waits for the flag, adds the parked partials to its own accumulator (fixed
// order) and runs the normal epilogue. No atomics on data, no zero fill; flags clean themselves for the next launch.
//
// Roles per CTA (256 threads): warp 0 TMA producer (completion bytes of both CTAs are signalled on the
// LEADER's full barrier), warp 1 MMA issuer (leader CTA only; commits multicast to both CTAs), warp 2
// TMEM allocator (`cta_group::2` alloc in both CTAs), warps 4-7 epilogue (each CTA drains its own 128
// accumulator rows and arrives on the leader's tmem_empty barrier).
#include <stdlib.h>

#include "egb_internal.hpp"
#include "ptx.cuh"

namespace egb {

namespace {

constexpr int BM2 = 128;                  // rows per CTA (256 per pair)
constexpr int BN2 = 256;                  // columns per pair (128 loaded by each CTA)
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int PLANE_BYTES = 128 * BK * 2; // 16 KiB: one 128-row plane tile
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;
constexpr int STAGES2 = 3;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int NUM_THREADS = 256;

struct K2Params {
  float* C;
  int ldc, M, N, K;
  int tiles_m, tiles_n;  // pair tiles: 256 x 256
  int accumulate;
  float alpha;
  // tail wave: tiles [0, full_units) are whole units, every later tile is `split` units of kbps k-blocks each
  int full_units, split, kbps, num_units;
  unsigned int* ws_flags;   // per split tile: {arrived, consumed}
  float4* ws_data;          // per split tile, part - 1, CTA of the pair: 128 x 256 fp32 in (chunk, j4, row) order
};

struct Unit {
  int tile, kb0, kb1, part, parts, slot;
};
__device__ __forceinline__ Unit get_unit(int u, const K2Params& p, int num_kb) {
  Unit r;
  if (u < p.full_units) {
    r.tile = u; r.kb0 = 0; r.kb1 = num_kb; r.part = 0; r.parts = 1; r.slot = 0;
  } else {
    const int v = u - p.full_units;
    r.slot = v / p.split;
    r.tile = p.full_units + r.slot;
    r.part = v % p.split;
    r.parts = p.split;
    r.kb0 = r.part * p.kbps;
    r.kb1 = min(num_kb, r.kb0 + p.kbps);
  }
  return r;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  // arrive on the barrier at the same offset in CTA 0 of the pair (peer bit cleared)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ptx::smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_2cta_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_mid,
                        const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_mid,
                        const K2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = (p.K + BK - 1) / BK;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a_hi);
    ptx::prefetch_tensormap(&tm_a_mid);
    ptx::prefetch_tensormap(&tm_b_hi);
    ptx::prefetch_tensormap(&tm_b_mid);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);    // leader producer's arrive.expect_tx (bytes of both CTAs)
      ptx::mbar_init(&empty_bar[s], 1);   // one multicast commit from the leader's MMA thread
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);   // multicast commit
      ptx::mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs (used on the leader only)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<2>(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / complete_tx
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    // convergent warp, elect.sync-guarded issue (no R2UR waterfall around UTMALDG, see gemm_tcgen05.cu)
    {
      pdl_wait();
      uint32_t it = 0;
      for (int u = pair; u < p.num_units; u += num_pairs) {
        const Unit un = get_unit(u, p, num_kb);
        const int tile = un.tile;
        const int m0 = (tile % p.tiles_m) * 2 * BM2 + (int)rank * BM2;   // this CTA's A rows
        const int n0 = (tile / p.tiles_m) * BN2 + (int)rank * (BN2 / 2); // this CTA's half of B
        for (int kb = un.kb0; kb < un.kb1; ++kb, ++it) {
          const int s = it % STAGES2;
          const uint32_t ph = (it / STAGES2) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1, 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int k0 = kb * BK;
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], 2u * STAGE_BYTES);
            ptx::tma_load_2d_cta2(st, &tm_a_hi, &full_bar[s], k0, m0);
            ptx::tma_load_2d_cta2(st + PLANE_BYTES, &tm_a_mid, &full_bar[s], k0, m0);
            ptx::tma_load_2d_cta2(st + 2 * PLANE_BYTES, &tm_b_hi, &full_bar[s], k0, n0);
            ptx::tma_load_2d_cta2(st + 3 * PLANE_BYTES, &tm_b_mid, &full_bar[s], k0, n0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      // instruction descriptor for the pair: M = 256, N = 256, bf16 x bf16 -> fp32, both K-major
      const uint32_t idesc = ptx::make_idesc_bf16_f32(2 * BM2, BN2);
      uint32_t it = 0, local_tile = 0;
      for (int u = pair; u < p.num_units; u += num_pairs, ++local_tile) {
        const Unit un = get_unit(u, p, num_kb);
        const uint32_t acc = local_tile & 1;
        const uint32_t use = local_tile >> 1;
        ptx::mbar_wait(&tmem_empty[acc], (use & 1) ^ 1, 2);   // both CTAs drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = un.kb0; kb < un.kb1; ++kb, ++it) {
          const int s = it % STAGES2;
          const uint32_t ph = (it / STAGES2) & 1;
          ptx::mbar_wait(&full_bar[s], ph, 3);   // the planes of both CTAs have landed
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t st = ptx::smem_u32(smem + s * STAGE_BYTES);
            const uint64_t a_hi = ptx::make_kmajor_sw128_desc(st);
            const uint64_t a_mid = ptx::make_kmajor_sw128_desc(st + PLANE_BYTES);
            const uint64_t b_hi = ptx::make_kmajor_sw128_desc(st + 2 * PLANE_BYTES);
            const uint64_t b_mid = ptx::make_kmajor_sw128_desc(st + 3 * PLANE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
              ptx::umma_f16<2>(d_tmem, a_mid + adv, b_hi + adv, idesc, (kb != un.kb0) || (k != 0));
              ptx::umma_f16<2>(d_tmem, a_hi + adv, b_mid + adv, idesc, 1);
              ptx::umma_f16<2>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
            ptx::umma_commit_cta2(&empty_bar[s], 3);                        // slot free in both CTAs
            if (kb == un.kb1 - 1) ptx::umma_commit_cta2(&tmem_full[acc], 3); // accumulators complete in both
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    uint32_t local_tile = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    for (int u = pair; u < p.num_units; u += num_pairs, ++local_tile) {
      const Unit un = get_unit(u, p, num_kb);
      const int tile = un.tile;
      const uint32_t acc = local_tile & 1;
      const uint32_t use = local_tile >> 1;
      const int m0 = (tile % p.tiles_m) * 2 * BM2 + (int)rank * BM2;
      const int n0 = (tile / p.tiles_m) * BN2;
      ptx::mbar_wait(&tmem_full[acc], use & 1, 4);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      const int row_local = q * 32 + lane;
      unsigned int* const flag = p.ws_flags + 2 * un.slot;
      if (un.parts > 1 && un.part > 0) {
        // ---- contributor: park the raw accumulator rows of this CTA (chunk, j4, row order: coalesced 512-byte runs)
        float4* const dst = p.ws_data + ((size_t)(un.slot * (un.parts - 1) + (un.part - 1)) * 2 + rank) * (BM2 * BN2 / 4);
        for (int c = 0; c < BN2; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + c, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            dst[((c >> 2) + (j >> 2)) * BM2 + row_local] = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        __threadfence();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          atomicAdd(flag, 1u);
          mbar_arrive_leader(&tmem_empty[acc]);
        }
        continue;
      }
      if (un.parts > 1) {
        // ---- finisher: the other parts of this tile (8 epilogue warps each: 4 per CTA of their pair) have parked
        if (lane == 0) {
          const unsigned int need = 8u * (unsigned)(un.parts - 1);
          const long long t0 = clock64();
          while (ld_acquire_gpu(flag) < need) {
            if (clock64() - t0 > 4000000000ll) {
              printf("egb gemm: split tile %d never received its partial sums\n", tile);
              __trap();
            }
          }
        }
        __syncwarp();
      }
      for (int c = 0; c < BN2; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c, r);
        ptx::tmem_ld_wait();
        if (un.parts > 1) {
          for (int part = 1; part < un.parts; ++part) {
            const float4* src = p.ws_data + ((size_t)(un.slot * (un.parts - 1) + (part - 1)) * 2 + rank) * (BM2 * BN2 / 4);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldcg(src + ((c >> 2) + (j >> 2)) * BM2 + row_local);
              r[j] = __float_as_uint(__uint_as_float(r[j]) + t.x);
              r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + t.y);
              r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + t.z);
              r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + t.w);
            }
          }
        }
        const int col0 = n0 + c;
        if (col0 >= p.N) break;
        if (row >= p.M) continue;
        const int ncols = min(32, p.N - col0);
        float* crow = p.C + (size_t)row * p.ldc + col0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (vec_ok && ncols == 32) {
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o = *reinterpret_cast<const float4*>(crow + j);
              v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) crow[j] = p.accumulate ? crow[j] + v[j] : v[j];
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_leader(&tmem_empty[acc]);
        if (un.parts > 1 && atomicAdd(flag + 1, 1u) == 7u) {   // last of the 8 finisher warps: reset for the next launch
          flag[0] = 0u;
          flag[1] = 0u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // nobody leaves while the peer may still signal into this CTA
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
  }
}

void encode_kmajor(Context& ctx, CUtensorMap* tm, const __nv_bfloat16* base, int rows, int K, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx.encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

}  // namespace

bool gemm_2cta_eligible(const GemmArgs& a) {
  static const bool disabled = getenv("EGB_GEMM_NO_2CTA") != nullptr;
  if (disabled) return false;
  if (a.a_mn || a.b_mn) return false;
  if (a.epi != EPI_NONE || a.colsum || a.bias || (a.flags & (GEMM_BIAS | GEMM_SPLIT_OUT)) || a.cluster_k > 1) return false;
  if ((a.lda & 7) || (a.ldb & 7)) return false;
  // worth it only when there are enough 256 x 256 pair tiles to fill the machine
  const long pair_tiles = (long)((a.M + 255) / 256) * ((a.N + 255) / 256);
  return a.M >= 512 && a.N >= 256 && a.K >= 256 && pair_tiles >= 64;
}

// workspace of the tail-wave split: flags (4 KB) + one parked 256 x 256 fp32 partial per remainder tile (at most
// pairs / 2 of them at split 2). Must be zero when first used (flags); the kernel leaves the flags zeroed.
size_t gemm_2cta_workspace_bytes(int sm_count) {
  const size_t pairs = (size_t)sm_count / 2;
  return 4096 + (pairs / 2) * (size_t)(2 * BM2) * BN2 * sizeof(float);
}

void launch_gemm_bf16x3_2cta(Context& ctx, const GemmArgs& a, cudaStream_t st) {
  K2Params p;
  p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K;
  p.tiles_m = (a.M + 2 * BM2 - 1) / (2 * BM2);
  p.tiles_n = (a.N + BN2 - 1) / BN2;
  p.accumulate = (a.flags & GEMM_ACCUMULATE) ? 1 : 0;
  p.alpha = a.alpha;
  CUtensorMap tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid;
  encode_kmajor(ctx, &tm_a_hi, a.a_hi, a.M, a.K, a.lda);
  encode_kmajor(ctx, &tm_a_mid, a.a_mid, a.M, a.K, a.lda);
  encode_kmajor(ctx, &tm_b_hi, a.b_hi, a.N, a.K, a.ldb);
  encode_kmajor(ctx, &tm_b_mid, a.b_mid, a.N, a.K, a.ldb);
  const size_t smem = 1024 + (size_t)STAGES2 * STAGE_BYTES + (2 * STAGES2 + 4) * 8 + 16;
  if (first_use_on_device(ctx, (const void*)gemm_bf16x3_2cta_kernel))
    EGB_CUDA(cudaFuncSetAttribute(gemm_bf16x3_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = p.tiles_m * p.tiles_n;
  int pairs = ctx.sm_count / 2;
  if (pairs > tiles) pairs = tiles;
  // tail wave: cut the remainder tiles along K when at least two units of each fit the last wave
  const int num_kb = (a.K + BK - 1) / BK;
  p.full_units = tiles;
  p.split = 1;
  p.kbps = num_kb;
  p.ws_flags = nullptr;
  p.ws_data = nullptr;
  static const bool no_tail = getenv("EGB_GEMM_NO_TAIL_SPLIT") != nullptr;
  const int rem = tiles % pairs;
  if (!no_tail && a.ws && rem > 0 && tiles > pairs && 2 * rem <= pairs && num_kb >= 16) {
    const int split = 2;   // (deeper splits would shorten the tail further but multiply the parked traffic)
    const size_t need = gemm_2cta_workspace_bytes(ctx.sm_count);
    if (a.ws_bytes >= need) {
      p.full_units = tiles - rem;
      p.split = split;
      p.kbps = (num_kb + split - 1) / split;
      p.ws_flags = reinterpret_cast<unsigned int*>(a.ws);
      p.ws_data = reinterpret_cast<float4*>(reinterpret_cast<char*>(a.ws) + 4096);
    }
  }
  p.num_units = p.full_units + (tiles - p.full_units) * p.split;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx.pdl ? 2 : 1;
  {
    Launch l(ctx, KC_GEMM, st);
    EGB_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16x3_2cta_kernel, tm_a_hi, tm_a_mid, tm_b_hi, tm_b_mid, p));
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
