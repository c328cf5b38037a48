// conv2 forward on the 5th-gen tensor cores (sm_100a) as an implicit GEMM:
//
//   out[n,y,x,f] (+)= sum_{dy,dx,c} img[n,y+dy,x+dx,c] * w[f,dy,dx,c]          (exprgrad/layers/dnn.nim:45-49)
//
// seen as  OUT[p, f] = sum_k A[p, k] * W[f, k]  with p = (n,y,x) linearised, k = (dy,dx,c): M = 128 output
// pixels per tile, N = F filters, K = KH*KW*C <= 32 taps. The CUDA-core kernel (conv2.cu) is issue-bound
// at 0.32 of the HBM roofline (ncu: FMA pipe 43 %, issue 68 %); here the multiply-adds run as six
// tcgen05.mma per tile and the SM's instruction budget goes to building operand tiles and streaming the
// 3.2 GB output.
//
// fp32 accuracy on bf16 tensor cores (same bf16x3 scheme as gemm_tcgen05.cu): every operand is split
// into hi = bf16(x), mid = bf16(x - hi). One 128-byte shared-memory row per pixel holds
// [hi(32 taps) | mid(32 taps)] in the K-major SWIZZLE_128B layout (written by the producer threads, not
// by TMA: the im2col row of a pixel is KH runs of KW*C contiguous floats), so that
//     A' x B1,  B1 = [w_hi | w_hi]  (K = 64)   gives  hi*hi + mid*hi
//     A' x B2,  B2 = [w_mid]        (K = 32)   gives  hi*mid
// accumulate into one fp32 TMEM tile.
//
// Roles (320 threads, two CTAs per SM, persistent over pixel tiles):
//   warps 0-3  producers: one pixel each - KH*KW*C loads, split, 8 swizzled 16-byte stores
//   warp  4    MMA issuer (elect.sync-guarded, convergent warp)
//   warp  5    TMEM allocator
//   warps 6-9  epilogue: tcgen05.ld -> padded staging tile in shared memory -> ONE TMA tensor store per warp and
//              tile (cp.async.bulk.tensor: a warp's 32 pixels x F floats are a contiguous 8 KB run of the NHWC
//              output; the box is F + 4 floats wide like the padded tile and the tensor map clips the padding).
//              The LSU no longer re-reads the tile and issues no global stores: ncu had the kernel at 0.66 of
//              the HBM roofline with the producers' gathers, the staging traffic and the store loop all
//              competing for the same load/store pipe. (An accumulating convolution keeps the store loop.)
#include <stdlib.h>

#include "egb_internal.hpp"
#include "ptx.cuh"

namespace egb {
namespace {

constexpr int TILE_P = 128;            // pixels per tile = MMA M
constexpr int KPAD = 32;               // taps padded to 32 (two k-steps of 16)
constexpr int A_BYTES = TILE_P * 128;  // one operand tile: 128 rows of [hi(32) | mid(32)] bf16
constexpr int STAGES = 3;
constexpr int THREADS = 320;
constexpr int OUT_PAD = 4;

struct TcParams {
  const float* img;
  const float* w;
  float* out;
  int N, H, W, C, F, OH, OW;
  long total;   // output pixels
  int ntiles;
  int accumulate;
  int tma_store;   // epilogue stores through the output tensor map
};

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// (x0, x1) -> packed bf16x2 hi and mid words: hi = bf16(x), mid = bf16(x - hi). One cvt.rn.bf16x2.f32 per pair
// (the scalar conversions are the producers' most expensive instructions).
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& h, uint32_t& m) {
  const __nv_bfloat162 hp = __floats2bfloat162_rn(x0, x1);
  h = *reinterpret_cast<const uint32_t*>(&hp);
  const float r0 = x0 - __uint_as_float(h << 16);           // low half = x0's bf16 bits
  const float r1 = x1 - __uint_as_float(h & 0xffff0000u);
  const __nv_bfloat162 mp = __floats2bfloat162_rn(r0, r1);
  m = *reinterpret_cast<const uint32_t*>(&mp);
}

// byte offset of 16-byte chunk `c` of row `r` in a K-major SWIZZLE_128B tile (rows of 128 bytes)
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <int KH, int KWC>
__global__ void __launch_bounds__(THREADS, 2) conv2_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_out, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int F = p.F;
  const int b_bytes = F * 128;
  uint8_t* sB1 = smem;
  uint8_t* sB2 = smem + b_bytes;
  uint8_t* sA = smem + 2 * b_bytes;                              // STAGES operand tiles (b_bytes is a multiple of 1024)
  float* sOut = reinterpret_cast<float*>(sA + STAGES * A_BYTES);  // 4 warps x 32 x (F + OUT_PAD)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sOut + 4 * 32 * (F + OUT_PAD));
  uint64_t* a_empty = a_full + STAGES;
  uint64_t* tmem_full = a_empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const uint32_t tmem_cols = F <= 16 ? 32 : (F <= 32 ? 64 : (F <= 64 ? 128 : (F <= 128 ? 256 : 512)));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&a_full[s], TILE_P);  // every producer thread arrives
      ptx::mbar_init(&a_empty[s], 1);      // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 5) ptx::tmem_alloc<1>(tmem_slot, tmem_cols);
  pdl_wait();
  pdl_launch_dependents();
  // filter tiles, built once per CTA: B1 row f = [w_hi | w_hi], B2 row f = [w_mid | 0]
  constexpr int K = KH * KWC;
  for (int i = threadIdx.x; i < F * 8; i += THREADS) {
    const int f = i >> 3, c = i & 7;        // 16-byte chunk c of row f holds taps 8*(c&3) .. +7
    uint32_t h[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = 8 * (c & 3) + 2 * e;
      const float x0 = k0 < K ? __ldg(p.w + (size_t)f * K + k0) : 0.0f;
      const float x1 = k0 + 1 < K ? __ldg(p.w + (size_t)f * K + k0 + 1) : 0.0f;
      split_pair(x0, x1, h[e], m[e]);
    }
    *reinterpret_cast<uint4*>(sB1 + sw128(f, c)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sB2 + sw128(f, c)) = c < 4 ? make_uint4(m[0], m[1], m[2], m[3]) : make_uint4(0, 0, 0, 0);
  }
  ptx::fence_proxy_async();   // generic-proxy writes above are read by the tensor core (async proxy)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================================================== producers: one im2col row per thread
    const int row = threadIdx.x;
    // Software pipeline: the taps of the NEXT tile are in flight while this one is converted and stored
    // (ncu on the first version: 4.8 long-scoreboard stalls per issue - one exposed HBM/L2 round trip per
    // tile and thread, since the shared-memory stages are filled by the same threads one after another).
    auto load_taps = [&](int tile, float (&v)[K]) {
      const long pix = (long)tile * TILE_P + row;
      if (tile < p.ntiles && pix < p.total) {
        const int x = (int)(pix % p.OW);
        const long t = pix / p.OW;
        const int y = (int)(t % p.OH);
        const int n = (int)(t / p.OH);
        const float* base = p.img + (((size_t)n * p.H + y) * p.W + x) * p.C;
#pragma unroll
        for (int dy = 0; dy < KH; ++dy)
#pragma unroll
          for (int j = 0; j < KWC; ++j) v[dy * KWC + j] = __ldg(base + (size_t)dy * p.W * p.C + j);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = 0.0f;
      }
    };
    float v[K], nv[K];
    load_taps(blockIdx.x, nv);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
#pragma unroll
      for (int k = 0; k < K; ++k) v[k] = nv[k];
      load_taps(tile + gridDim.x, nv);
      ptx::mbar_wait_sleepy(&a_empty[s], ph ^ 1, 11);   // the MMAs that read this stage have retired
      uint8_t* a = sA + s * A_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) {   // chunk c: taps 8c .. 8c+7 (hi) and the same taps' mid in chunk c + 4
        uint32_t h[4], m[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = 8 * c + 2 * e;
          const float x0 = k0 < K ? v[k0 < K ? k0 : 0] : 0.0f;
          const float x1 = k0 + 1 < K ? v[k0 + 1 < K ? k0 + 1 : 0] : 0.0f;
          split_pair(x0, x1, h[e], m[e]);
        }
        *reinterpret_cast<uint4*>(a + sw128(row, c)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(a + sw128(row, c + 4)) = make_uint4(m[0], m[1], m[2], m[3]);
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&a_full[s]);
    }
  } else if (warp == 4) {
    // ===================================================== MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(TILE_P, F);
    const uint64_t a_desc0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA));
    const uint64_t b1_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB1));
    const uint64_t b2_desc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB2));
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      const uint32_t acc = it & 1;
      ptx::mbar_wait_sleepy(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 12);
      ptx::mbar_wait_sleepy(&a_full[s], ph, 13);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)F;
        const uint64_t a_desc = a_desc0 + (uint64_t)((uint32_t)(s * A_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // [hi | mid] x [w_hi | w_hi]; 16 taps = 32 bytes per step
          ptx::umma_f16<1>(d_tmem, a_desc + 2 * k, b1_desc + 2 * k, idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 2; ++k)   // hi x w_mid
          ptx::umma_f16<1>(d_tmem, a_desc + 2 * k, b2_desc + 2 * k, idesc, 1);
        ptx::umma_commit(&a_empty[s]);
        ptx::umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 6) {
    // ===================================================== epilogue
    const int q = warp & 3;   // TMEM lane quarter of this warp: pixels q*32 .. q*32+31 of the tile
    float* stage = sOut + q * 32 * (F + OUT_PAD);
    const int f4 = F >> 2;    // float4 per pixel
    const int f4_shift = 31 - __clz(f4);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      ptx::mbar_wait_sleepy(&tmem_full[acc], (it >> 1) & 1, 14);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)F;
      if (p.tma_store) {
        // the previous tile's tensor store must have finished reading the staging tile (it had a whole tile period)
        if (lane == 0) ptx::bulk_wait_group_read0();
        __syncwarp();
      }
      for (int c = 0; c < F; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c, r);
        ptx::tmem_ld_wait();
        float* mine = stage + lane * (F + OUT_PAD) + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(mine + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      ptx::tc_fence_before();
      if (p.tma_store) ptx::fence_proxy_async();   // this lane's staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);   // the accumulator is free while the tile streams out
      const long pix0 = (long)tile * TILE_P + q * 32;
      if (p.tma_store) {
        // rows past the last pixel and the 4 padding columns of every row lie outside the tensor: not written
        if (lane == 0 && pix0 < p.total) {
          ptx::tma_store_2d(&tm_out, stage, 0, (int)pix0);
          ptx::bulk_commit_group();
        }
        continue;
      }
      const long left = p.total - pix0;
      const int nrows = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
      float* dst = p.out + (size_t)pix0 * F;   // nrows * F contiguous floats
      const int total4 = nrows * f4;
      // F is a power of two: row / column of float4 i by shift and mask; four independent copies in flight
#pragma unroll 4
      for (int i = lane; i < total4; i += 32) {
        const int rr = i >> f4_shift, cc = (i & (f4 - 1)) << 2;
        float4 x = *reinterpret_cast<const float4*>(stage + rr * (F + OUT_PAD) + cc);
        if (p.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(dst + (size_t)i * 4);
          x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
        }
        *reinterpret_cast<float4*>(dst + (size_t)i * 4) = x;
      }
      __syncwarp();   // the staging rows are rewritten by the next tile
    }
    if (p.tma_store && lane == 0) ptx::bulk_wait_group0();   // every tensor store has landed before the CTA retires
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// d_images: dimg[n,y+dy,x+dx,c] += sum_f dout[n,y,x,f] * w[f,dy,dx,c]   (derive()d adjoint of conv2,
// exprgrad/passes.nim:519-549; the reference scatters, only n and c are independent loops).
// Per tile of 128 output pixels of one output row:  T[p, k] = sum_f dout[p, f] * w[f, k]  (M = 128,
// N = 32 taps, K = F = 64) on the tensor cores, then col2im: every input-row segment the tile touches
// (3 rows x 130 pixels x 3 channels) is gathered from T in shared memory (3 adds per element, no atomics
// inside the tile), accumulated in a three-row ring across the vertically adjacent tiles of a CTA and added
// to dimg with 16-byte REDs once per row - runs of different CTAs and the 2-pixel halo of neighbouring
// column tiles still overlap, so the image must be zeroed (or hold the value to accumulate onto) beforehand.
constexpr int DI_GATHER = 256;    // gather threads (8 warps; 4 measured 1.02 ms)
constexpr int DI_THREADS = 448 + DI_GATHER;   // 8 converter warps, MMA, loader/TMEM allocator, 4 drain + 8 gather warps
constexpr int DI_STAGES = 2;
constexpr int DI_TLD = 33;        // row stride of the T tile in shared memory (floats)
constexpr int RING_LD = 392;      // one input-row segment of a tile: (128 + 2) pixels x 3 channels, padded to 16 bytes

struct DimgParams {
  const float* dout;
  const float* w;
  float* dimg;
  int N, H, W, OH, OW;
  int tiles_per_row, ntiles;
  int vec4;   // image rows are 16-byte multiples and the image is 16-byte aligned: vector REDs
};

// fp32 -> (hi, mid) bf16 of 32 consecutive floats of one row, written as chunks c0..c0+3 of the row
__device__ __forceinline__ void split_store_32(const float4 (&x)[8], uint8_t* a_hi, uint8_t* a_mid, int row, int c0) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v[8] = {x[2 * i].x, x[2 * i].y, x[2 * i].z, x[2 * i].w, x[2 * i + 1].x, x[2 * i + 1].y, x[2 * i + 1].z, x[2 * i + 1].w};
    uint32_t h[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      split_pair(v[2 * e], v[2 * e + 1], h[e], m[e]);
    }
    *reinterpret_cast<uint4*>(a_hi + sw128(row, c0 + i)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_mid + sw128(row, c0 + i)) = make_uint4(m[0], m[1], m[2], m[3]);
  }
}

// Raw fp32 dout tile [128 pixels x 64 filters] (as the bulk copy delivered it) -> swizzled hi / mid bf16
// tiles. 256 threads; a warp converts 4 rows per step: lane l takes the 8 filters of chunk l % 8 of row
// l / 8 (32 contiguous bytes per lane, 256 per row: conflict-free reads and swizzled 16-byte writes).
// Rows >= valid_rows (tail tile / end of an output row) become zeros.
__device__ __forceinline__ void convert_dout_tile(const float* raw, uint8_t* a_hi, uint8_t* a_mid, int t, int valid_rows) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = j * 32 + (t >> 3), c = t & 7;
    float v[8];
    if (row < valid_rows) {
      const float4 x0 = *reinterpret_cast<const float4*>(raw + row * 64 + c * 8);
      const float4 x1 = *reinterpret_cast<const float4*>(raw + row * 64 + c * 8 + 4);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.0f;
    }
    uint32_t h[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      split_pair(v[2 * e], v[2 * e + 1], h[e], m[e]);
    }
    *reinterpret_cast<uint4*>(a_hi + sw128(row, c)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_mid + sw128(row, c)) = make_uint4(m[0], m[1], m[2], m[3]);
  }
}
constexpr int RAW_BYTES = TILE_P * 64 * 4;   // one raw fp32 dout tile
constexpr int RAW_STAGES = 3;

__global__ void __launch_bounds__(DI_THREADS, 1) conv2_dimg_tc_kernel(const DimgParams p) {
  constexpr int KH = 3, KW = 3, C = 3, F = 64, K = KH * KW * C, NT = 32;
  constexpr int SEG = (TILE_P + KW - 1) * C;   // floats of one input-row segment a tile touches (390)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* sWhi = smem;                       // [32 taps][64 f] bf16, K-major (K = f)
  uint8_t* sWmid = smem + NT * 128;
  uint8_t* sA = smem + 2 * NT * 128;          // DI_STAGES x (A_hi, A_mid), 16 KB each
  uint8_t* sRaw = sA + DI_STAGES * 2 * A_BYTES;   // RAW_STAGES raw fp32 dout tiles (bulk-copy ring)
  float* sT0 = reinterpret_cast<float*>(sRaw + RAW_STAGES * RAW_BYTES);   // two T tiles [128][DI_TLD]
  float* sRing = sT0 + 2 * TILE_P * DI_TLD + 3;                            // three input-row segments (16-byte aligned)
  sRing = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sRing) + 15) & ~uintptr_t(15));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sRing + 3 * RING_LD);
  a_full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(a_full) + 7) & ~uintptr_t(7));
  uint64_t* a_empty = a_full + DI_STAGES;
  uint64_t* tmem_full = a_empty + DI_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* raw_full = tmem_empty + 2;
  uint64_t* raw_empty = raw_full + RAW_STAGES;
  uint64_t* t_full = raw_empty + RAW_STAGES;   // T tile drained from TMEM -> gather warps
  uint64_t* t_free = t_full + 2;               // gather warps done with a T tile -> drain warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < DI_STAGES; ++s) {
      ptx::mbar_init(&a_full[s], 256);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int r = 0; r < RAW_STAGES; ++r) {
      ptx::mbar_init(&raw_full[r], 1);
      ptx::mbar_init(&raw_empty[r], 256);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);   // the four drain warps
      ptx::mbar_init(&t_full[a], 128);     // every drain thread
      ptx::mbar_init(&t_free[a], DI_GATHER);     // every gather thread
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<1>(tmem_slot, 64);
  pdl_wait();
  pdl_launch_dependents();
  // filter tiles: row k (tap), 64 f along K; rows >= 27 are zero
  for (int i = threadIdx.x; i < NT * 8; i += DI_THREADS) {
    const int k = i >> 3, c = i & 7;
    uint32_t h[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f0 = 8 * c + 2 * e;
      const float x0 = k < K ? __ldg(p.w + (size_t)f0 * K + k) : 0.0f;
      const float x1 = k < K ? __ldg(p.w + (size_t)(f0 + 1) * K + k) : 0.0f;
      split_pair(x0, x1, h[e], m[e]);
    }
    *reinterpret_cast<uint4*>(sWhi + sw128(k, c)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sWmid + sw128(k, c)) = make_uint4(m[0], m[1], m[2], m[3]);
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Tiles are numbered (n, x-tile, y) with y fastest and every CTA takes one contiguous range, so that the
  // tiles a CTA processes one after another are vertically adjacent (see the row ring in the epilogue).
  const int t0 = (int)((long)blockIdx.x * p.ntiles / gridDim.x), t1 = (int)((long)(blockIdx.x + 1) * p.ntiles / gridDim.x);
  auto decode = [&](int tile, int& x0, long& ny) {   // -> first output column, n * OH + y
    const int y = tile % p.OH, col = tile / p.OH;
    x0 = (col % p.tiles_per_row) * TILE_P;
    ny = (long)(col / p.tiles_per_row) * p.OH + y;
  };

  if (warp < 8) {
    // ===================================================== producers: convert the raw dout tile of the ring
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int s = it % DI_STAGES;
      const uint32_t ph = (it / DI_STAGES) & 1;
      const int rs = it % RAW_STAGES;
      const uint32_t rph = (it / RAW_STAGES) & 1;
      int x0;
      long ny;
      decode(tile, x0, ny);
      ptx::mbar_wait_sleepy(&a_empty[s], ph ^ 1, 21);
      ptx::mbar_wait_sleepy(&raw_full[rs], rph, 25);
      uint8_t* a_hi = sA + s * 2 * A_BYTES;
      convert_dout_tile(reinterpret_cast<const float*>(sRaw + rs * RAW_BYTES), a_hi, a_hi + A_BYTES, threadIdx.x,
                        min(TILE_P, p.OW - x0));
      ptx::mbar_arrive(&raw_empty[rs]);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&a_full[s]);
    }
  } else if (warp == 9) {
    // ===================================================== loader: bulk copies of the raw dout row segments
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int rs = it % RAW_STAGES;
      const uint32_t rph = (it / RAW_STAGES) & 1;
      ptx::mbar_wait_sleepy(&raw_empty[rs], rph ^ 1, 26);
      if (ptx::elect_one()) {
        int x0;
        long ny;
        decode(tile, x0, ny);
        const uint32_t bytes = (uint32_t)min(TILE_P, p.OW - x0) * F * 4;
        ptx::mbar_arrive_expect_tx(&raw_full[rs], bytes);
        ptx::bulk_load(sRaw + rs * RAW_BYTES, p.dout + ((size_t)ny * p.OW + x0) * F, bytes, &raw_full[rs]);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ===================================================== MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(TILE_P, NT);
    const uint64_t a_desc0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA));
    const uint64_t wh = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWhi));
    const uint64_t wm = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sWmid));
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int s = it % DI_STAGES;
      const uint32_t ph = (it / DI_STAGES) & 1;
      const uint32_t acc = it & 1;
      ptx::mbar_wait_sleepy(&tmem_empty[acc], ((it >> 1) & 1) ^ 1, 22);
      ptx::mbar_wait_sleepy(&a_full[s], ph, 23);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * NT;
        const uint64_t ah = a_desc0 + (uint64_t)((uint32_t)(s * 2 * A_BYTES) >> 4);
        const uint64_t am = ah + (uint64_t)(A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < F / 16; ++k) {
          ptx::umma_f16<1>(d_tmem, am + 2 * k, wh + 2 * k, idesc, k != 0);
          ptx::umma_f16<1>(d_tmem, ah + 2 * k, wm + 2 * k, idesc, 1);
          ptx::umma_f16<1>(d_tmem, ah + 2 * k, wh + 2 * k, idesc, 1);
        }
        ptx::umma_commit(&a_empty[s]);
        ptx::umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 10) {
    // ===================================================== epilogue: T -> shared memory -> col2im -> row ring -> RED
    // Two groups of four warps, decoupled by a double-buffered T tile: the DRAIN warps (one per TMEM lane
    // quarter) move an accumulator into shared memory and free it for the next MMAs; the GATHER warps turn it
    // into input-row contributions. The col2im contributions of a tile go to three input rows (y + dy);
    // consecutive tiles of this CTA are vertically adjacent, so the rows are accumulated in a three-slot ring
    // in shared memory and an input row is added to the image once (one 16-byte RED per four elements) when
    // the last tile that touches it inside this CTA's run has been processed - instead of three REDs per
    // element. Partial rows at the ends of a run simply meet the neighbouring CTA's part in memory.
    if (warp < 14) {
      // ---------------------------------------------------- drain warps (10..13 cover the four lane quarters)
      const int q = warp & 3;
      const int trow = q * 32 + lane;
      uint32_t it = 0;
      for (int tile = t0; tile < t1; ++tile, ++it) {
        const uint32_t acc = it & 1, buf = it & 1;
        int x0;
        long ny;
        decode(tile, x0, ny);
        ptx::mbar_wait_sleepy(&tmem_full[acc], (it >> 1) & 1, 24);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * NT, r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);   // the accumulator is free again
        ptx::mbar_wait_sleepy(&t_free[buf], ((it >> 1) & 1) ^ 1, 27);   // the gather warps are done with this buffer
        float* sT = sT0 + buf * TILE_P * DI_TLD;
        const bool live = x0 + trow < p.OW;   // rows past the end of the output row carry zeros anyway
#pragma unroll
        for (int j = 0; j < K; ++j) sT[trow * DI_TLD + j] = live ? __uint_as_float(r[j]) : 0.0f;
        ptx::mbar_arrive(&t_full[buf]);
      }
    } else {
      // ---------------------------------------------------- gather warps (14..17)
      const int gt = threadIdx.x - 448;   // 0..DI_GATHER-1
      for (int i = gt; i < 3 * RING_LD; i += DI_GATHER) sRing[i] = 0.0f;
      asm volatile("bar.sync 2, %0;" ::"n"(DI_GATHER) : "memory");
      uint32_t it = 0;
      for (int tile = t0; tile < t1; ++tile, ++it) {
        const uint32_t buf = it & 1;
        int x0;
        long ny;
        decode(tile, x0, ny);
        const int y = (int)(ny % p.OH);
        const long n = ny / p.OH;
        const float* sT = sT0 + buf * TILE_P * DI_TLD;
        ptx::mbar_wait_sleepy(&t_full[buf], (it >> 1) & 1, 28);
        constexpr int XS = TILE_P + KW - 1;   // input pixels of a row segment
        for (int idx = gt; idx < KH * XS; idx += DI_GATHER) {   // one (dy, input pixel) per thread: 9 taps -> 3 channels
          const int dy = idx / XS, X = idx - dy * XS;
          float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll
          for (int dx = 0; dx < KW; ++dx) {
            const int xx = X - dx;
            if (xx >= 0 && xx < TILE_P) {
              const float* t = sT + xx * DI_TLD + (dy * KW + dx) * C;
              s0 += t[0]; s1 += t[1]; s2 += t[2];
            }
          }
          float* ring = sRing + ((y + dy) % 3) * RING_LD + X * C;
          ring[0] += s0; ring[1] += s1; ring[2] += s2;
        }
        ptx::mbar_arrive(&t_free[buf]);                  // this thread no longer reads the T tile
        asm volatile("bar.sync 2, %0;" ::"n"(DI_GATHER) : "memory");   // the ring holds this tile
        // input row y is complete as far as this run goes; at the end of a column / of the run also y+1, y+2
        const int nflush = (y == p.OH - 1 || tile == t1 - 1) ? 3 : 1;
        const int seg_len = min(SEG, (p.W - x0) * C);   // the segment ends with the image row
        constexpr int Q = (SEG + 3) / 4;
        for (int idx = gt; idx < nflush * Q; idx += DI_GATHER) {
          const int fr = idx / Q, j = (idx - fr * Q) * 4;
          float* ring = sRing + ((y + fr) % 3) * RING_LD + j;
          const float4 val = *reinterpret_cast<const float4*>(ring);
          *reinterpret_cast<float4*>(ring) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (j >= seg_len) continue;
          float* dst = p.dimg + (((size_t)n * p.H + y + fr) * p.W + x0) * C + j;
          if (p.vec4 && j + 4 <= seg_len) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w)
                         : "memory");
          } else {
            const float e[4] = {val.x, val.y, val.z, val.w};
            for (int k2 = 0; k2 < 4 && j + k2 < seg_len; ++k2) atomicAdd(dst + k2, e[k2]);
          }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(DI_GATHER) : "memory");   // ring slots are reused by the next tile
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 64);
  }
}

// ------------------------------------------------------------------------------------------------------
// d_filters: dw[f,dy,dx,c] += sum_{n,y,x} dout[n,y,x,f] * img[n,y+dy,x+dx,c]   (derive()d adjoint of conv2;
// a reduction over all 12.6 M output pixels). As one MMA shape per 16 pixels:
//     D[128 x 64] += [dout_hi ; dout_mid]^T (M = 2 x 64 filters, MN-major)  x  [a_hi | a_mid] (N = 2 x 32 taps, MN-major)
// i.e. both operands are used exactly as the other two kernels build them (one 128-byte row per pixel),
// only described to the tensor core as MN-major with the pixel as K. Rows 0-63 of D hold hi*hi | hi*mid,
// rows 64-127 mid*hi | (mid*mid, unused): dw[f,k] = D[f,k] + D[f,32+k] + D[64+f,k].
// The tensor core's fp32 accumulation truncates, so D is drained into registers every DW_FLUSH tiles
// (512 pixels) and summed there with rounded adds; the CTAs meet in dw through atomics at the very end.
constexpr int DW_THREADS = 448;   // 8 producer warps, MMA, TMEM allocator, 4 epilogue warps
constexpr int DW_STAGES = 2;
constexpr int DW_RAW_STAGES = 3;   // raw dout tiles in flight (four - 128 KB per SM - measured the same 0.807 ms: not bound by bytes in flight)
constexpr int DW_FLUSH = 4;

struct DwParams {
  const float* img;
  const float* dout;
  float* dw;
  int N, H, W, OH, OW;
  long total;
  int ntiles;
};

__global__ void __launch_bounds__(DW_THREADS, 1) conv2_dw_tc_kernel(const DwParams p) {
  constexpr int KH = 3, KWC = 9, C = 3, F = 64, K = KH * KWC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* sS = smem;   // DW_STAGES x (dout_hi, dout_mid, im2col), 16 KB each
  uint8_t* sRaw = sS + DW_STAGES * 3 * A_BYTES;   // DW_RAW_STAGES raw fp32 dout tiles (bulk-copy ring)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sRaw + DW_RAW_STAGES * RAW_BYTES);
  uint64_t* a_empty = a_full + DW_STAGES;
  uint64_t* tmem_full = a_empty + DW_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* raw_full = tmem_empty + 2;
  uint64_t* raw_empty = raw_full + DW_RAW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + DW_RAW_STAGES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      ptx::mbar_init(&a_full[s], 256);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int r = 0; r < DW_RAW_STAGES; ++r) {
      ptx::mbar_init(&raw_full[r], 1);      // the loader's arrive.expect_tx
      ptx::mbar_init(&raw_empty[r], 256);   // every converter thread
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<1>(tmem_slot, 128);
  pdl_wait();
  pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) ++my_tiles;
  const int my_groups = (my_tiles + DW_FLUSH - 1) / DW_FLUSH;

  if (warp < 8) {
    // ===================================================== producers: convert the raw dout tile + half an im2col row
    // (dout arrives through the loader warp's bulk-copy ring, DW_RAW_STAGES tiles ahead: with per-thread
    // global loads ncu showed 8.7 long-scoreboard stalls per issue and issue slots 23 % busy)
    const int row = threadIdx.x & 127, half = threadIdx.x >> 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % DW_STAGES;
      const uint32_t ph = (it / DW_STAGES) & 1;
      const int rs = it % DW_RAW_STAGES;
      const uint32_t rph = (it / DW_RAW_STAGES) & 1;
      const long pix = (long)tile * TILE_P + row;
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.0f;
      if (pix < p.total) {
        const int xo = (int)(pix % p.OW);
        const long t = pix / p.OW;
        const int y = (int)(t % p.OH);
        const long n = t / p.OH;
        const float* base = p.img + (((size_t)n * p.H + y) * p.W + xo) * C;
        if (half == 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = __ldg(base + (size_t)(k / KWC) * p.W * C + (k % KWC));
        } else {
#pragma unroll
          for (int k = 16; k < K; ++k) v[k - 16] = __ldg(base + (size_t)(k / KWC) * p.W * C + (k % KWC));
        }
      }
      ptx::mbar_wait_sleepy(&a_empty[s], ph ^ 1, 31);
      uint8_t* d_hi = sS + s * 3 * A_BYTES;
      ptx::mbar_wait_sleepy(&raw_full[rs], rph, 35);
      const long left = p.total - (long)tile * TILE_P;
      convert_dout_tile(reinterpret_cast<const float*>(sRaw + rs * RAW_BYTES), d_hi, d_hi + A_BYTES, threadIdx.x,
                        left >= TILE_P ? TILE_P : (int)left);
      ptx::mbar_arrive(&raw_empty[rs]);
      uint8_t* a = d_hi + 2 * A_BYTES;   // im2col row: chunks 0-3 hi, 4-7 mid; this thread owns taps 16*half .. +15
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t h[4], m[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = v[8 * c + 2 * e], x1 = v[8 * c + 2 * e + 1];
          split_pair(x0, x1, h[e], m[e]);
        }
        *reinterpret_cast<uint4*>(a + sw128(row, half * 2 + c)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(a + sw128(row, 4 + half * 2 + c)) = make_uint4(m[0], m[1], m[2], m[3]);
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&a_full[s]);
    }
  } else if (warp == 9) {
    // ===================================================== loader: bulk copies of the raw dout tiles
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int rs = it % DW_RAW_STAGES;
      const uint32_t rph = (it / DW_RAW_STAGES) & 1;
      ptx::mbar_wait_sleepy(&raw_empty[rs], rph ^ 1, 36);
      if (ptx::elect_one()) {
        const long left = p.total - (long)tile * TILE_P;
        const uint32_t bytes = (uint32_t)(left >= TILE_P ? TILE_P : left) * 64 * 4;
        ptx::mbar_arrive_expect_tx(&raw_full[rs], bytes);
        ptx::bulk_load(sRaw + rs * RAW_BYTES, p.dout + (size_t)tile * TILE_P * 64, bytes, &raw_full[rs]);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ===================================================== MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, true, true);
    // A: two groups of 64 M-elements (dout_hi tile, dout_mid tile), A_BYTES apart; B: one group of 64 N-elements
    const uint64_t a_desc0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(sS), A_BYTES);
    const uint64_t b_desc0 = ptx::make_mnmajor_sw128_desc(ptx::smem_u32(sS + 2 * A_BYTES), A_BYTES);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % DW_STAGES;
      const uint32_t ph = (it / DW_STAGES) & 1;
      const uint32_t group = it / DW_FLUSH, in_group = it % DW_FLUSH;
      const uint32_t acc = group & 1;
      if (in_group == 0) ptx::mbar_wait_sleepy(&tmem_empty[acc], ((group >> 1) & 1) ^ 1, 32);
      ptx::mbar_wait_sleepy(&a_full[s], ph, 33);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * 64;
        const uint64_t so = (uint64_t)((uint32_t)(s * 3 * A_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < TILE_P / 16; ++k)   // 16 pixels = 16 rows of 128 bytes per step
          ptx::umma_f16<1>(d_tmem, a_desc0 + so + 128 * k, b_desc0 + so + 128 * k, idesc, (in_group | k) != 0);
        ptx::umma_commit(&a_empty[s]);
        if (in_group == DW_FLUSH - 1 || (int)it == my_tiles - 1) ptx::umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 10) {
    // ===================================================== epilogue: drain D every DW_FLUSH tiles, sum in registers
    const int q = warp & 3;
    const int drow = q * 32 + lane;   // row of D: filter drow (hi) or drow - 64 (mid)
    float sum[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) sum[j] = 0.0f;
    for (int g = 0; g < my_groups; ++g) {
      const uint32_t acc = g & 1;
      ptx::mbar_wait_sleepy(&tmem_full[acc], (g >> 1) & 1, 34);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[c + j] += __uint_as_float(r[j]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
    const int f = drow & 63;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float val = drow < 64 ? sum[k] + sum[32 + k] : sum[k];
      atomicAdd(p.dw + (size_t)f * K + k, val);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 128);
  }
}

size_t tc_smem_bytes(int F) {
  return 1024 + 2 * (size_t)F * 128 + (size_t)STAGES * A_BYTES + (size_t)4 * 32 * (F + OUT_PAD) * 4 + (2 * STAGES + 4) * 8 + 16;
}

}  // namespace

bool conv2_fwd_tc_supported(const float* out, int C, int F, int KH, int KW) {
  static const bool disabled = getenv("EGB_CONV_NO_TC") != nullptr;
  if (disabled) return false;
  // taps fit one 32-wide row, F is a legal MMA N with whole 32-column TMEM chunks and 1024-byte filter tiles
  return KH == 3 && KW == 3 && (C == 3 || C == 1) && (F == 32 || F == 64 || F == 128) &&
         (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

bool conv2_dimg_tc_supported(const float* dout, int C, int F, int KH, int KW) {
  static const bool disabled = getenv("EGB_CONV_NO_TC") != nullptr;
  return !disabled && KH == 3 && KW == 3 && C == 3 && F == 64 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0;
}

void launch_conv2_dimg_tc(Context& ctx, const float* dout, const float* w, float* dimg, int N, int H, int W, int C, int F,
                          int KH, int KW, bool accumulate, cudaStream_t st) {
  (void)C; (void)F;
  DimgParams p;
  p.dout = dout; p.w = w; p.dimg = dimg;
  p.N = N; p.H = H; p.W = W;
  p.OH = H - KH + 1; p.OW = W - KW + 1;
  if (N <= 0 || p.OH <= 0 || p.OW <= 0) return;
  p.tiles_per_row = (p.OW + TILE_P - 1) / TILE_P;
  const long tiles = (long)N * p.OH * p.tiles_per_row;
  if (tiles > 0x7fffffffL) fail(EGB_ERR_GPU, "conv2 d_images: too many tiles");
  p.ntiles = (int)tiles;
  p.vec4 = ((W * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(dimg) & 15) == 0) ? 1 : 0;
  // the tiles overlap (rows by dy, columns by the halo) and meet through RED: start from zero unless the
  // caller accumulates onto existing data
  if (!accumulate) {
    Launch lz(ctx, KC_CONV_DIMG, st);   // counted (and timed) with the kernel it belongs to
    EGB_CUDA(cudaMemsetAsync(dimg, 0, (size_t)N * H * W * 3 * sizeof(float), st));
  }
  const size_t smem = 1024 + 2 * 32 * 128 + (size_t)DI_STAGES * 2 * A_BYTES + (size_t)RAW_STAGES * RAW_BYTES +
                      (size_t)(2 * TILE_P * DI_TLD + 8 + 3 * RING_LD) * 4 + (2 * DI_STAGES + 8 + 2 * RAW_STAGES) * 8 + 32;
  int grid = ctx.sm_count;
  if (grid > p.ntiles) grid = p.ntiles;
  EGB_CUDA(cudaFuncSetAttribute(conv2_dimg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Launch l(ctx, KC_CONV_DIMG, st);
  launch_kernel(ctx, conv2_dimg_tc_kernel, dim3(grid), dim3(DI_THREADS), smem, st, p);
  EGB_CUDA(cudaGetLastError());
}

bool conv2_dw_tc_supported(const float* dout, int C, int F, int KH, int KW) {
  static const bool disabled = getenv("EGB_CONV_NO_TC") != nullptr;
  return !disabled && KH == 3 && KW == 3 && C == 3 && F == 64 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0;
}

// dw must be zero (or hold the value to accumulate onto): the CTAs add their partial sums atomically
void launch_conv2_dw_tc(Context& ctx, const float* img, const float* dout, float* dw, int N, int H, int W, int C, int F,
                        int KH, int KW, cudaStream_t st) {
  (void)C; (void)F;
  DwParams p;
  p.img = img; p.dout = dout; p.dw = dw;
  p.N = N; p.H = H; p.W = W;
  p.OH = H - KH + 1; p.OW = W - KW + 1;
  p.total = (long)N * p.OH * p.OW;
  if (p.total <= 0) return;
  p.ntiles = (int)((p.total + TILE_P - 1) / TILE_P);
  const size_t smem = 1024 + (size_t)DW_STAGES * 3 * A_BYTES + (size_t)DW_RAW_STAGES * RAW_BYTES +
                      (2 * DW_STAGES + 4 + 2 * DW_RAW_STAGES) * 8 + 32;
  int grid = ctx.sm_count;
  if (grid > p.ntiles) grid = p.ntiles;
  EGB_CUDA(cudaFuncSetAttribute(conv2_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Launch l(ctx, KC_CONV_DW, st);
  launch_kernel(ctx, conv2_dw_tc_kernel, dim3(grid), dim3(DW_THREADS), smem, st, p);
  EGB_CUDA(cudaGetLastError());
}

void launch_conv2_fwd_tc(Context& ctx, const float* img, const float* w, float* out, int N, int H, int W, int C, int F,
                         int KH, int KW, bool accumulate, cudaStream_t st) {
  TcParams p;
  p.img = img; p.w = w; p.out = out;
  p.N = N; p.H = H; p.W = W; p.C = C; p.F = F;
  p.OH = H - KH + 1; p.OW = W - KW + 1;
  p.total = (long)N * p.OH * p.OW;
  if (p.total <= 0) return;
  p.ntiles = (int)((p.total + TILE_P - 1) / TILE_P);
  p.accumulate = accumulate ? 1 : 0;
  // output as a 2-D tensor [pixels, F] with a box of 32 pixels x (F + OUT_PAD) floats = the padded staging tile
  CUtensorMap tm_out;
  memset(&tm_out, 0, sizeof(tm_out));
  static const bool no_tma_store = getenv("EGB_CONV_NO_TMA_STORE") != nullptr;
  p.tma_store = 0;
  if (!accumulate && !no_tma_store && ctx.encode_tiled && p.total < (long)1 << 31) {
    cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)p.total};
    cuuint64_t strides[1] = {(cuuint64_t)F * 4};
    cuuint32_t box[2] = {(cuuint32_t)(F + OUT_PAD), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx.encode_tiled(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)out, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p.tma_store = r == CUDA_SUCCESS ? 1 : 0;
  }
  const size_t smem = tc_smem_bytes(F);
  int grid = ctx.sm_count * 2;
  if (grid > p.ntiles) grid = p.ntiles;
  Launch l(ctx, KC_CONV, st);
  if (C == 3) {
    EGB_CUDA(cudaFuncSetAttribute(conv2_fwd_tc_kernel<3, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(ctx, conv2_fwd_tc_kernel<3, 9>, dim3(grid), dim3(THREADS), smem, st, tm_out, p);
  } else {
    EGB_CUDA(cudaFuncSetAttribute(conv2_fwd_tc_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(ctx, conv2_fwd_tc_kernel<3, 3>, dim3(grid), dim3(THREADS), smem, st, tm_out, p);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
