// Direct fp32 kernels for the reference's `conv2` layer (exprgrad/layers/dnn.nim:45-49: NHWC images,
// filters [filter, dy, dx, chan], valid padding, stride 1) and the two adjoint kernels `derive`
// generates for it (exprgrad/passes.nim:519-549):
//
//   forward    out[n,y,x,f]        += img[n,y+dy,x+dx,c] * w[f,dy,dx,c]
//   d_filters  dw[f,dy,dx,c]       += dout[n,y,x,f]      * img[n,y+dy,x+dx,c]      (reduction over n,y,x)
//   d_images   dimg[n,y+dy,x+dx,c] += dout[n,y,x,f]      * w[f,dy,dx,c]            (scatter in the reference,
//                                                                                   gathered per input pixel here)
//
// At BASELINE config 4 (256x224x224x3 images, 64 3x3x3 filters) the arithmetic intensity is ~13 flop/B:
// all three kernels are bound by streaming the 3.2 GB output / output-gradient tensor through HBM, and
// K = 27 is far too short for tensor-core tiles to pay off, so they run on the fp32 CUDA cores with
// register tiling (4 pixels x 8 filters per thread), shared-memory staging of the input rows and filter
// bank, and 128-bit coalesced loads/stores of the big tensor. Filters are processed in chunks of 64.
#include "egb_internal.hpp"

namespace egb {

namespace {

constexpr int FC = 64;        // filters per chunk
constexpr int TX_BWD = 64;    // pixels per block (both adjoint kernels)
constexpr int MAX_CC = 4;     // channels per pass

struct ConvDims {
  int N, H, W, C, F, KH, KW, OH, OW;
};

// ------------------------------------------------------------------ forward
// block: 8 filter groups (filters g*4..g*4+3 and 32+g*4..32+g*4+3) x TQ pixel groups (PXT consecutive x each), TQ <= 32 chosen by
// the host so that the strips divide the output row evenly. A block walks ROWS_FWD consecutive output
// rows of its strip: the filter bank is staged once, the KH input rows live in a ring buffer so that
// every new output row stages only one new input row.
constexpr int ROWS_FWD = 8;
constexpr int PXT = 4;          // output pixels per thread (forward; 8 was measured slower: 1.88 vs 1.63 ms)

template <int KW_T>
__global__ void __launch_bounds__(256) conv2_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                        float* __restrict__ out, ConvDims d, int accumulate, int tq_n) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int KW = KW_T > 0 ? KW_T : d.KW;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int tx = tid & 7, tq = tid >> 3;
  const int fchunks = (d.F + FC - 1) / FC;
  const int n = blockIdx.z / fchunks, f0 = (blockIdx.z % fchunks) * FC;
  const int y0 = blockIdx.y * ROWS_FWD, x0 = blockIdx.x * tq_n * PXT;
  const int y1 = min(y0 + ROWS_FWD, d.OH);
  const int in_w = tq_n * PXT + KW - 1;             // staged input pixels per row
  float* in_s = smem;                               // ring: [KH][in_w * cc]
  float* w_s = smem + d.KH * in_w * MAX_CC;         // [KH*KW*cc][FC]
  const int fb = f0 + tx * 4;                       // this thread's filters: fb..fb+3 and fb+32..fb+35
  const bool vec = (d.F & 3) == 0 && fb + 36 <= d.F;

  for (int c0 = 0; c0 < d.C; c0 += MAX_CC) {
    const int cc = min(MAX_CC, d.C - c0);
    const int row_len = in_w * cc;
    __syncthreads();
    for (int i = tid; i < d.KH * KW * cc * FC; i += nthreads) {
      const int f = i % FC, k = i / FC;
      const int ch = k % cc, dx = (k / cc) % KW, dy = k / (cc * KW);
      w_s[i] = (f0 + f) < d.F ? __ldg(w + (((size_t)(f0 + f) * d.KH + dy) * KW + dx) * d.C + c0 + ch) : 0.0f;
    }
    for (int y = y0; y < y1; ++y) {
      // stage input rows y .. y+KH-1 into ring slots (row % KH); after the first output row only the
      // newest input row is missing
      __syncthreads();
      for (int r = (y == y0 ? 0 : d.KH - 1); r < d.KH; ++r) {
        const int iy = y + r;
        float* dst = in_s + (iy % d.KH) * row_len;
        const float* src = img + (((size_t)n * d.H + iy) * d.W + x0) * d.C + c0;
        if (cc == d.C) {
          const int valid = min(in_w, d.W - x0) * cc;   // contiguous run
          for (int j = tid; j < row_len; j += nthreads) dst[j] = j < valid ? __ldg(src + j) : 0.0f;
        } else {
          for (int j = tid; j < row_len; j += nthreads) {
            const int px = j / cc, ch = j % cc;
            dst[j] = (x0 + px) < d.W ? __ldg(src + (size_t)px * d.C + ch) : 0.0f;
          }
        }
      }
      __syncthreads();
      float acc[PXT][8];
#pragma unroll
      for (int p = 0; p < PXT; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = 0.0f;
      if (tq < tq_n) {
        for (int dy = 0; dy < d.KH; ++dy) {
          const float* row = in_s + ((y + dy) % d.KH) * row_len + tq * PXT * cc;
          if (KW_T > 0) {
            for (int ch = 0; ch < cc; ++ch) {
              float iv[PXT + (KW_T > 0 ? KW_T : 1) - 1];
#pragma unroll
              for (int t = 0; t < PXT + KW_T - 1; ++t) iv[t] = row[t * cc + ch];
#pragma unroll
              for (int dx = 0; dx < KW_T; ++dx) {
                const float4* wp = reinterpret_cast<const float4*>(w_s + ((dy * KW_T + dx) * cc + ch) * FC) + tx;
                const float4 w0 = wp[0], w1 = wp[8];   // filters tx*4.. and 32+tx*4..: conflict-free 128-byte rows
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < PXT; ++p)
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(iv[p + dx], wv[j], acc[p][j]);
              }
            }
          } else {
            for (int dx = 0; dx < KW; ++dx)
              for (int ch = 0; ch < cc; ++ch) {
                const float4* wp = reinterpret_cast<const float4*>(w_s + ((dy * KW + dx) * cc + ch) * FC) + tx;
                const float4 w0 = wp[0], w1 = wp[8];
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < PXT; ++p) {
                  const float iv = row[(p + dx) * cc + ch];
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(iv, wv[j], acc[p][j]);
                }
              }
          }
        }
        // channel chunks after the first accumulate into what the first one stored
        const bool acc_out = accumulate || c0 > 0;
#pragma unroll
        for (int p = 0; p < PXT; ++p) {
          const int x = x0 + tq * PXT + p;
          if (x >= d.OW) continue;
          float* o = out + (((size_t)n * d.OH + y) * d.OW + x) * d.F + fb;
          if (vec) {
            float4 a = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
            float4 b = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
            if (acc_out) {
              const float4 oa = *reinterpret_cast<const float4*>(o), ob = *reinterpret_cast<const float4*>(o + 32);
              a.x += oa.x; a.y += oa.y; a.z += oa.z; a.w += oa.w;
              b.x += ob.x; b.y += ob.y; b.z += ob.z; b.w += ob.w;
            }
            *reinterpret_cast<float4*>(o) = a;
            *reinterpret_cast<float4*>(o + 32) = b;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int fo = (j & 3) + (j >> 2) * 32;
              if (fb + fo < d.F) o[fo] = acc_out ? o[fo] + acc[p][j] : acc[p][j];
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ d_filters
// Persistent blocks; a work item is (n, y, <=64-pixel chunk of the output row). 256 threads = 8 pixel
// sets x (8 filter groups: filters g*4..g*4+3 and 32+g*4..32+g*4+3) x (4 tap groups); a thread
// accumulates 8 filters x 8 filter taps (k = kc*32 + kg + 4*i) in registers over all its work items,
// the block reduces the 8 pixel sets through shared memory and flushes once with atomicAdd.
__global__ void __launch_bounds__(256) conv2_dw_kernel(const float* __restrict__ img, const float* __restrict__ dout,
                                                       float* __restrict__ dw, ConvDims d, int txw) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int K = d.KH * d.KW * d.C;
  const int kchunks = (K + 31) / 32, fchunks = (d.F + FC - 1) / FC;
  const int kc = blockIdx.y % kchunks, f0 = (blockIdx.y / kchunks) * FC;
  const int tid = threadIdx.x;
  const int ps = tid >> 5, r = tid & 31, fg = r & 7, kg = r >> 3;
  const int in_w = (txw + d.KW - 1) * d.C;          // staged floats per input row (all channels)
  (void)fchunks;
  int koff[8];
  bool kval[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kc * 32 + kg + 4 * i;
    kval[i] = k < K;
    const int kk = kval[i] ? k : 0;
    const int ch = kk % d.C, dx = (kk / d.C) % d.KW, dy = kk / (d.C * d.KW);
    koff[i] = dy * in_w + dx * d.C + ch;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  const int xchunks = (d.OW + txw - 1) / txw;
  const long items = (long)d.N * d.OH * xchunks;
  const int buf_floats = (TX_BWD * FC + d.KH * in_w + 3) & ~3;   // one staging buffer: dout tile + input rows
  const bool f_vec = (d.F & 3) == 0;
  // cp.async double buffering: the tile of work item i+1 streams in while item i is being reduced
  auto stage = [&](long it, int buf) {
    const int xc = (int)(it % xchunks);
    const int y = (int)((it / xchunks) % d.OH);
    const int n = (int)(it / ((long)xchunks * d.OH));
    const int x0 = xc * txw;
    float* dsm = smem + buf * buf_floats;
    float* ism = dsm + TX_BWD * FC;
    for (int i = tid; i < txw * (FC / 4); i += 256) {
      const int px = i / (FC / 4), f4 = (i % (FC / 4)) * 4;
      const int x = x0 + px;
      float* dst = dsm + px * FC + f4;
      const float* src = dout + (((size_t)n * d.OH + y) * d.OW + min(x, d.OW - 1)) * d.F + f0 + f4;
      if (f_vec && f0 + f4 + 4 <= d.F) {
        const unsigned bytes = x < d.OW ? 16u : 0u;   // zero-fill outside the row
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
                     "r"(bytes) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = (x < d.OW && f0 + f4 + j < d.F) ? __ldg(src + j) : 0.0f;
      }
    }
    for (int i = tid; i < d.KH * in_w; i += 256) {
      const int dy = i / in_w, rr = i % in_w;
      const int gx = x0 + rr / d.C;
      const float* src = img + (((size_t)n * d.H + y + dy) * d.W + x0) * d.C + (gx < d.W ? rr : 0);
      const unsigned bytes = gx < d.W ? 4u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(ism + i)), "l"(src),
                   "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if ((long)blockIdx.x < items) stage(blockIdx.x, 0);
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const long next = it + gridDim.x;
    if (next < items) {
      stage(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* dout_c = smem + buf * buf_floats;
    const float* img_c = dout_c + TX_BWD * FC;
    const int x0 = (int)(it % xchunks) * txw;
    const int npx = min(txw, d.OW - x0);
    for (int px = ps; px < npx; px += 8) {
      const float4 d0 = *reinterpret_cast<const float4*>(dout_c + px * FC + fg * 4);
      const float4 d1 = *reinterpret_cast<const float4*>(dout_c + px * FC + 32 + fg * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float iv = img_c[koff[i] + px * d.C];
        acc[i][0] = fmaf(iv, d0.x, acc[i][0]);
        acc[i][1] = fmaf(iv, d0.y, acc[i][1]);
        acc[i][2] = fmaf(iv, d0.z, acc[i][2]);
        acc[i][3] = fmaf(iv, d0.w, acc[i][3]);
        acc[i][4] = fmaf(iv, d1.x, acc[i][4]);
        acc[i][5] = fmaf(iv, d1.y, acc[i][5]);
        acc[i][6] = fmaf(iv, d1.z, acc[i][6]);
        acc[i][7] = fmaf(iv, d1.w, acc[i][7]);
      }
    }
    __syncthreads();   // everyone is done with this buffer before it is refilled
    buf ^= 1;
  }
  // reduce the 8 pixel sets, then one atomic per (filter, tap) and block
  __syncthreads();
  float* red = smem;  // [8 sets][32 threads][64 values (+1 pad)]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(ps * 32 + r) * 65 + i * 8 + j] = acc[i][j];
  __syncthreads();
  // 32 threads x 64 values = 2048 sums: every thread of the block takes 8 of them
  for (int e = tid; e < 32 * 64; e += 256) {
    const int rr = e >> 6, idx = e & 63;
    const int i = idx >> 3, j = idx & 7;
    const int kg2 = rr >> 3, fg2 = rr & 7;
    const int k = kc * 32 + kg2 + 4 * i;
    const int f = f0 + (j & 3) + (j >> 2) * 32 + fg2 * 4;
    if (k >= K || f >= d.F) continue;
    float s = 0.0f;
#pragma unroll
    for (int set = 0; set < 8; ++set) s += red[(set * 32 + rr) * 65 + idx];
    atomicAdd(dw + (size_t)f * K + k, s);
  }
}

// ------------------------------------------------------------------ d_images
// block = (n, ROWS_DIMG consecutive input rows, strip of TQ*4 input pixels); threads = TQ pixel quads x
// 16 filter slices (4 filters each). The KH rows of dout that feed an input row live in a ring buffer,
// so every new input row stages one new dout row. A thread gathers, for its 4 pixels and up to 4
// channels, the contributions of its 4 filters over all (dy, dx); the 16 slices are combined with
// xor-shuffles.
constexpr int ROWS_DIMG = 8;

template <int KW_T>
__global__ void __launch_bounds__(256) conv2_dimg_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                         float* __restrict__ dimg, ConvDims d, int accumulate, int tq_n) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  pdl_wait();
  constexpr int KW = KW_T;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int q = tid >> 4, fs = tid & 15;
  const int n = blockIdx.z, iy0 = blockIdx.y * ROWS_DIMG, ix0 = blockIdx.x * tq_n * 4;
  const int iy1 = min(iy0 + ROWS_DIMG, d.H);
  const int tw = tq_n * 4 + KW - 1;                  // staged dout pixels per row
  float* dout_s = smem;                              // ring: [KH+1][tw][FC]
  float* w_s = smem + (d.KH + 1) * tw * FC;          // [KH][KW][4 filters of a slice][16 slices][4 channels]
  const int fchunks = (d.F + FC - 1) / FC;
  for (int c0 = 0; c0 < d.C; c0 += MAX_CC) {
    const int cc = min(MAX_CC, d.C - c0);
    for (int fchunk = 0; fchunk < fchunks; ++fchunk) {
      const int f0 = fchunk * FC;
      __syncthreads();
      for (int i = tid; i < d.KH * KW * FC * MAX_CC; i += nthreads) {
        const int c = i % MAX_CC, sl = (i / MAX_CC) % 16, j = (i / (MAX_CC * 16)) % 4;
        const int dx = (i / (MAX_CC * FC)) % KW, dy = i / (MAX_CC * FC * KW);
        const int f = f0 + sl * 4 + j;
        w_s[i] = (c < cc && f < d.F) ? __ldg(w + (((size_t)f * d.KH + dy) * KW + dx) * d.C + c0 + c) : 0.0f;
      }
      // ring of KH+1 dout rows (slot = oy mod (KH+1)): row iy+1 streams in with cp.async while row iy is
      // being processed
      const int ring = d.KH + 1;
      const bool f_vec = (d.F & 3) == 0;
      auto stage_row = [&](int oy) {
        float* dst = dout_s + (((oy % ring) + ring) % ring) * tw * FC;
        const bool row_ok = oy >= 0 && oy < d.OH;
        const int oyc = min(max(oy, 0), d.OH - 1);
        for (int i = tid; i < tw * (FC / 4); i += nthreads) {
          const int f4 = (i % (FC / 4)) * 4, t = i / (FC / 4);
          const int ox = ix0 - (KW - 1) + t;
          const bool ok = row_ok && ox >= 0 && ox < d.OW;
          const int oxc = min(max(ox, 0), d.OW - 1);
          const float* src = dout + (((size_t)n * d.OH + oyc) * d.OW + oxc) * d.F + f0 + f4;
          float* o = dst + t * FC + f4;
          if (f_vec && f0 + f4 + 4 <= d.F) {
            const unsigned bytes = ok ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(o)), "l"(src),
                         "r"(bytes) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = (ok && f0 + f4 + j < d.F) ? __ldg(src + j) : 0.0f;
          }
        }
      };
      for (int r = 0; r < d.KH; ++r) stage_row(iy0 - (d.KH - 1) + r);
      asm volatile("cp.async.commit_group;" ::: "memory");
      for (int iy = iy0; iy < iy1; ++iy) {
        if (iy + 1 < iy1) {
          stage_row(iy + 1);
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        float acc[4][MAX_CC];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < MAX_CC; ++c) acc[p][c] = 0.0f;
        for (int dy = 0; dy < d.KH; ++dy) {
          const int oy = iy - dy;
          const float* rowp = dout_s + (((oy % ring) + ring) % ring) * tw * FC;
          // pixel ix = ix0 + q*4 + p receives dout[.., ix - dx, ..]: staged column t = q*4 + p - dx + KW-1
          float4 e[4 + KW - 1];
#pragma unroll
          for (int t = 0; t < 4 + KW - 1; ++t) e[t] = *reinterpret_cast<const float4*>(rowp + (q * 4 + t) * FC + fs * 4);
#pragma unroll
          for (int dx = 0; dx < KW; ++dx) {
            const float4* wp = reinterpret_cast<const float4*>(w_s + (size_t)(dy * KW + dx) * FC * MAX_CC) + fs;
            const float4 w0 = wp[0], w1 = wp[16], w2 = wp[32], w3 = wp[48];  // filters fs*4+0..3, 4 channels each
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float4 ev = e[p - dx + KW - 1];
              acc[p][0] = fmaf(ev.x, w0.x, fmaf(ev.y, w1.x, fmaf(ev.z, w2.x, fmaf(ev.w, w3.x, acc[p][0]))));
              acc[p][1] = fmaf(ev.x, w0.y, fmaf(ev.y, w1.y, fmaf(ev.z, w2.y, fmaf(ev.w, w3.y, acc[p][1]))));
              acc[p][2] = fmaf(ev.x, w0.z, fmaf(ev.y, w1.z, fmaf(ev.z, w2.z, fmaf(ev.w, w3.z, acc[p][2]))));
              acc[p][3] = fmaf(ev.x, w0.w, fmaf(ev.y, w1.w, fmaf(ev.z, w2.w, fmaf(ev.w, w3.w, acc[p][3]))));
            }
          }
        }
        // combine the 16 filter slices (lanes differing in the low 4 bits)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < MAX_CC; ++c) {
            float v = acc[p][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            acc[p][c] = v;
          }
        if (fs == 0) {
          const bool acc_out = accumulate || fchunk > 0;   // later filter chunks add to the first one's result
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int ix = ix0 + q * 4 + p;
            if (ix >= d.W) continue;
            float* o = dimg + (((size_t)n * d.H + iy) * d.W + ix) * d.C + c0;
#pragma unroll
            for (int c = 0; c < MAX_CC; ++c)
              if (c < cc) o[c] = acc_out ? o[c] + acc[p][c] : acc[p][c];
          }
        }
        __syncthreads();   // the ring slot of the oldest row is refilled next
      }
    }
  }
}

void check_dims(const ConvDims& d) {
  if (d.N <= 0 || d.OH <= 0 || d.OW <= 0 || d.F <= 0 || d.C <= 0) fail(EGB_ERR_GPU, "conv2: empty tensor");
  if (d.OH != d.H - d.KH + 1 || d.OW != d.W - d.KW + 1) fail(EGB_ERR_GPU, "conv2: inconsistent shapes");
  if (d.OH > 65535 || (long)d.N * ((d.F + FC - 1) / FC) > 65535 || d.H > 65535)
    fail(EGB_ERR_GPU, "conv2: tensor too large for the launch grid");
}

}  // namespace

// Pixel quads per block: the smallest block count that covers the row, then the smallest strip that
// still covers it with that many blocks (e.g. OW = 222 -> 2 strips of 28 quads instead of 32 + 24).
static int groups_per_block(int pixels, int group, int max_groups) {
  const int groups = (pixels + group - 1) / group;
  const int blocks = (groups + max_groups - 1) / max_groups;
  return (groups + blocks - 1) / blocks;
}
static int quads_per_block(int pixels, int max_quads) { return groups_per_block(pixels, 4, max_quads); }

void launch_conv2_fwd(Context& ctx, const float* img, const float* w, float* out, int N, int H, int W, int C, int F,
                      int KH, int KW, bool accumulate, cudaStream_t st) {
  ConvDims d = {N, H, W, C, F, KH, KW, H - KH + 1, W - KW + 1};
  check_dims(d);
  if (conv2_fwd_tc_supported(out, C, F, KH, KW)) {
    launch_conv2_fwd_tc(ctx, img, w, out, N, H, W, C, F, KH, KW, accumulate, st);
    return;
  }
  const int cc = C < MAX_CC ? C : MAX_CC;
  const int tq = groups_per_block(d.OW, PXT, 32);
  const size_t smem = ((size_t)KH * (tq * PXT + KW - 1) * MAX_CC + (size_t)KH * KW * cc * FC) * sizeof(float);
  dim3 grid((d.OW + tq * PXT - 1) / (tq * PXT), (d.OH + ROWS_FWD - 1) / ROWS_FWD, N * ((F + FC - 1) / FC));
  Launch l(ctx, KC_CONV, st);
  if (KW == 3) {
    EGB_CUDA(cudaFuncSetAttribute(conv2_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(ctx, conv2_fwd_kernel<3>, grid, dim3(tq * 8), smem, st, img, w, out, d, accumulate ? 1 : 0, tq);
  } else {
    EGB_CUDA(cudaFuncSetAttribute(conv2_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(ctx, conv2_fwd_kernel<0>, grid, dim3(tq * 8), smem, st, img, w, out, d, accumulate ? 1 : 0, tq);
  }
}

void launch_conv2_dw(Context& ctx, const float* img, const float* dout, float* dw, int N, int H, int W, int C, int F,
                     int KH, int KW, cudaStream_t st) {
  ConvDims d = {N, H, W, C, F, KH, KW, H - KH + 1, W - KW + 1};
  check_dims(d);
  if (conv2_dw_tc_supported(dout, C, F, KH, KW)) {
    launch_conv2_dw_tc(ctx, img, dout, dw, N, H, W, C, F, KH, KW, st);
    return;
  }
  const int K = KH * KW * C;
  const int txw = quads_per_block(d.OW, TX_BWD / 4) * 4;   // even chunks of the output row, <= 64 pixels
  const size_t stage = 2 * (((size_t)TX_BWD * FC + (size_t)KH * (txw + KW - 1) * C + 3) & ~size_t(3)) * sizeof(float);
  const size_t red = (size_t)8 * 32 * 65 * sizeof(float);
  const size_t smem = stage > red ? stage : red;
  dim3 grid(ctx.sm_count * 3, ((K + 31) / 32) * ((F + FC - 1) / FC));
  EGB_CUDA(cudaFuncSetAttribute(conv2_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Launch l(ctx, KC_CONV_DW, st);
  launch_kernel(ctx, conv2_dw_kernel, grid, dim3(256), smem, st, img, dout, dw, d, txw);
}

void launch_conv2_dimg(Context& ctx, const float* dout, const float* w, float* dimg, int N, int H, int W, int C, int F,
                       int KH, int KW, bool accumulate, cudaStream_t st) {
  ConvDims d = {N, H, W, C, F, KH, KW, H - KH + 1, W - KW + 1};
  check_dims(d);
  if (conv2_dimg_tc_supported(dout, C, F, KH, KW)) {
    launch_conv2_dimg_tc(ctx, dout, w, dimg, N, H, W, C, F, KH, KW, accumulate, st);
    return;
  }
  const int tq = quads_per_block(W, 16);
  const size_t smem = ((size_t)(KH + 1) * (tq * 4 + KW - 1) * FC + (size_t)KH * KW * FC * MAX_CC) * sizeof(float);
  dim3 grid((W + tq * 4 - 1) / (tq * 4), (H + ROWS_DIMG - 1) / ROWS_DIMG, N);
  Launch l(ctx, KC_CONV_DIMG, st);
#define EGB_DIMG(KWT)                                                                                       \
  EGB_CUDA(cudaFuncSetAttribute(conv2_dimg_kernel<KWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  launch_kernel(ctx, conv2_dimg_kernel<KWT>, grid, dim3(tq * 16), smem, st, dout, w, dimg, d, accumulate ? 1 : 0, tq);
  switch (KW) {
    case 1: EGB_DIMG(1) break;
    case 2: EGB_DIMG(2) break;
    case 3: EGB_DIMG(3) break;
    case 4: EGB_DIMG(4) break;
    case 5: EGB_DIMG(5) break;
    case 7: EGB_DIMG(7) break;
    default: fail(EGB_ERR_GPU, "conv2 d_images: unsupported filter width %d", KW);
  }
#undef EGB_DIMG
}

bool conv2_dimg_supported(int KW) { return KW == 1 || KW == 2 || KW == 3 || KW == 4 || KW == 5 || KW == 7; }

}  // namespace egb
