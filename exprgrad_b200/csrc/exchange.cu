// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory, fused with the optimizer update.
//
// The one exchange step of a data-parallel train target sits between the last adjoint kernel and
// the first optimizer kernel (exprgrad/parser.nim:757-766 places the optimizer effects behind the GenBackwards
// block, passes.nim:638-640): every rank holds the gradients of its batch shard in one contiguous bucket, the
// optimizer needs their average. With NCCL this was ncclAllReduce(avg) + one gradientDescent kernel per parameter
// (base.nim:37-38): +43 / +53 / +112 us on a 68 us step at 2 / 4 / 8 GPUs. Here it is ONE kernel per rank:
//
//   1  reduce-scatter by PUSH: every rank stores slice s of its gradients into an inbox on rank s (N-1 remote
//      stores of 1/N of the bucket each);
//   2  rank r sums slice r (own part + N-1 inbox slots, all local reads, fixed rank order, so the result does not
//      depend on timing), divides by N and stores the averaged slice into a second inbox on every rank (all-gather
//      by push);
//   3  the gradientDescent update P += (0 - g) * rate of every parameter straight from the averaged values
//      (parameters never leave the rank; replicas stay bit-identical because every rank applies the same values).
// Lines carry their own epoch tag (see st_ll): no fences, no separate flags, no grid-wide barrier, no atomics.
//
// CTA b of a rank only ever talks to CTA b of the other ranks, so there is no grid-wide barrier; the epoch every
// line is tagged with lives in device memory (one slot per CTA, advanced by the kernel itself), which lets the kernel
// sit in a replayed CUDA graph with constant parameters. Per rank and step 2 x (N - 1) / N of the bucket is stored
// into peer memory. Waits are bounded: a peer that never arrives traps the kernel instead of hanging the GPU.
#include <stdlib.h>

#include "egb_internal.hpp"
#include "runtime.hpp"

namespace egb {

namespace {

constexpr int EX_THREADS = 256;
constexpr size_t FLAG_BYTES = (size_t)(2 * EX_MAX_WORLD * EX_MAX_CTAS + EX_MAX_CTAS) * sizeof(uint32_t) + 16 * 8;
constexpr size_t INBOX_OFF = (FLAG_BYTES + 255) & ~(size_t)255;

__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// peer / freshly written data: always from L2 or the link, never from this SM's L1
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// ---- "LL" lines: 16 payload bytes travel as 32 bytes, every 8-byte word = (4 bytes of data, 4-byte epoch tag).
// An 8-byte word is written atomically, so a reader that sees the tag of this step has the data - no fence and
// no separate flag behind a burst of remote stores. (Measured on this box, tools/probes/p2p_probe.cu: a fence.sys
// behind 0.3-1.3 MB of peer stores completes after 8-14 us whatever the size, memcpyPeer of the same bytes takes
// 10-11 us, while a lone 4-byte store + fence crosses in 2 us: it is the fence after a burst that is slow, not the
// link. The first two versions of this kernel - pull, then push + flags - spent 2 x 10-17 us in exactly that.)
__device__ __forceinline__ void st_ll(float* line, float4 v, uint32_t tag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(line), "r"(__float_as_uint(v.x)), "r"(tag), "r"(__float_as_uint(v.y)) : "memory");
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(line + 4), "r"(__float_as_uint(v.z)), "r"(tag), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ float4 ld_ll(const float* line, uint32_t tag, uint32_t ll_backoff_ns) {
  uint32_t a0, f0, a1, f1, b0, g0, b1, g1;
  const unsigned long long t0 = globaltimer();
  uint32_t spins = 0;
  while (true) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(f0), "=r"(a1), "=r"(f1) : "l"(line) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(g0), "=r"(b1), "=r"(g1) : "l"(line + 4) : "memory");
    if (f0 == tag && f1 == tag && g0 == tag && g1 == tag) break;
    // 32 K threads that re-read their line as fast as they can load the L2 with ~4 TB/s of polling - the same L2 the
    // peers' stores are arriving in; sleep between looks
    __nanosleep(ll_backoff_ns);
    if ((++spins & 0xff) == 0 && globaltimer() - t0 > 2000000000ull) {
      printf("egb exchange: line %p never reached epoch %u (peer missing)\n", (const void*)line, tag);
      __trap();
    }
  }
  return make_float4(__uint_as_float(a0), __uint_as_float(a1), __uint_as_float(b0), __uint_as_float(b1));
}

// Per step and CTA b (chunk b of every slice; CTA b of a rank only ever exchanges lines with CTA b of its peers):
//   1  push: chunk b of slice s of this rank's gradients -> inbox 1, slot `me`, of rank s (every s != me);
//   2  chunk b of MY slice: own part + the peers' lines as they arrive, summed in rank order (the result does not
//      depend on timing), averaged, stored into my bucket and pushed into inbox 2, slot `me`, of every peer;
//   3  every slice: the averaged chunk (own: from the bucket, others: inbox 2 as it arrives) goes into the bucket
//      (optimizers other than gradientDescent read it there) and through the fused gradientDescent update.
// Re-use of the inboxes by the next step is safe without further synchronisation: a rank enters step t + 1 only
// after ITS kernel of step t has ended, i.e. after it received every averaged line of step t, which the peers
// produced after consuming inbox 1; and a peer writes inbox 2 of step t + 1 only after it has received this rank's
// step t + 1 contributions, sent after this rank's step t kernel (and its inbox 2 reads) had ended.
__global__ void __launch_bounds__(EX_THREADS, 1) dp_exchange_sgd_kernel(const __grid_constant__ ExchangeParams p) {
  __shared__ uint32_t epoch_sm;
  pdl_wait();   // every gradient kernel of this rank has completed (no early launch_dependents: the grid spins)
  const int tid = threadIdx.x, b = blockIdx.x, G = gridDim.x, N = p.world, me = p.rank;
  uint32_t* const my_flags = p.flags[me];
  uint32_t* const epoch_slot = my_flags + 2 * EX_MAX_WORLD * EX_MAX_CTAS + b;
  if (tid == 0) epoch_sm = *epoch_slot + 1;
  __syncthreads();
  const uint32_t epoch = epoch_sm;
  // phase stamps of CTA 0 and the last CTA (globaltimer, ns): cheap, always on; printed by EGB_EXCHANGE_TRACE
  unsigned long long* const stamps =
      reinterpret_cast<unsigned long long*>(my_flags + 2 * EX_MAX_WORLD * EX_MAX_CTAS + EX_MAX_CTAS) + (b == 0 ? 0 : 8);
  const bool stamp = tid == 0 && (b == 0 || b == G - 1);
  if (stamp) stamps[0] = globaltimer();
  const long long n4 = p.n >> 2;                       // 16-byte groups (bucket tensors are 256-byte aligned)
  const long long S = (n4 + N - 1) / N;                // groups per rank slice
  const long long C = (S + G - 1) / G;                 // groups per CTA and slice
  const size_t inbox2_off = INBOX_OFF + (size_t)N * S * 32;
  float* const own = p.bucket[me];
  auto inbox = [&](int rank, size_t off, int slot, long long gi) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(p.flags[rank]) + off) + (((long long)slot * S + gi) << 3);
  };
  // ---- 1: push my contributions. Item = (peer index, group of the chunk); four loads in flight per thread.
  {
    const long long items = C * (N - 1);
    for (long long base = 0; base < items; base += 4 * EX_THREADS) {
      float4 v[4];
      float* dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long it = base + (long long)u * EX_THREADS + tid;
        dst[u] = nullptr;
        if (it < items) {
          const int pi = (int)(it / C);
          const int sdst = pi + (pi >= me ? 1 : 0);
          const long long gi = (long long)b * C + (it - (long long)pi * C);      // group inside the slice
          const long long g = (long long)sdst * S + gi;
          if (gi < S && g < n4) {
            v[u] = ld_cg4(own + (g << 2));
            dst[u] = inbox(sdst, INBOX_OFF, me, gi);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u]) st_ll(dst[u], v[u], epoch);
    }
  }
  if (stamp) stamps[1] = globaltimer();
  // ---- 2: reduce chunk b of my slice (rank order), average, store it into my bucket and push it to every peer.
  //         Four groups per thread and pass: all own loads are issued first, and while the thread waits for one
  //         peer line the other three are arriving.
  {
    const float nf = (float)N;
    const bool pow2 = (N & (N - 1)) == 0;
    const float inv = 1.0f / nf;      // exact for a power of two: x * inv == x / N bit for bit
    const long long c_lo = (long long)b * C, c_hi = min((long long)(b + 1) * C, S);
    for (long long base = c_lo; base < c_hi; base += 4 * EX_THREADS) {
      float4 acc[4], mine[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long gi = base + (long long)u * EX_THREADS + tid;
        ok[u] = gi < c_hi && (long long)me * S + gi < n4;
        if (ok[u]) mine[u] = ld_cg4(own + (((long long)me * S + gi) << 2));
      }
      for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          const long long gi = base + (long long)u * EX_THREADS + tid;
          const float4 v = r == me ? mine[u] : ld_ll(inbox(me, INBOX_OFF, r, gi), epoch, p.backoff_ns);
          if (r == 0) acc[u] = v;
          else { acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w; }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const long long gi = base + (long long)u * EX_THREADS + tid;
        float4 a = acc[u];
        if (pow2) { a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv; }
        else { a.x = __fdiv_rn(a.x, nf); a.y = __fdiv_rn(a.y, nf); a.z = __fdiv_rn(a.z, nf); a.w = __fdiv_rn(a.w, nf); }
        st_cg4(own + (((long long)me * S + gi) << 2), a);
        for (int r = 0; r < N; ++r)
          if (r != me) st_ll(inbox(r, inbox2_off, me, gi), a, epoch);
      }
    }
  }
  __syncthreads();   // (phase 3 re-reads this CTA's part of the bucket with another thread mapping at the slice ends)
  if (stamp) stamps[2] = globaltimer();
  // ---- 3: the averaged chunk b of every slice: into the bucket, and gradientDescent straight from it (four groups
  //         per thread and pass: gradient loads / line waits, then parameter loads, then stores)
  for (int t = 0; t < N; ++t) {
    const int r = (me + t) % N;   // own slice first (nothing to wait for), the peers' averages are in flight meanwhile
    const long long lo = (long long)r * S + (long long)b * C;
    const long long hi = min(min(lo + C, (long long)(r + 1) * S), n4);
    for (long long base = lo; base < hi; base += 4 * EX_THREADS) {
      float4 gv[4], pv[4];
      float* dst[4];
      long long left[4];
      float rate[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long g = base + (long long)u * EX_THREADS + tid;
        dst[u] = nullptr;
        left[u] = -1;
        if (g >= hi) continue;
        const long long i = g << 2;
        left[u] = 0;
        if (r == me) gv[u] = ld_cg4(own + i);
        for (int q = 0; q < p.nseg; ++q)
          if (i >= p.seg[q].off && i < p.seg[q].off + p.seg[q].len) {
            dst[u] = p.seg[q].param + (i - p.seg[q].off);
            left[u] = p.seg[q].off + p.seg[q].len - i;
            rate[u] = p.seg[q].rate;
          }
        if (dst[u] && left[u] >= 4) pv[u] = *reinterpret_cast<const float4*>(dst[u]);
      }
      if (r != me) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long g = base + (long long)u * EX_THREADS + tid;
          if (left[u] < 0) continue;
          gv[u] = ld_ll(inbox(me, inbox2_off, r, g - (long long)r * S), epoch, p.backoff_ns);
          st_cg4(own + (g << 2), gv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!dst[u]) continue;   // no fused update, or alignment padding between tensors
        if (left[u] >= 4) {
          // P += (0 - g) * rate   (base.nim:37-38; negate is 0 - x, llvm.nim:333-336; un-contracted)
          float4 q = pv[u];
          q.x = __fadd_rn(q.x, __fmul_rn(0.0f - gv[u].x, rate[u]));
          q.y = __fadd_rn(q.y, __fmul_rn(0.0f - gv[u].y, rate[u]));
          q.z = __fadd_rn(q.z, __fmul_rn(0.0f - gv[u].z, rate[u]));
          q.w = __fadd_rn(q.w, __fmul_rn(0.0f - gv[u].w, rate[u]));
          *reinterpret_cast<float4*>(dst[u]) = q;
        } else {
          const float ge[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
          for (int e = 0; e < (int)left[u]; ++e) dst[u][e] = __fadd_rn(dst[u][e], __fmul_rn(0.0f - ge[e], rate[u]));
        }
      }
    }
  }
  if (tid == 0) *epoch_slot = epoch;
  if (stamp) stamps[3] = globaltimer();
}

}  // namespace

size_t exchange_flag_bytes() { return FLAG_BYTES; }
// flag area + the two inboxes (one slot of ceil(n4 / world) lines per source rank each), one peer-mapped allocation
size_t exchange_area_bytes(size_t bucket_bytes, int world) {
  const size_t n4 = bucket_bytes / 16, S = (n4 + world - 1) / world;
  return INBOX_OFF + 2 * (size_t)world * S * 32 + 256;   // two inboxes of `world` slots, 32 bytes per 16 payload bytes
}

void launch_exchange(Context& ctx, const ExchangeParams& p, cudaStream_t st) {
  if (p.world < 1 || p.world > EX_MAX_WORLD) fail(EGB_ERR_GPU, "exchange: world size %d is not supported (max %d)", p.world, EX_MAX_WORLD);
  if (p.n % 4 != 0) fail(EGB_ERR_GPU, "exchange: bucket length must be a multiple of 4 floats");
  int grid = ctx.sm_count < EX_MAX_CTAS ? ctx.sm_count : EX_MAX_CTAS;
  grid = grid / 8 * 8;
  if (p.ctas > 0 && p.ctas < grid) grid = p.ctas;
  ExchangeParams q = p;
  static const int backoff = getenv("EGB_DP_BACKOFF_NS") ? atoi(getenv("EGB_DP_BACKOFF_NS")) : 400;
  q.backoff_ns = backoff;
  Launch l(ctx, KC_EXCHANGE, st);
  launch_kernel(ctx, dp_exchange_sgd_kernel, dim3((unsigned)grid), dim3(EX_THREADS), 0, st, q);
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
