// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory, fused with the optimizer update.
//
// The one exchange step of a data-parallel train target sits between the last adjoint kernel and
// the first optimizer kernel (exprgrad/parser.nim:757-766 places the optimizer effects behind the GenBackwards
// block, passes.nim:638-640): every rank holds the gradients of its batch shard in one contiguous bucket, the
// optimizer needs their average. With NCCL this was ncclAllReduce(avg) + one gradientDescent kernel per parameter
// (base.nim:37-38): +43 / +53 / +112 us on a 68 us step at 2 / 4 / 8 GPUs. Here it is ONE kernel per rank:
//
//   A  every CTA tells its partner CTAs on all peers "my rank's gradients are complete" (a flag store into the
//      peer's memory) and waits for theirs;
//   B  reduce-scatter: rank r sums slice r of all N buckets (N-1 of them read through NVLink, fixed rank order,
//      so the result does not depend on timing), divides by N and
//      all-gather: stores the averaged slice into every rank's bucket (N-1 remote stores);
//   C  flags again: "slice r has landed everywhere";
//   D  the gradientDescent update P += (0 - g) * rate of every parameter, straight from the averaged bucket
//      (parameters never leave the rank; replicas stay bit-identical because every rank applies the same values).
//
// CTA b of a rank only ever talks to CTA b of the other ranks (per-CTA flag slots), so there is no grid-wide
// barrier and no atomics; flags carry a per-CTA epoch that lives in device memory, which lets the kernel sit in a
// replayed CUDA graph with constant parameters. Per rank and step (N - 1) / N of the bucket crosses NVLink in
// each direction. Waits are bounded: a peer that never arrives traps the kernel instead of hanging the GPU.
#include "egb_internal.hpp"
#include "runtime.hpp"

namespace egb {

namespace {

constexpr int EX_THREADS = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// peer / freshly written data: always from L2 or the link, never from this SM's L1
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// wait until *flag >= epoch (flags only grow); trap after ~2 s so a missing peer cannot hang the device
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch) {
  const unsigned long long t0 = globaltimer();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (globaltimer() - t0 > 2000000000ull) {
      printf("egb exchange: rank flag %p never reached epoch %u (peer missing)\n", (const void*)flag, epoch);
      __trap();
    }
  }
}

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
  // ---- A: gradients of this rank are complete -> tell CTA b of every rank, wait for CTA b of every rank
  if (tid < N) {
    __threadfence_system();
    st_release_sys(p.flags[tid] + (size_t)me * EX_MAX_CTAS + b, epoch);
    wait_flag(my_flags + (size_t)tid * EX_MAX_CTAS + b, epoch);
  }
  __syncthreads();
  if (stamp) stamps[1] = globaltimer();
  // ---- B: reduce slice `me` from all buckets, average, store it into every bucket
  const long long n4 = p.n >> 2;                       // 16-byte groups (bucket tensors are 256-byte aligned)
  const long long S = (n4 + N - 1) / N;                // groups per rank slice
  const long long C = (S + G - 1) / G;                 // groups per CTA and slice
  {
    const long long lo = (long long)me * S + (long long)b * C;
    const long long hi = min(min(lo + C, (long long)(me + 1) * S), n4);
    const float nf = (float)N;
    for (long long g = lo + tid; g < hi; g += EX_THREADS) {
      float4 v[EX_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < EX_MAX_WORLD; ++r)
        if (r < N) v[r] = ld_cg4(p.bucket[r] + (g << 2));
      float4 acc = v[0];
#pragma unroll
      for (int r = 1; r < EX_MAX_WORLD; ++r)
        if (r < N) {
          acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w;
        }
      acc.x = __fdiv_rn(acc.x, nf); acc.y = __fdiv_rn(acc.y, nf);
      acc.z = __fdiv_rn(acc.z, nf); acc.w = __fdiv_rn(acc.w, nf);
#pragma unroll
      for (int r = 0; r < EX_MAX_WORLD; ++r)
        if (r < N) *reinterpret_cast<float4*>(p.bucket[r] + (g << 2)) = acc;
    }
  }
  // ---- C: slice `me`, chunk b has landed everywhere
  __syncthreads();
  if (stamp) stamps[2] = globaltimer();
  if (tid < N) {
    __threadfence_system();
    st_release_sys(p.flags[tid] + (size_t)(EX_MAX_WORLD + me) * EX_MAX_CTAS + b, epoch);
    wait_flag(my_flags + (size_t)(EX_MAX_WORLD + tid) * EX_MAX_CTAS + b, epoch);
  }
  __syncthreads();
  if (stamp) stamps[3] = globaltimer();
  // ---- D: gradientDescent from the averaged bucket: chunk b of every slice (exactly the chunks whose flags
  //         this CTA has just seen)
  if (p.nseg > 0) {
    float* const mine = p.bucket[me];
    for (int r = 0; r < N; ++r) {
      const long long lo = (long long)r * S + (long long)b * C;
      const long long hi = min(min(lo + C, (long long)(r + 1) * S), n4);
      for (long long g = lo + tid; g < hi; g += EX_THREADS) {
        const long long i = g << 2;
        int s = -1;
        for (int q = 0; q < p.nseg; ++q)
          if (i >= p.seg[q].off && i < p.seg[q].off + p.seg[q].len) s = q;
        if (s < 0) continue;   // alignment padding between tensors
        const ExchangeSeg sg = p.seg[s];
        const float4 gv = ld_cg4(mine + i);
        float* dst = sg.param + (i - sg.off);
        const long long left = sg.off + sg.len - i;
        if (left >= 4) {
          float4 pv = *reinterpret_cast<const float4*>(dst);
          // P += (0 - g) * rate   (base.nim:37-38; negate is 0 - x, llvm.nim:333-336; un-contracted)
          pv.x = __fadd_rn(pv.x, __fmul_rn(0.0f - gv.x, sg.rate));
          pv.y = __fadd_rn(pv.y, __fmul_rn(0.0f - gv.y, sg.rate));
          pv.z = __fadd_rn(pv.z, __fmul_rn(0.0f - gv.z, sg.rate));
          pv.w = __fadd_rn(pv.w, __fmul_rn(0.0f - gv.w, sg.rate));
          *reinterpret_cast<float4*>(dst) = pv;
        } else {
          const float ge[4] = {gv.x, gv.y, gv.z, gv.w};
          for (int e = 0; e < (int)left; ++e) dst[e] = __fadd_rn(dst[e], __fmul_rn(0.0f - ge[e], sg.rate));
        }
      }
    }
  }
  if (tid == 0) *epoch_slot = epoch;
  if (stamp) stamps[4] = globaltimer();
}

}  // namespace

size_t exchange_flag_bytes() { return (size_t)(2 * EX_MAX_WORLD * EX_MAX_CTAS + EX_MAX_CTAS) * sizeof(uint32_t) + 16 * 8; }

void launch_exchange(Context& ctx, const ExchangeParams& p, cudaStream_t st) {
  if (p.world < 1 || p.world > EX_MAX_WORLD) fail(EGB_ERR_GPU, "exchange: world size %d is not supported (max %d)", p.world, EX_MAX_WORLD);
  if (p.n % 4 != 0) fail(EGB_ERR_GPU, "exchange: bucket length must be a multiple of 4 floats");
  int grid = ctx.sm_count < EX_MAX_CTAS ? ctx.sm_count : EX_MAX_CTAS;
  grid = grid / 8 * 8;
  Launch l(ctx, KC_EXCHANGE, st);
  launch_kernel(ctx, dp_exchange_sgd_kernel, dim3((unsigned)grid), dim3(EX_THREADS), 0, st, p);
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
