// Generic loop-nest kernel: executes any lowered exprgrad kernel (exprgrad/ir.nim:211-220) on the
// device, so that every `++=` expression runs under the B200 backend even when no specialised
// kernel matches (the role clgen.nim:74-257 plays for OpenCL in the reference). The expression is a
// small register program held in the kernel parameter (constant bank); every thread evaluates it
// for its share of the iteration space. Opcode semantics follow the reference's CPU lowering
// (exprgrad/llvmgen.nim:193-321): separate fmul/fadd (this file is built with -fmad=false), `0 - x`
// negate, ordered float compares, truncating integer division, select evaluating both arms.
//
// Thread layout: a block of 256 threads covers PB output points x RB reduction slices
// (PB * RB == 256). `points_fast` chooses which of the two is the fastest-varying thread index, so
// that the dimension that is contiguous in memory is the one neighbouring lanes walk. Partial sums
// of the RB slices are combined through shared memory. RB == 1 keeps the reference's sequential
// accumulation order (bit-exact mode).
#include <math.h>

#include <type_traits>

#include "egb_internal.hpp"
#include "interp.hpp"

namespace egb {
namespace {

union Slot {
  float f;
  int64_t i;
  uint64_t u;
};

constexpr int IP_THREADS = 256;

// The register file of the interpreted program. In local memory every first touch of a slot is a
// cache miss on the thread's stack (the dominant cost of small kernels: ~5 us for a ten-instruction
// program); programs that fit IP_SMEM_SLOTS keep it in shared memory instead, laid out [slot][thread].
struct LocalSlots {
  Slot* s;
  __device__ __forceinline__ Slot& operator[](int i) const { return s[i]; }
};
struct SharedSlots {
  Slot* base;  // already offset by the thread index
  __device__ __forceinline__ Slot& operator[](int i) const { return base[i * IP_THREADS]; }
};

template <class S>
__device__ __forceinline__ void run_instrs(const IpInstr* __restrict__ ins, int n, S s, const IpProgram& p) {
  for (int k = 0; k < n; ++k) {
    const IpInstr in = ins[k];
    Slot r;
    r.u = 0;
    switch (in.op) {
      case IP_FADD: r.f = s[in.a].f + s[in.b].f; break;
      case IP_FSUB: r.f = s[in.a].f - s[in.b].f; break;
      case IP_FMUL: r.f = s[in.a].f * s[in.b].f; break;
      case IP_FDIV: r.f = __fdiv_rn(s[in.a].f, s[in.b].f); break;
      case IP_FNEG: r.f = 0.0f - s[in.a].f; break;
      case IP_SIN: r.f = sinf(s[in.a].f); break;
      case IP_COS: r.f = cosf(s[in.a].f); break;
      case IP_EXP: r.f = expf(s[in.a].f); break;
      case IP_LN: r.f = logf(s[in.a].f); break;
      case IP_SQRT: r.f = __fsqrt_rn(s[in.a].f); break;
      case IP_POW: r.f = powf(s[in.a].f, s[in.b].f); break;
      case IP_LOG10: r.f = log10f(s[in.a].f); break;
      case IP_LOG2: r.f = log2f(s[in.a].f); break;
      case IP_LOGB: r.f = __fdiv_rn(logf(s[in.a].f), logf(s[in.b].f)); break;
      case IP_IADD: r.i = s[in.a].i + s[in.b].i; break;
      case IP_ISUB: r.i = s[in.a].i - s[in.b].i; break;
      case IP_IMUL: r.i = s[in.a].i * s[in.b].i; break;
      case IP_IDIV: { const int64_t d = s[in.b].i; r.i = d ? s[in.a].i / d : 0; break; }
      case IP_IMOD: { const int64_t d = s[in.b].i; r.i = d ? s[in.a].i % d : 0; break; }
      case IP_IWRAP: {
        const int64_t d = s[in.b].i;
        r.i = d ? ((s[in.a].i % d) + d) % d : 0;
        break;
      }
      case IP_INEG: r.i = 0 - s[in.a].i; break;
      case IP_FEQ: r.i = s[in.a].f == s[in.b].f; break;
      case IP_FLT: r.i = s[in.a].f < s[in.b].f; break;
      case IP_FLE: r.i = s[in.a].f <= s[in.b].f; break;
      case IP_IEQ: r.i = s[in.a].i == s[in.b].i; break;
      case IP_ILT: r.i = s[in.a].i < s[in.b].i; break;
      case IP_ILE: r.i = s[in.a].i <= s[in.b].i; break;
      case IP_BEQ: r.i = (s[in.a].i != 0) == (s[in.b].i != 0); break;
      case IP_AND: r.i = (s[in.a].i != 0) & (s[in.b].i != 0); break;
      case IP_OR: r.i = (s[in.a].i != 0) | (s[in.b].i != 0); break;
      case IP_SELECT: r.u = s[in.a].i ? s[in.b].u : s[in.c].u; break;
      case IP_TOSCALAR: r.f = (float)s[in.a].i; break;
      case IP_TOINDEX: r.i = (int64_t)s[in.a].f; break;
      case IP_ARRAY_READ: r.u = s[p.array_table[in.imm + (int)s[in.a].i]].u; break;
      default: break;
    }
    s[in.dst] = r;
  }
}

template <class S>
__device__ __forceinline__ int64_t flat_index(const IpTensorOp& op, S s) {
  int64_t idx = op.offset;
  for (int t = 0; t < op.nterms; ++t) idx += op.coef[t] * s[op.slot[t]].i;
  return idx;
}

// Decode a linear index into the iterators of loops [lo, hi) (last loop fastest).
template <class S>
__device__ __forceinline__ void decode(const IpProgram& p, int lo, int hi, int64_t lin, S s) {
  if (lin < 0x7fffffffLL) {
    uint32_t v = (uint32_t)lin;
    for (int l = hi - 1; l >= lo; --l) {
      const uint32_t c = (uint32_t)p.loops[l].count;
      const uint32_t q = v / c;
      s[p.loops[l].slot].i = p.loops[l].start + p.loops[l].step * (int64_t)(v - q * c);
      v = q;
    }
  } else {
    int64_t v = lin;
    for (int l = hi - 1; l >= lo; --l) {
      const int64_t c = p.loops[l].count;
      const int64_t q = v / c;
      s[p.loops[l].slot].i = p.loops[l].start + p.loops[l].step * (v - q * c);
      v = q;
    }
  }
}

template <bool kStrict, bool kSmemSlots>
__global__ void __launch_bounds__(IP_THREADS) interp_kernel(const __grid_constant__ IpProgram p, int pb, int rb,
                                                            int points_fast, int rsplit) {
  __shared__ float partial[IP_THREADS];
  extern __shared__ __align__(16) unsigned char slot_smem[];
  Slot local_slots[kSmemSlots ? 1 : IP_MAX_SLOTS];
  typename std::conditional<kSmemSlots, SharedSlots, LocalSlots>::type s;
  if constexpr (kSmemSlots) s.base = reinterpret_cast<Slot*>(slot_smem) + threadIdx.x;
  else s.s = local_slots;
  const int t = threadIdx.x;
  const int pl = points_fast ? t % pb : t / rb;  // point within the block
  const int rl = points_fast ? t / pb : t % rb;  // reduction slice
  const int64_t nblocks = (p.npoints + pb - 1) / pb;
  float* const out = reinterpret_cast<float*>(p.write.base);
  pdl_launch_dependents();
  for (int i = 0; i < p.nlits; ++i) s[p.lit_slot[i]].u = p.lits[i];
  pdl_wait();

  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t point = blk * pb + pl;
    const bool active = point < p.npoints;
    float acc = 0.0f;
    int64_t widx = 0;
    if (active) {
      decode(p, 0, p.npar, point, s);
      if (!p.scatter) {
        // the write index does not depend on reduction loops: evaluate it once
        if (p.nred > 0) {
          decode(p, p.npar, p.nloops, 0, s);
          run_instrs(p.index_instrs, p.nindex_instrs, s, p);
          widx = flat_index(p.write, s);
        }
        if (kStrict && p.accumulate && p.nred > 0) acc = out[widx];
      }
      // long reductions with few outputs are additionally split over blockIdx.y; the partial sums of
      // the splits meet in the output through atomicAdd (the output then always accumulates)
      int64_t r_begin = 0, r_end = p.nred;
      if (rsplit > 1) {
        const int64_t chunk = ((p.nred + rsplit - 1) / rsplit + rb - 1) / rb * rb;
        r_begin = (int64_t)blockIdx.y * chunk;
        r_end = min(p.nred, r_begin + chunk);
      }
      for (int64_t r = r_begin + rl; r < r_end; r += rb) {
        decode(p, p.npar, p.nloops, r, s);
        run_instrs(p.index_instrs, p.nindex_instrs, s, p);
        for (int k = 0; k < p.nreads; ++k) {
          const IpTensorOp& op = p.reads[k];
          s[op.dst].u = 0;
          s[op.dst].f = reinterpret_cast<const float*>(op.base)[flat_index(op, s)];
        }
        run_instrs(p.instrs, p.ninstrs, s, p);
        const float v = s[p.write.dst].f;
        if (p.scatter) {
          const int64_t w = flat_index(p.write, s);
          out[w] = p.accumulate ? out[w] + v : v;
        } else {
          acc = acc + v;
        }
      }
    }
    if (p.scatter) continue;
    if (rb > 1) {
      partial[t] = acc;
      __syncthreads();
      for (int half = rb >> 1; half > 0; half >>= 1) {
        if (rl < half) {
          const int other = points_fast ? t + half * pb : t + half;
          partial[t] += partial[other];
        }
        __syncthreads();
      }
      acc = partial[t];
      __syncthreads();
    }
    if (active && rl == 0 && p.nred > 0) {
      if (kStrict) out[widx] = acc;
      else if (rsplit > 1) atomicAdd(out + widx, acc);
      else out[widx] = p.accumulate ? out[widx] + acc : acc;
    }
  }
}

// ------------------------------------------------------------------ row-chain kernel
// A run of small kernels that only couple elements of the same row (softmax sums -> normalise ->
// cross-entropy adjoints ..., exprgrad/layers/dnn.nim:90-94, base.nim:66-67 and their derive()d
// kernels) executes in ONE launch: each warp owns rows and runs the whole chain for its row,
// exchanging intermediate tensors through global memory with only __syncwarp() in between.
// Programs are in "row form": loops[0] is the row loop (start 0, step 1).
__global__ void __launch_bounds__(IP_THREADS) interp_rowchain_kernel(const IpProgram* __restrict__ gprogs, int nprogs,
                                                                     int64_t rows) {
  // the programs are staged into shared memory once per block: fetching them field by field from
  // global memory costs an L2 round trip per cache line on every warp's critical path
  extern __shared__ __align__(16) unsigned char chain_smem[];
  IpProgram* progs = reinterpret_cast<IpProgram*>(chain_smem);
  SharedSlots s;  // register file in shared memory, behind the programs (the host guarantees that it fits)
  s.base = reinterpret_cast<Slot*>(chain_smem + (size_t)nprogs * sizeof(IpProgram)) + threadIdx.x;
  pdl_launch_dependents();
  {
    const uint4* src = reinterpret_cast<const uint4*>(gprogs);
    uint4* dst = reinterpret_cast<uint4*>(chain_smem);
    const int n16 = (int)(nprogs * sizeof(IpProgram) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);  // written before the graph ran
  }
  __syncthreads();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < rows; row += nwarps) {
    for (int k = 0; k < nprogs; ++k) {
      const IpProgram& p = progs[k];
      float* const out = reinterpret_cast<float*>(p.write.base);
      for (int i = 0; i < p.nlits; ++i) s[p.lit_slot[i]].u = p.lits[i];
      s[p.loops[0].slot].i = row;
      const int64_t pts = p.npoints / rows;  // points of this row
      if (pts == 1 && p.nred > 1) {
        // one output per row: the lanes share the reduction
        decode(p, 1, p.npar, 0, s);
        float acc = 0.0f;
        for (int64_t r = lane; r < p.nred; r += 32) {
          decode(p, p.npar, p.nloops, r, s);
          run_instrs(p.index_instrs, p.nindex_instrs, s, p);
          for (int q = 0; q < p.nreads; ++q) {
            const IpTensorOp& op = p.reads[q];
            s[op.dst].u = 0;
            s[op.dst].f = reinterpret_cast<const float*>(op.base)[flat_index(op, s)];
          }
          run_instrs(p.instrs, p.ninstrs, s, p);
          acc += s[p.write.dst].f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          decode(p, p.npar, p.nloops, 0, s);
          run_instrs(p.index_instrs, p.nindex_instrs, s, p);
          const int64_t w = flat_index(p.write, s);
          out[w] = p.accumulate ? out[w] + acc : acc;
        }
      } else {
        for (int64_t pt = lane; pt < pts; pt += 32) {
          decode(p, 1, p.npar, pt, s);
          float acc = 0.0f;
          for (int64_t r = 0; r < p.nred; ++r) {
            decode(p, p.npar, p.nloops, r, s);
            run_instrs(p.index_instrs, p.nindex_instrs, s, p);
            for (int q = 0; q < p.nreads; ++q) {
              const IpTensorOp& op = p.reads[q];
              s[op.dst].u = 0;
              s[op.dst].f = reinterpret_cast<const float*>(op.base)[flat_index(op, s)];
            }
            run_instrs(p.instrs, p.ninstrs, s, p);
            acc += s[p.write.dst].f;
          }
          const int64_t w = flat_index(p.write, s);
          out[w] = p.accumulate ? out[w] + acc : acc;
        }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------ 4-wide streaming fast path
// Pure elementwise kernels (one unit-stride loop, fp32/boolean expression): every thread evaluates the
// register program on 4 consecutive elements at a time, with 128-bit loads and stores where the
// access is 16-byte aligned. Same opcode semantics as above, lane by lane.
struct V4 {
  float x, y, z, w;
};
__device__ __forceinline__ float b2f(bool b) { return __uint_as_float(b ? 1u : 0u); }
__device__ __forceinline__ bool f2b(float f) { return __float_as_uint(f) != 0u; }

#define EGB_V4_BIN(expr)                                    \
  {                                                         \
    const V4 a = s[in.a], b = s[in.b];                      \
    r.x = expr(a.x, b.x); r.y = expr(a.y, b.y); r.z = expr(a.z, b.z); r.w = expr(a.w, b.w); \
  }
#define EGB_V4_UN(expr)                                     \
  {                                                         \
    const V4 a = s[in.a];                                   \
    r.x = expr(a.x); r.y = expr(a.y); r.z = expr(a.z); r.w = expr(a.w); \
  }

// register file of the 4-wide paths: shared memory ([slot][thread], 16 bytes per thread) when it fits,
// else local memory
struct V4Local {
  V4* s;
  __device__ __forceinline__ V4& operator[](int i) const { return s[i]; }
};
struct V4Shared {
  V4* base;
  __device__ __forceinline__ V4& operator[](int i) const { return base[i * IP_THREADS]; }
};

template <class S>
__device__ __forceinline__ void run_v4(const IpProgram& p, S s) {
  for (int k = 0; k < p.ninstrs; ++k) {
    const IpInstr in = p.instrs[k];
    V4 r = V4{0.f, 0.f, 0.f, 0.f};
    switch (in.op) {
      case IP_FADD: EGB_V4_BIN([](float a, float b) { return a + b; }) break;
      case IP_FSUB: EGB_V4_BIN([](float a, float b) { return a - b; }) break;
      case IP_FMUL: EGB_V4_BIN([](float a, float b) { return a * b; }) break;
      case IP_FDIV: EGB_V4_BIN(__fdiv_rn) break;
      case IP_FNEG: EGB_V4_UN([](float a) { return 0.0f - a; }) break;
      case IP_SIN: EGB_V4_UN(sinf) break;
      case IP_COS: EGB_V4_UN(cosf) break;
      case IP_EXP: EGB_V4_UN(expf) break;
      case IP_LN: EGB_V4_UN(logf) break;
      case IP_SQRT: EGB_V4_UN(__fsqrt_rn) break;
      case IP_POW: EGB_V4_BIN(powf) break;
      case IP_LOG10: EGB_V4_UN(log10f) break;
      case IP_LOG2: EGB_V4_UN(log2f) break;
      case IP_LOGB: EGB_V4_BIN([](float a, float b) { return __fdiv_rn(logf(a), logf(b)); }) break;
      case IP_FEQ: EGB_V4_BIN([](float a, float b) { return b2f(a == b); }) break;
      case IP_FLT: EGB_V4_BIN([](float a, float b) { return b2f(a < b); }) break;
      case IP_FLE: EGB_V4_BIN([](float a, float b) { return b2f(a <= b); }) break;
      case IP_BEQ: EGB_V4_BIN([](float a, float b) { return b2f(f2b(a) == f2b(b)); }) break;
      case IP_AND: EGB_V4_BIN([](float a, float b) { return b2f(f2b(a) && f2b(b)); }) break;
      case IP_OR: EGB_V4_BIN([](float a, float b) { return b2f(f2b(a) || f2b(b)); }) break;
      case IP_SELECT: {
        const V4 c = s[in.a], a = s[in.b], b = s[in.c];
        r.x = f2b(c.x) ? a.x : b.x; r.y = f2b(c.y) ? a.y : b.y; r.z = f2b(c.z) ? a.z : b.z; r.w = f2b(c.w) ? a.w : b.w;
        break;
      }
      default: break;
    }
    s[in.dst] = r;
  }
}

// The 4-wide kernels keep iterator values in a small array indexed by SLOT (the loop iterators own the
// first slots of a lowered kernel), so the flat element index of an access is a dot product of its terms.
__device__ __forceinline__ int64_t flat_v4(const IpTensorOp& op, const int64_t* itv) {
  int64_t idx = op.offset;
  for (int t = 0; t < op.nterms; ++t) idx += op.coef[t] * itv[op.slot[t]];
  return idx;
}

// q = a / b, a -= q * b with a 32-bit divide when both fit (they almost always do; 64-bit division is a
// ~100-instruction subroutine)
__device__ __forceinline__ int64_t divmod(int64_t& a, int64_t b) {
  int64_t q;
  if (((uint64_t)a | (uint64_t)b) >> 32 == 0) q = (int64_t)((uint32_t)a / (uint32_t)b);
  else q = a / b;
  a -= q * b;
  return q;
}

// iterator values of loops [first, last] from a linear index (last loop fastest)
__device__ __forceinline__ void decode_loops(const IpProgram& p, int first, int last, int64_t v, int64_t* itv) {
  for (int l = last; l >= first; --l) {
    const int64_t c = p.loops[l].count;
    int64_t r = v;
    v = l > first ? divmod(r, c) : 0;
    itv[p.loops[l].slot] = p.loops[l].start + p.loops[l].step * r;
  }
}

// the 4 consecutive elements (along the streaming loop) of one read
__device__ __forceinline__ V4 load_one_v4(const IpTensorOp& op, const int64_t* itv, int64_t i4, int64_t n) {
  const float* src = reinterpret_cast<const float*>(op.base) + flat_v4(op, itv);
  V4 v;
  if (!op.streaming) {
    const float f = __ldg(src);
    v = V4{f, f, f, f};
  } else if (i4 + 3 < n && op.aligned16) {
    const float4 t = *reinterpret_cast<const float4*>(src);
    v = V4{t.x, t.y, t.z, t.w};
  } else {
    v.x = src[0];
    v.y = i4 + 1 < n ? src[1] : 0.0f;
    v.z = i4 + 2 < n ? src[2] : 0.0f;
    v.w = i4 + 3 < n ? src[3] : 0.0f;
  }
  return v;
}

// Software pipeline: the first IP_V4_PRE reads of the NEXT work item are fetched into registers while the
// current item is evaluated, so every thread keeps loads in flight during the interpretive part.
constexpr int IP_V4_PRE = 4;
__device__ __forceinline__ void prefetch_v4(const IpProgram& p, const int64_t* itv, int64_t i4, int64_t n, V4 (&pre)[IP_V4_PRE]) {
#pragma unroll
  for (int k = 0; k < IP_V4_PRE; ++k)
    if (k < p.nreads) pre[k] = load_one_v4(p.reads[k], itv, i4, n);
}
template <class S>
__device__ __forceinline__ void commit_v4(const IpProgram& p, S s, const V4 (&pre)[IP_V4_PRE]) {
#pragma unroll
  for (int k = 0; k < IP_V4_PRE; ++k)
    if (k < p.nreads) s[p.reads[k].dst] = pre[k];
}
template <class S>
__device__ __forceinline__ void load_rest_v4(const IpProgram& p, S s, const int64_t* itv, int64_t i4, int64_t n) {
  for (int k = IP_V4_PRE; k < p.nreads; ++k) s[p.reads[k].dst] = load_one_v4(p.reads[k], itv, i4, n);
}

#define EGB_V4_SLOTS(s)                                                               \
  extern __shared__ __align__(16) unsigned char v4_smem[];                              \
  V4 v4_local[kSmem ? 1 : 64];                                                          \
  typename std::conditional<kSmem, V4Shared, V4Local>::type s;                          \
  if constexpr (kSmem) s.base = reinterpret_cast<V4*>(v4_smem) + (threadIdx.x + threadIdx.y * blockDim.x); \
  else s.s = v4_local;                                                                  \
  for (int i = 0; i < p.nlits; ++i) {                                                   \
    const float f = __uint_as_float((uint32_t)p.lits[i]);                               \
    s[p.lit_slot[i]] = V4{f, f, f, f};                                                  \
  }

// mode 1, elementwise: no reduction loop, the innermost independent loop streams
template <bool kSmem>
__global__ void __launch_bounds__(IP_THREADS) interp_vec4_kernel(const __grid_constant__ IpProgram p) {
  pdl_launch_dependents();
  EGB_V4_SLOTS(s)
  pdl_wait();
  const int inner = p.npar - 1;
  const int inner_slot = p.loops[inner].slot;
  const int64_t n = p.loops[inner].count;              // row length (innermost loop)
  const int64_t groups = (n + 3) >> 2;                 // 4-element groups per row
  const int64_t total = (p.npoints / n) * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t itv[IP_MAX_LOOPS];
  V4 pre[IP_V4_PRE];
  auto locate = [&](int64_t item) {                    // iterator values of a work item; returns i4
    int64_t g = item;
    const int64_t rowi = divmod(g, groups);
    if (inner > 0) decode_loops(p, 0, inner - 1, rowi, itv);
    itv[inner_slot] = p.loops[inner].start + (g << 2);
    return g << 2;
  };
  int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i4 = 0;
  if (item < total) {
    i4 = locate(item);
    prefetch_v4(p, itv, i4, n, pre);
  }
  while (item < total) {
    commit_v4(p, s, pre);
    load_rest_v4(p, s, itv, i4, n);
    float* o = reinterpret_cast<float*>(p.write.base) + flat_v4(p.write, itv);
    const int64_t cur_i4 = i4;
    item += stride;
    if (item < total) {
      i4 = locate(item);
      prefetch_v4(p, itv, i4, n, pre);
    }
    run_v4(p, s);
    V4 v = s[p.write.dst];
    if (cur_i4 + 3 < n && p.write.aligned16) {
      if (p.accumulate) {
        const float4 t = *reinterpret_cast<const float4*>(o);
        v.x = t.x + v.x; v.y = t.y + v.y; v.z = t.z + v.z; v.w = t.w + v.w;
      }
      *reinterpret_cast<float4*>(o) = make_float4(v.x, v.y, v.z, v.w);
    } else {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      for (int j = 0; j < 4; ++j)
        if (cur_i4 + j < n) o[j] = p.accumulate ? o[j] + vv[j] : vv[j];
    }
  }
}

// mode 2, streaming reduction: exactly one reduction loop (unit stride, loops[npar]) that the reads stream
// along; blockIdx.y walks the output points, the blocks along x share one point's reduction range;
// 128-bit loads, per-lane fp32 partial sums, warp-shuffle + shared-memory block reduction, one
// atomicAdd per block.
template <bool kSmem>
__global__ void __launch_bounds__(IP_THREADS) interp_vec4_reduce_kernel(const __grid_constant__ IpProgram p) {
  __shared__ float warp_sums[IP_THREADS / 32];
  pdl_launch_dependents();
  EGB_V4_SLOTS(s)
  pdl_wait();
  const int red = p.npar;                               // index of the reduction loop
  const int red_slot = p.loops[red].slot;
  const int64_t n = p.loops[red].count;
  const int64_t groups = (n + 3) >> 2;
  const int64_t chunk = (groups + gridDim.x - 1) / gridDim.x;
  const int64_t g_begin = (int64_t)blockIdx.x * chunk, g_end = min(groups, g_begin + chunk);
  float* const out = reinterpret_cast<float*>(p.write.base);
  int64_t itv[IP_MAX_LOOPS];
  V4 pre[IP_V4_PRE];
  for (int64_t point = blockIdx.y; point < p.npoints; point += gridDim.y) {
    if (red > 0) decode_loops(p, 0, red - 1, point, itv);
    V4 acc = V4{0.f, 0.f, 0.f, 0.f};
    int64_t g = g_begin + threadIdx.x;
    if (g < g_end) {
      itv[red_slot] = p.loops[red].start + (g << 2);
      prefetch_v4(p, itv, g << 2, n, pre);
    }
    while (g < g_end) {
      const int64_t i4 = g << 2;
      commit_v4(p, s, pre);
      load_rest_v4(p, s, itv, i4, n);
      g += IP_THREADS;
      if (g < g_end) {
        itv[red_slot] = p.loops[red].start + (g << 2);
        prefetch_v4(p, itv, g << 2, n, pre);
      }
      run_v4(p, s);
      const V4 v = s[p.write.dst];
      acc.x += v.x;
      if (i4 + 1 < n) acc.y += v.y;
      if (i4 + 2 < n) acc.z += v.z;
      if (i4 + 3 < n) acc.w += v.w;
    }
    float t = (acc.x + acc.y) + (acc.z + acc.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float total = 0.0f;
      for (int w = 0; w < IP_THREADS / 32; ++w) total += warp_sums[w];
      atomicAdd(out + flat_v4(p.write, itv), total);
    }
    __syncthreads();
  }
}

// mode 3, reduction with streaming output points (column sums, naive contractions): the innermost
// independent loop streams (4 outputs per thread, 128-bit loads along it), one reduction loop with
// arbitrary strides. blockDim = (GX point groups) x (GY reduction lanes), blockIdx.y splits the
// reduction range; partial sums meet in shared memory and leave with one atomicAdd per output.
template <bool kSmem>
__global__ void __launch_bounds__(IP_THREADS) interp_vec4_pointsum_kernel(const __grid_constant__ IpProgram p) {
  __shared__ V4 partial[IP_THREADS];
  pdl_launch_dependents();
  EGB_V4_SLOTS(s)
  pdl_wait();
  const int inner = p.npar - 1, red = p.npar;
  const int inner_slot = p.loops[inner].slot, red_slot = p.loops[red].slot;
  const int64_t n = p.loops[inner].count;
  const int64_t groups = (n + 3) >> 2;
  const int64_t total = (p.npoints / n) * groups;
  const int64_t R = p.loops[red].count;
  const int64_t chunk = (R + gridDim.y - 1) / gridDim.y;
  const int64_t r_begin = (int64_t)blockIdx.y * chunk, r_end = min(R, r_begin + chunk);
  const int gx = blockDim.x, gy = blockDim.y, tid = threadIdx.x + threadIdx.y * gx;
  float* const out = reinterpret_cast<float*>(p.write.base);
  int64_t itv[IP_MAX_LOOPS];
  V4 pre[IP_V4_PRE];
  for (int64_t base = (int64_t)blockIdx.x * gx; base < total; base += (int64_t)gridDim.x * gx) {
    const int64_t item = base + threadIdx.x;
    const bool valid = item < total;
    V4 acc = V4{0.f, 0.f, 0.f, 0.f};
    int64_t i4 = 0;
    if (valid) {
      int64_t g = item;
      const int64_t rowi = divmod(g, groups);
      if (inner > 0) decode_loops(p, 0, inner - 1, rowi, itv);
      i4 = g << 2;
      itv[inner_slot] = p.loops[inner].start + i4;
      int64_t r = r_begin + threadIdx.y;
      if (r < r_end) {
        itv[red_slot] = p.loops[red].start + p.loops[red].step * r;
        prefetch_v4(p, itv, i4, n, pre);
      }
      while (r < r_end) {
        commit_v4(p, s, pre);
        load_rest_v4(p, s, itv, i4, n);
        r += gy;
        if (r < r_end) {
          itv[red_slot] = p.loops[red].start + p.loops[red].step * r;
          prefetch_v4(p, itv, i4, n, pre);
        }
        run_v4(p, s);
        const V4 v = s[p.write.dst];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (gy > 1) {
      partial[tid] = acc;
      __syncthreads();
      if (threadIdx.y == 0) {
        for (int j = 1; j < gy; ++j) {
          const V4 o = partial[threadIdx.x + j * gx];
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
      }
      __syncthreads();
    }
    if (valid && threadIdx.y == 0) {
      float* o = out + flat_v4(p.write, itv);   // the write has no term on the reduction loop
      atomicAdd(o, acc.x);
      if (i4 + 1 < n) atomicAdd(o + 1, acc.y);
      if (i4 + 2 < n) atomicAdd(o + 2, acc.z);
      if (i4 + 3 < n) atomicAdd(o + 3, acc.w);
    }
  }
}

}  // namespace

void launch_interp_rowchain(Context& ctx, const IpProgram* dev_progs, int nprogs, int max_slots, int64_t rows,
                            cudaStream_t st) {
  if (rows <= 0 || nprogs <= 0) return;
  const int64_t warps_per_block = IP_THREADS / 32;
  const int64_t nb = (rows + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)ctx.sm_count * 8;
  {
    Launch l(ctx, KC_INTERP, st);
    const size_t smem = (size_t)nprogs * sizeof(IpProgram) + (size_t)max_slots * IP_THREADS * sizeof(Slot);
    if (first_use_on_device(ctx, (const void*)interp_rowchain_kernel))
      EGB_CUDA(cudaFuncSetAttribute(interp_rowchain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    launch_kernel(ctx, interp_rowchain_kernel, dim3((int)(nb < cap ? nb : cap)), dim3(IP_THREADS), smem, st, dev_progs, nprogs,
                  rows);
  }
  EGB_CUDA(cudaGetLastError());
}

int interp_reduction_splits(const IpProgram& prog, int pb, int rb, bool strict, int sm_count) {
  if (strict || prog.scatter) return 1;
  if (prog.vec4 >= 2) return 2;  // the 4-wide reductions always combine their blocks with atomicAdd
  if (prog.vec4 || prog.nred < (1 << 16)) return 1;
  const int64_t nblocks = (prog.npoints + pb - 1) / pb;
  if (nblocks >= 2 * sm_count) return 1;
  int64_t s = (4 * (int64_t)sm_count) / nblocks;
  const int64_t max_s = prog.nred / ((int64_t)rb * 64);
  if (s > max_s) s = max_s;
  if (s > 1024) s = 1024;
  return s < 2 ? 1 : (int)s;
}

void launch_interp(Context& ctx, const IpProgram& prog, int pb, int rb, int points_fast, bool strict,
                   cudaStream_t st, int rsplit) {
  if (prog.npoints <= 0 || prog.nred <= 0) return;
  if (prog.vec4 && !strict) {
    const bool smem_slots = prog.nslots <= 24;   // 24 slots x 256 threads x 16 B = 96 KB
    const size_t smem = smem_slots ? (size_t)prog.nslots * IP_THREADS * sizeof(V4) : 0;
    if (first_use_on_device(ctx, (const void*)interp_vec4_kernel<true>)) {
      EGB_CUDA(cudaFuncSetAttribute(interp_vec4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      EGB_CUDA(cudaFuncSetAttribute(interp_vec4_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      EGB_CUDA(cudaFuncSetAttribute(interp_vec4_pointsum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    if (prog.vec4 == 1) {
      const int64_t row_len = prog.loops[prog.npar - 1].count;
      const int64_t groups = (prog.npoints / row_len) * ((row_len + 3) / 4);
      const int64_t nb = (groups + IP_THREADS - 1) / IP_THREADS;
      const int64_t capb = (int64_t)ctx.sm_count * 16;
      const dim3 grid((int)(nb < capb ? nb : capb));
      Launch l(ctx, KC_ELTWISE, st);
      if (smem_slots) launch_kernel(ctx, interp_vec4_kernel<true>, grid, dim3(IP_THREADS), smem, st, prog);
      else launch_kernel(ctx, interp_vec4_kernel<false>, grid, dim3(IP_THREADS), 0, st, prog);
    } else if (prog.vec4 == 2) {
      if (!prog.accumulate) fail(EGB_ERR_GPU, "interp: a streaming reduction needs an accumulating output");
      const int64_t groups = (prog.nred + 3) / 4;
      int64_t by = prog.npoints < 65535 ? prog.npoints : 65535;
      int64_t bx = (8 * (int64_t)ctx.sm_count + by - 1) / by;            // ~8 blocks per SM in total
      const int64_t max_bx = (groups + IP_THREADS * 4 - 1) / (IP_THREADS * 4);  // >= 4 groups per thread
      if (bx > max_bx) bx = max_bx;
      if (bx < 1) bx = 1;
      Launch l(ctx, KC_REDUCE, st);
      if (smem_slots) launch_kernel(ctx, interp_vec4_reduce_kernel<true>, dim3((int)bx, (int)by), dim3(IP_THREADS), smem, st, prog);
      else launch_kernel(ctx, interp_vec4_reduce_kernel<false>, dim3((int)bx, (int)by), dim3(IP_THREADS), 0, st, prog);
    } else {
      if (!prog.accumulate) fail(EGB_ERR_GPU, "interp: a split reduction needs an accumulating output");
      const int64_t row_len = prog.loops[prog.npar - 1].count;
      const int64_t items = (prog.npoints / row_len) * ((row_len + 3) / 4);
      int gx = 8;                                        // >= 8 lanes x 16 B: whole 128-byte lines
      while (gx < IP_THREADS && gx < items) gx *= 2;
      const int gy = IP_THREADS / gx;
      int64_t bx = (items + gx - 1) / gx;
      if (bx > 65535) bx = 65535;
      int64_t by = (8 * (int64_t)ctx.sm_count + bx - 1) / bx;            // ~8 blocks per SM in total
      const int64_t max_by = prog.nred / ((int64_t)gy * 8);              // >= 8 steps per thread
      if (by > max_by) by = max_by;
      if (by > 65535) by = 65535;
      if (by < 1) by = 1;
      Launch l(ctx, KC_REDUCE, st);
      if (smem_slots) launch_kernel(ctx, interp_vec4_pointsum_kernel<true>, dim3((int)bx, (int)by), dim3(gx, gy), smem, st, prog);
      else launch_kernel(ctx, interp_vec4_pointsum_kernel<false>, dim3((int)bx, (int)by), dim3(gx, gy), 0, st, prog);
    }
    EGB_CUDA(cudaGetLastError());
    return;
  }
  if (pb * rb != IP_THREADS) fail(EGB_ERR_GPU, "interp: PB*RB must be %d", IP_THREADS);
  const int64_t nblocks = (prog.npoints + pb - 1) / pb;
  const int64_t cap = (int64_t)ctx.sm_count * 8;
  const int grid = (int)(nblocks < cap ? nblocks : cap);
  if (rsplit > 1 && !prog.accumulate) fail(EGB_ERR_GPU, "interp: a split reduction needs an accumulating output");
  const dim3 grid3(grid, rsplit > 1 ? rsplit : 1);
  {
    Launch l(ctx, KC_INTERP, st);
    const bool smem_slots = prog.nslots <= IP_SMEM_SLOTS;
    const size_t smem = smem_slots ? (size_t)prog.nslots * IP_THREADS * sizeof(Slot) : 0;
    if (strict) {
      if (smem_slots) launch_kernel(ctx, interp_kernel<true, true>, grid3, dim3(IP_THREADS), smem, st, prog, pb, rb, points_fast, 1);
      else launch_kernel(ctx, interp_kernel<true, false>, grid3, dim3(IP_THREADS), 0, st, prog, pb, rb, points_fast, 1);
    } else {
      if (smem_slots) launch_kernel(ctx, interp_kernel<false, true>, grid3, dim3(IP_THREADS), smem, st, prog, pb, rb, points_fast, rsplit);
      else launch_kernel(ctx, interp_kernel<false, false>, grid3, dim3(IP_THREADS), 0, st, prog, pb, rb, points_fast, rsplit);
    }
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
