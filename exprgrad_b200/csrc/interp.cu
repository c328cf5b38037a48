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

__global__ void __launch_bounds__(IP_THREADS) interp_vec4_kernel(const __grid_constant__ IpProgram p) {
  V4 s[64];
  pdl_launch_dependents();
  for (int i = 0; i < p.nlits; ++i) {
    const float f = __uint_as_float((uint32_t)p.lits[i]);
    s[p.lit_slot[i]] = V4{f, f, f, f};
  }
  pdl_wait();
  const int inner = p.npar - 1;
  const int64_t n = p.loops[inner].count;              // row length (innermost loop)
  const int64_t groups = (n + 3) >> 2;                 // 4-element groups per row
  const int64_t total = (p.npoints / n) * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += stride) {
    const int64_t rowi = item / groups;
    const int64_t i4 = (item - rowi * groups) << 2;
    // iterators: outer loops decoded from the row index, the inner one starts at i4
    int64_t it[IP_MAX_LOOPS];
    {
      int64_t v = rowi;
      for (int l = inner - 1; l >= 0; --l) {
        const int64_t c = p.loops[l].count;
        const int64_t q = v / c;
        it[l] = p.loops[l].start + p.loops[l].step * (v - q * c);
        v = q;
      }
      it[inner] = p.loops[inner].start + i4;
    }
    auto flat = [&](const IpTensorOp& op) {
      int64_t idx = op.offset;
      for (int t = 0; t < op.nterms; ++t)
        for (int l = 0; l <= inner; ++l)
          if (p.loops[l].slot == op.slot[t]) idx += op.coef[t] * it[l];
      return idx;
    };
    const bool full = i4 + 3 < n;
    for (int k = 0; k < p.nreads; ++k) {
      const IpTensorOp& op = p.reads[k];
      const float* src = reinterpret_cast<const float*>(op.base) + flat(op);
      V4 v;
      if (!op.streaming) {
        const float f = __ldg(src);
        v = V4{f, f, f, f};
      } else if (full && op.aligned16) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        v = V4{t.x, t.y, t.z, t.w};
      } else {
        v.x = src[0];
        v.y = i4 + 1 < n ? src[1] : 0.0f;
        v.z = i4 + 2 < n ? src[2] : 0.0f;
        v.w = i4 + 3 < n ? src[3] : 0.0f;
      }
      s[op.dst] = v;
    }
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
    V4 v = s[p.write.dst];
    float* o = reinterpret_cast<float*>(p.write.base) + flat(p.write);
    if (full && p.write.aligned16) {
      if (p.accumulate) {
        const float4 t = *reinterpret_cast<const float4*>(o);
        v.x = t.x + v.x; v.y = t.y + v.y; v.z = t.z + v.z; v.w = t.w + v.w;
      }
      *reinterpret_cast<float4*>(o) = make_float4(v.x, v.y, v.z, v.w);
    } else {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      for (int j = 0; j < 4; ++j)
        if (i4 + j < n) o[j] = p.accumulate ? o[j] + vv[j] : vv[j];
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
    static bool attr_set = false;
    if (!attr_set) {
      EGB_CUDA(cudaFuncSetAttribute(interp_rowchain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    launch_kernel(ctx, interp_rowchain_kernel, dim3((int)(nb < cap ? nb : cap)), dim3(IP_THREADS), smem, st, dev_progs, nprogs,
                  rows);
  }
  EGB_CUDA(cudaGetLastError());
}

int interp_reduction_splits(const IpProgram& prog, int pb, int rb, bool strict, int sm_count) {
  if (strict || prog.scatter || prog.vec4 || prog.nred < (1 << 16)) return 1;
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
    const int64_t row_len = prog.loops[prog.npar - 1].count;
    const int64_t groups = (prog.npoints / row_len) * ((row_len + 3) / 4);
    const int64_t nb = (groups + IP_THREADS - 1) / IP_THREADS;
    const int64_t capb = (int64_t)ctx.sm_count * 16;
    {
      Launch l(ctx, KC_ELTWISE, st);
      launch_kernel(ctx, interp_vec4_kernel, dim3((int)(nb < capb ? nb : capb)), dim3(IP_THREADS), 0, st, prog);
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
