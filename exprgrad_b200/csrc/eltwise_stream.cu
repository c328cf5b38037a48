// HBM-streaming map kernels for the fixed elementwise / optimizer forms of exprgrad's layer library:
// activations and their adjoints (exprgrad/layers/dnn.nim:26-40 + derive, passes.nim:383-549), the optimizer
// updates (gradientDescent base.nim:37-38, adam base.nim:40-53), tensor arithmetic (base.nim:19-25) and the
// row-broadcast bias add (dnn.nim:22-24). They replace, for these forms, what the reference's OpenCL generator
// emitted per kernel (clgen.nim:74-190: one work-item per element, scalar loads) and what interp.cu does as a
// register interpreter (issue-bound at 0.37 of the HBM roofline, profiles/r01j_eltwise_ncu_full.txt).
//
// One template instance per form - no interpreter loop, no index arithmetic beyond a grid-stride counter:
// every thread moves four independent 128-bit groups per tensor and iteration (all loads issued before the
// first store), 256-thread blocks, a grid of a few CTAs per SM. Tensors that cannot stay in the 126 MB L2 are
// read with ld.global.cs / written with st.global.cs (evict-first) so that they do not sweep the L2 contents
// of their neighbours; small ones keep the default policy because the next kernel re-reads them from L2.
// Arithmetic restates the IR expression operation by operation (file compiled with -fmad=false: the reference
// emits separate fmul / fadd, llvmgen.nim:219-221; negate is 0 - x, llvm.nim:333-336; division and sqrt IEEE).
#include <stdlib.h>

#include "egb_internal.hpp"
#include "pattern.hpp"
#include "runtime.hpp"

namespace egb {

namespace {

constexpr int ELT_THREADS = 256;
constexpr int ELT_UNROLL = 4;
constexpr int ELT_SCALAR_VARIANT = 1;   // store schedule of the scalar-operand kinds (see elt_stream_kernel; measured: profiles/r04_*)
constexpr int ELT_SCALAR_SMEM = 73728;  // dynamic shared memory request of variant 1 (bytes; residency bound, never touched)
constexpr int ELT_LATE_TRIGGER = 0;     // launches that stream more than 64 MB release their dependents at exit

struct EltArgs {
  float* out;
  const float* in0;
  const float* in1;
  float p0, p1, p2, p3;
  long long n;     // elements
  long long row;   // ELT_BIAS_ROW: row length (multiple of 4)
  int accumulate;  // out += f(...) (InstrWrite) instead of out = f(...) (InstrOverwrite)
  int streaming;   // evict-first cache policy, bit 0: operand loads, bit 1: load of the destination, bit 2: stores
};

template <int KIND>
struct EltFn;

#define EGB_ELT_FN(KIND_, NIN_, BODY)                                                  \
  template <>                                                                          \
  struct EltFn<KIND_> {                                                                \
    static constexpr int kInputs = NIN_;                                               \
    static __device__ __forceinline__ float apply(float x, float y, const EltArgs& a) { \
      (void)x; (void)y; (void)a;                                                       \
      BODY                                                                             \
    }                                                                                  \
  };

EGB_ELT_FN(ELT_COPY, 1, return x;)
EGB_ELT_FN(ELT_BIAS_ROW, 1, return x;)
EGB_ELT_FN(ELT_RELU, 1, return (0.0f <= x) ? x : 0.0f;)
EGB_ELT_FN(ELT_LEAKY, 1, return ((0.0f <= x) ? 1.0f : a.p0) * x;)
EGB_ELT_FN(ELT_SIGMOID, 1, return __fdiv_rn(1.0f, 1.0f + expf(0.0f - x));)
EGB_ELT_FN(ELT_TANH, 1, const float e = expf(x); const float f = expf(0.0f - x); return __fdiv_rn(e - f, e + f);)
EGB_ELT_FN(ELT_SCALE, 1, return x * a.p0;)
EGB_ELT_FN(ELT_SCALE_NEG, 1, return (0.0f - x) * a.p0;)
EGB_ELT_FN(ELT_DIV_CONST, 1, return __fdiv_rn(x, a.p0);)
EGB_ELT_FN(ELT_ADD, 2, return x + y;)
EGB_ELT_FN(ELT_SUB, 2, return x - y;)
EGB_ELT_FN(ELT_MUL, 2, return x * y;)
EGB_ELT_FN(ELT_RELU_ADJ, 2, return (0.0f <= x) ? y : 0.0f;)
EGB_ELT_FN(ELT_LEAKY_ADJ, 2, return y * ((0.0f <= x) ? 1.0f : a.p0);)
// negate(mul(mul(negate(1), div(g, mul(s, s))), e)) with e = exp(negate(x)), s = add(1, e)
EGB_ELT_FN(ELT_SIGMOID_ADJ, 2, const float e = expf(0.0f - x); const float s = 1.0f + e;
           return 0.0f - ((-1.0f * __fdiv_rn(y, s * s)) * e);)
// derive() of (e - f) / (e + f), e = exp(x), f = exp(negate(x)):
//   t = negate(e - f) * (g / ((e + f) * (e + f)));  result = negate((t + negate(g / (e + f))) * f) + (t + g / (e + f)) * e
EGB_ELT_FN(ELT_TANH_ADJ, 2, const float e = expf(x); const float f = expf(0.0f - x); const float s = e + f;
           const float t = (0.0f - (e - f)) * __fdiv_rn(y, s * s); const float q = __fdiv_rn(y, s);
           return (0.0f - ((t + (0.0f - q)) * f)) + ((t + q) * e);)
EGB_ELT_FN(ELT_ADAM_M, 2, return (x * a.p0) + (a.p1 * y);)
EGB_ELT_FN(ELT_ADAM_V, 2, return (x * a.p0) + (a.p1 * (y * y));)
// div(mul(negate(eta), div(m, c1)), add(sqrt(div(v, c2)), eps)); c1 = 1 - pow(b1, epoch), c2 = 1 - pow(b2, epoch)
EGB_ELT_FN(ELT_ADAM_STEP, 2, return __fdiv_rn(a.p0 * __fdiv_rn(x, a.p1), __fsqrt_rn(__fdiv_rn(y, a.p2)) + a.p3);)
// adjoint of sq(x) under a scalar loss, add(mul(g, x), mul(g, x)): y is ONE element (in1[0], the adjoint seed), loaded
// once per thread instead of streamed
EGB_ELT_FN(ELT_SQ_ADJ, 2, return (y * x) + (y * x);)
template <int KIND>
constexpr bool kScalarIn1 = KIND == ELT_SQ_ADJ;

// (the policy bits are uniform across the grid: the branches cost one predicate each)
__device__ __forceinline__ float4 ld4(const float* p, bool cs, bool lu = false) {
  if (lu) return __ldlu(reinterpret_cast<const float4*>(p));   // last use: the line need not stay (the store rewrites it whole)
  if (cs) return __ldcs(reinterpret_cast<const float4*>(p));
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v, bool cs) {
  if (cs) __stcs(reinterpret_cast<float4*>(p), v);
  else *reinterpret_cast<float4*>(p) = v;
}

// column of flat element i in a row-major [rows, row] tensor (32-bit divide when it fits)
__device__ __forceinline__ long long row_offset(long long i, const EltArgs& a) {
  if ((a.n >> 31) == 0) return (long long)((unsigned)i % (unsigned)a.row);
  return i % a.row;
}

// VARIANT (scalar-operand kinds only): 0 = the general body (destination operand, accumulate path, 4-wide copy of the
// one-element operand), 1 = a body without them (overwrite only), what ships
template <int KIND, int UNROLL, int VARIANT = 0>
__global__ void __launch_bounds__(ELT_THREADS, 3) elt_stream_kernel(const EltArgs a) {
  using Fn = EltFn<KIND>;
  const bool cs_in = (a.streaming & 1) != 0, cs_out_ld = (a.streaming & 2) != 0, cs_st = (a.streaming & 4) != 0;
  const bool lu_out = (a.streaming & 8) != 0;
  // bit 4: no early trigger - the dependent launch is released when this grid's CTAs exit. A long streaming launch
  // gains nothing from an early start of its successor, and a successor that becomes co-resident (a one-CTA-per-SM
  // tensor-core kernel whose warps poll mbarriers while they wait) takes issue slots and LSU bandwidth from it.
  if (!(a.streaming & 16)) pdl_launch_dependents();
  pdl_wait();
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * ELT_THREADS;
  const long long first = (long long)blockIdx.x * ELT_THREADS + threadIdx.x;
  // in-place forms (adam: m += f(m, g)) read the destination as an operand already
  const bool alias0 = KIND != ELT_BIAS_ROW && a.in0 == a.out;
  float sc = 0.0f;
  if constexpr (kScalarIn1<KIND>) sc = __ldg(a.in1);
  for (long long base = first; base < n4; base += stride * UNROLL) {
    float4 x[UNROLL], y[UNROLL], o[UNROLL];
    // phase 1: every load of this iteration
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long g = base + (long long)u * stride;
      x[u] = y[u] = o[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < n4) {
        if constexpr (KIND == ELT_BIAS_ROW) x[u] = *reinterpret_cast<const float4*>(a.in0 + row_offset(g << 2, a));
        else x[u] = ld4(a.in0 + (g << 2), alias0 ? cs_out_ld : cs_in, alias0 && lu_out);
        if constexpr (kScalarIn1<KIND>) y[u] = make_float4(sc, sc, sc, sc);
        else if constexpr (Fn::kInputs >= 2) y[u] = ld4(a.in1 + (g << 2), cs_in);
        if constexpr (!(kScalarIn1<KIND> && VARIANT == 1))
          if (a.accumulate) o[u] = alias0 ? x[u] : ld4(a.out + (g << 2), cs_out_ld, lu_out);
      }
    }
    // phase 2: arithmetic + stores
    if constexpr (kScalarIn1<KIND> && VARIANT == 1) {
      // overwrite only (the launcher refuses the accumulating form), no destination operand, no broadcast copies of
      // the seed: 48 registers and 243 instructions where the general body has 67 and 392. Measured 5.7 TB/s standalone
      // on a 3.2 GB tensor against 4.3 for the general body (profiles/r04d_stream_variants.txt; ncu of the general
      // body: 56 % DRAM activity where relu has 76 %, r04c). Which difference matters was not isolated: ptxas sinks each
      // group's arithmetic next to its store in both and reuses one register group for all four results (the empty
      // asm below does not prevent it), so operand reuse behind the stores is NOT the explanation.
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        x[u].x = Fn::apply(x[u].x, sc, a); x[u].y = Fn::apply(x[u].y, sc, a);
        x[u].z = Fn::apply(x[u].z, sc, a); x[u].w = Fn::apply(x[u].w, sc, a);
      }
      static_assert(UNROLL == 4 || UNROLL == 2, "register pinning below is written for 2 or 4 groups");
      if constexpr (UNROLL == 4)
        asm volatile("" : "+f"(x[0].x), "+f"(x[0].y), "+f"(x[0].z), "+f"(x[0].w), "+f"(x[1].x), "+f"(x[1].y), "+f"(x[1].z),
                          "+f"(x[1].w), "+f"(x[2].x), "+f"(x[2].y), "+f"(x[2].z), "+f"(x[2].w), "+f"(x[3].x), "+f"(x[3].y),
                          "+f"(x[3].z), "+f"(x[3].w) : : "memory");
      else
        asm volatile("" : "+f"(x[0].x), "+f"(x[0].y), "+f"(x[0].z), "+f"(x[0].w), "+f"(x[1].x), "+f"(x[1].y), "+f"(x[1].z),
                          "+f"(x[1].w) : : "memory");
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long g = base + (long long)u * stride;
        if (g < n4) st4(a.out + (g << 2), x[u], cs_st);
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long g = base + (long long)u * stride;
      if (g < n4) {
        float4 r;
        r.x = Fn::apply(x[u].x, y[u].x, a);
        r.y = Fn::apply(x[u].y, y[u].y, a);
        r.z = Fn::apply(x[u].z, y[u].z, a);
        r.w = Fn::apply(x[u].w, y[u].w, a);
        if (a.accumulate) {
          r.x = o[u].x + r.x; r.y = o[u].y + r.y; r.z = o[u].z + r.z; r.w = o[u].w + r.w;
        }
        st4(a.out + (g << 2), r, cs_st);
      }
    }
  }
  // tail (n mod 4 elements)
  const long long t = (n4 << 2) + first;
  if (t < a.n) {
    const float x = KIND == ELT_BIAS_ROW ? a.in0[row_offset(t, a)] : a.in0[t];
    float y = 0.0f;
    if constexpr (kScalarIn1<KIND>) y = sc;
    else if constexpr (Fn::kInputs >= 2) y = a.in1[t];
    const float r = Fn::apply(x, y, a);
    a.out[t] = (!kScalarIn1<KIND> && a.accumulate) ? a.out[t] + r : r;
  }
}

// adam (exprgrad/layers/base.nim:40-53) as ONE pass: the reference's three kernels
//   m += m * (b1 - 1) + (1 - b1) * g;   v += v * (b2 - 1) + (1 - b2) * g * g;
//   p += (-eta * (m / c1)) / (sqrt(v / c2) + eps),   c = 1 - pow(b, epoch)
// read g twice and m, v twice and write every line they touch: 40 bytes per element. Fused they move 28 (g, m, v, p in;
// m, v, p out); the arithmetic is the three kernels' own, operation by operation, on the values the third one would
// have re-read (so results are bit-identical to the separate launches).
struct AdamArgs {
  float* p; float* m; float* v; const float* g;
  float a0, a1, b0, b1, s0, s1, s2, s3;
  long long n;
};
__global__ void __launch_bounds__(ELT_THREADS, 3) adam_fused_kernel(const AdamArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * ELT_THREADS;
  const long long first = (long long)blockIdx.x * ELT_THREADS + threadIdx.x;
  auto one = [&](float g, float& m, float& v, float& p) {
    m = m + ((m * a.a0) + (a.a1 * g));
    v = v + ((v * a.b0) + (a.b1 * (g * g)));
    p = p + __fdiv_rn(a.s0 * __fdiv_rn(m, a.s1), __fsqrt_rn(__fdiv_rn(v, a.s2)) + a.s3);
  };
  for (long long base = first; base < n4; base += stride * 2) {
    float4 g[2], m[2], v[2], p[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = base + (long long)u * stride;
      if (i < n4) {
        g[u] = *reinterpret_cast<const float4*>(a.g + (i << 2));
        m[u] = *reinterpret_cast<const float4*>(a.m + (i << 2));
        v[u] = *reinterpret_cast<const float4*>(a.v + (i << 2));
        p[u] = *reinterpret_cast<const float4*>(a.p + (i << 2));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = base + (long long)u * stride;
      if (i < n4) {
        one(g[u].x, m[u].x, v[u].x, p[u].x); one(g[u].y, m[u].y, v[u].y, p[u].y);
        one(g[u].z, m[u].z, v[u].z, p[u].z); one(g[u].w, m[u].w, v[u].w, p[u].w);
        *reinterpret_cast<float4*>(a.m + (i << 2)) = m[u];
        *reinterpret_cast<float4*>(a.v + (i << 2)) = v[u];
        *reinterpret_cast<float4*>(a.p + (i << 2)) = p[u];
      }
    }
  }
  const long long t = (n4 << 2) + first;
  if (t < a.n) {
    float m = a.m[t], v = a.v[t], p = a.p[t];
    one(a.g[t], m, v, p);
    a.m[t] = m; a.v[t] = v; a.p[t] = p;
  }
}

template <int KIND>
void launch_kind(Context& ctx, const EltArgs& a, cudaStream_t st) {
  const long long n4 = a.n >> 2;
  // Four 16-byte groups per thread in flight (EGB_ELT_UNROLL=2 for measurements: relu-adjoint reaches 0.99 of the copy
  // roofline with two, relu, bias add and the in-place forms lose 5-13 points).
  static const int forced = getenv("EGB_ELT_UNROLL") ? atoi(getenv("EGB_ELT_UNROLL")) : 0;
  const int unroll = forced == 2 ? 2 : ELT_UNROLL;
  long long blocks = (n4 + (long long)ELT_THREADS * unroll - 1) / ((long long)ELT_THREADS * unroll);
  const long long cap = (long long)ctx.sm_count * 3;   // three resident CTAs per SM (launch bounds: 85 registers), one wave
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  Launch l(ctx, KC_ELTWISE, st);
  if constexpr (kScalarIn1<KIND>) {
    // EGB_ELT_SCALAR_VARIANT: measurement knob for the two store schedules of the scalar-operand kinds
    static const int variant = getenv("EGB_ELT_SCALAR_VARIANT") ? atoi(getenv("EGB_ELT_SCALAR_VARIANT")) : ELT_SCALAR_VARIANT;
    if (variant == 1) {
      // Variant 1 needs 48 registers: five of its CTAs fit on an SM. The grid (3 x SMs, one static slice of the tensor per
      // CTA) is only balanced when every SM gets exactly three, which the hardware does not guarantee for a launch
      // that trickles in while its predecessor's CTAs retire; an (unused) dynamic shared memory request of a third
      // of the SM's capacity bounds the residency at three.
      static const int smem = getenv("EGB_ELT_SCALAR_SMEM") ? atoi(getenv("EGB_ELT_SCALAR_SMEM")) : ELT_SCALAR_SMEM;
      if (smem > 48 * 1024 && first_use_on_device(ctx, (const void*)elt_stream_kernel<KIND, ELT_UNROLL, 1>))
        EGB_CUDA(cudaFuncSetAttribute(elt_stream_kernel<KIND, ELT_UNROLL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      launch_kernel(ctx, elt_stream_kernel<KIND, ELT_UNROLL, 1>, dim3((unsigned)blocks), dim3(ELT_THREADS), (size_t)smem, st, a);
      return;
    }
  }
  if (unroll == 2) launch_kernel(ctx, elt_stream_kernel<KIND, 2>, dim3((unsigned)blocks), dim3(ELT_THREADS), 0, st, a);
  else launch_kernel(ctx, elt_stream_kernel<KIND, ELT_UNROLL>, dim3((unsigned)blocks), dim3(ELT_THREADS), 0, st, a);
}

}  // namespace

bool eltwise_stream_supported(const EltLaunch& e) {
  if (e.kind == ELT_ADAM_FUSED) {
    auto al = [](const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return e.n > 0 && al(e.out) && al(e.in[0]) && al(e.adam_m) && al(e.adam_v);
  }
  if (e.kind <= ELT_NONE || e.kind >= ELT_KIND_COUNT || e.n <= 0) return false;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!e.out || !aligned(e.out) || !e.in[0] || !aligned(e.in[0])) return false;
  const bool scalar1 = e.kind == ELT_SQ_ADJ;   // in[1] is one element; these kernels only overwrite
  if (scalar1 && e.accumulate) return false;
  if (e.nreads >= 2 && (!e.in[1] || (!scalar1 && !aligned(e.in[1])))) return false;
  if (e.nreads > 2) return false;
  if (e.kind == ELT_BIAS_ROW && (e.row <= 0 || (e.row & 3) != 0 || e.n % e.row != 0)) return false;
  return true;
}

void launch_eltwise_stream(Context& ctx, const EltLaunch& e, cudaStream_t st) {
  if (!eltwise_stream_supported(e)) fail(EGB_ERR_GPU, "eltwise: unsupported launch (kind %d)", e.kind);
  if (e.kind == ELT_ADAM_FUSED) {
    AdamArgs q;
    q.p = e.out; q.m = e.adam_m; q.v = e.adam_v; q.g = e.in[0];
    q.a0 = e.p[0]; q.a1 = e.p[1]; q.b0 = e.p2[0]; q.b1 = e.p2[1];
    q.s0 = e.p3[0]; q.s1 = e.p3[1]; q.s2 = e.p3[2]; q.s3 = e.p3[3];
    q.n = e.n;
    const long long n4 = e.n >> 2;
    long long blocks = (n4 + (long long)ELT_THREADS * 2 - 1) / ((long long)ELT_THREADS * 2);
    const long long cap = (long long)ctx.sm_count * 3;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {
      Launch l(ctx, KC_ELTWISE, st);
      launch_kernel(ctx, adam_fused_kernel, dim3((unsigned)blocks), dim3(ELT_THREADS), 0, st, q);
    }
    EGB_CUDA(cudaGetLastError());
    return;
  }
  EltArgs a;
  a.out = e.out;
  a.in0 = e.in[0];
  a.in1 = e.nreads >= 2 ? e.in[1] : nullptr;
  a.p0 = e.p[0]; a.p1 = e.p[1]; a.p2 = e.p[2]; a.p3 = e.p[3];
  a.n = e.n;
  a.row = e.row > 0 ? e.row : 1;
  a.accumulate = e.accumulate ? 1 : 0;
  // bytes this launch moves; beyond about half of the L2 it streams (evict-first)
  const int streamed_reads = e.kind == ELT_BIAS_ROW ? 0 : e.kind == ELT_SQ_ADJ ? 1 : e.nreads;
  const double bytes = 4.0 * (double)e.n * (1 + streamed_reads + (e.accumulate ? 1 : 0));
  // Out-of-place maps stream (evict-first loads and stores: 0.94-0.96 of the copy roofline on 512 MiB tensors).
  // In-place forms (optimizer updates: the destination is also read) are DRAM-limited differently: the write-back
  // of a line follows its read by a few MB, same DRAM bytes but 62 % instead of 76 % DRAM activity in ncu
  // (profiles/r02c_eltwise_sgd_ncu.txt); they measured best with the default policy (0.80 vs 0.75 with evict-first).
  const bool in_place = e.accumulate || e.out == e.in[0] || (e.nreads >= 2 && e.out == e.in[1]);
  a.streaming = bytes > 64.0e6 ? (in_place ? 0 : 7) : 0;
  if (const char* pol = getenv("EGB_ELT_POLICY")) a.streaming = bytes > 64.0e6 ? atoi(pol) : 0;   // measurement knob
  static const int late = getenv("EGB_ELT_LATE_TRIGGER") ? atoi(getenv("EGB_ELT_LATE_TRIGGER")) : ELT_LATE_TRIGGER;
  if (late && bytes > 64.0e6) a.streaming |= 16;
  switch (e.kind) {
#define EGB_ELT_CASE(K) case K: launch_kind<K>(ctx, a, st); break;
    EGB_ELT_CASE(ELT_COPY) EGB_ELT_CASE(ELT_RELU) EGB_ELT_CASE(ELT_LEAKY) EGB_ELT_CASE(ELT_SIGMOID) EGB_ELT_CASE(ELT_TANH)
    EGB_ELT_CASE(ELT_SCALE) EGB_ELT_CASE(ELT_SCALE_NEG) EGB_ELT_CASE(ELT_DIV_CONST) EGB_ELT_CASE(ELT_ADD) EGB_ELT_CASE(ELT_SUB)
    EGB_ELT_CASE(ELT_MUL) EGB_ELT_CASE(ELT_RELU_ADJ) EGB_ELT_CASE(ELT_LEAKY_ADJ) EGB_ELT_CASE(ELT_SIGMOID_ADJ)
    EGB_ELT_CASE(ELT_TANH_ADJ) EGB_ELT_CASE(ELT_ADAM_M) EGB_ELT_CASE(ELT_ADAM_V) EGB_ELT_CASE(ELT_ADAM_STEP)
    EGB_ELT_CASE(ELT_BIAS_ROW) EGB_ELT_CASE(ELT_SQ_ADJ)
#undef EGB_ELT_CASE
    default: fail(EGB_ERR_GPU, "eltwise: unknown kind %d", e.kind);
  }
  EGB_CUDA(cudaGetLastError());
}

}  // namespace egb
