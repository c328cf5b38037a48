// Internal declarations shared by the C-ABI layer, the host-side program runtime and the kernels.
#pragma once
#include <cuda.h>
#include "../../include/egb200.h"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace egb {

// All internal failures are C++ exceptions; the C-ABI boundary converts them into a status
// code plus a thread-local message (mirrors `GpuError`, reference exprgrad/runtimes/cl.nim:41-43).
struct Error : std::runtime_error {
  int code;
  Error(int code_, const std::string& msg) : std::runtime_error(msg), code(code_) {}
};

// status codes: EGB_OK / EGB_ERR_* macros from include/egb200.h

[[noreturn]] void fail(int code, const char* fmt, ...);

#define EGB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      ::egb::fail(EGB_ERR_GPU, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                  __FILE__, __LINE__);                                                          \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

struct Context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  PFN_encodeTiled encode_tiled = nullptr;
  // scratch for operand splitting (bf16 hi/mid planes) used by the standalone GEMM entry point
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  size_t launches = 0;  // number of kernel launches issued through this context
  void* ensure_scratch(size_t bytes);
  bool scratch_tail_ws_zeroed = false;   // standalone GEMM entry point: tail-wave split flags inside the scratch
  void* scratch_tail_ws = nullptr;

  // auxiliary streams + events used to capture independent plan nodes as parallel graph branches
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> fork_events;

  // Programmatic dependent launch: every kernel is launched with the programmatic-stream-serialization
  // attribute and begins with griddepcontrol.wait, so that the launch latency and prologue of kernel
  // N+1 overlap the tail of kernel N (also inside captured CUDA graphs).
  bool pdl = true;

  // Optional per-kernel-class device timing (bench.py's live roofline measurement): when enabled,
  // every launch is bracketed by CUDA events on the launching stream.
  bool timing = false;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t get_event();

  // Debug timeline (EGB_GEMM_TRACE=<file>): CTA 0 of every contraction launch records globaltimer /
  // clock64 stamps at its phase boundaries into this pinned, device-mapped buffer (slot 0 = cursor);
  // dumped as text when the context is destroyed. Null in normal operation.
  unsigned long long* trace = nullptr;
  unsigned long long trace_next = 0;
};
constexpr int TRACE_SLOT_WORDS = 24;
constexpr int TRACE_SLOTS = 4096;

// Kernel classes for the timing interface (egb_context_kernel_time).
enum KernelClass { KC_GEMM = 0, KC_SPLIT = 1, KC_FILL = 2, KC_INTERP = 3, KC_REDUCE = 4, KC_ELTWISE = 5,
                   KC_CONV = 6 /* conv2 forward */, KC_OTHER = 7, KC_CONV_DW = 8, KC_CONV_DIMG = 9,
                   KC_EXCHANGE = 10 /* data-parallel gradient exchange + optimizer */, KC_COUNT = 11 };

// Function attributes (cudaFuncSetAttribute) belong to the CURRENT device: a process that opens contexts on several
// devices has to set them once per device, not once per process. True the first time `key` (a kernel's address, or any
// address that is unique to the call site) is seen on the context's device.
inline bool first_use_on_device(const Context& ctx, const void* key) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> seen;
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert(std::make_pair(ctx.device, key)).second;
}

// Kernel launch through cudaLaunchKernelEx so that the PDL attribute can be attached.
template <typename... KArgs, typename... Args>
inline void launch_kernel(Context& ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx.pdl ? 1 : 0;
  EGB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// Device side of PDL: wait for the prerequisite grid (no-op when launched without the attribute) and
// allow the dependent grid to start its own prologue.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// RAII bracket around one kernel launch: counts it and, if timing is on, records events.
struct Launch {
  Context& ctx;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  int cls;
  Launch(Context& c, int cls_, cudaStream_t s) : ctx(c), st(s), cls(cls_) {
    if (ctx.timing) {
      cudaEvent_t a = ctx.get_event();
      b = ctx.get_event();
      cudaEventRecord(a, st);
      ctx.spans.push_back({cls, a, b});
    }
  }
  ~Launch() {
    if (b) cudaEventRecord(b, st);
    ctx.launches++;
  }
};

// ------------------------------------------------------------------ kernels (host launchers)

// fp32 -> (hi, mid) bf16 planes with x ~= hi + mid (+ lo dropped, |lo| <= 2^-17 |x|).
// src is [rows, cols] row-major with leading dimension ld. If transpose, the planes are written as
// [cols, rows] (so that the reduction dimension becomes contiguous = "K-major").
// dst leading dimension is dst_ld elements. act: 0 none, 1 relu (select(0<=x, x, 0)).
void launch_split_bf16(Context& ctx, const float* src, int rows, int cols, int ld, bool transpose,
                       __nv_bfloat16* hi, __nv_bfloat16* mid, int dst_ld, int act, cudaStream_t st);

// Several non-transposing splits in one launch (split.cu)
struct SplitJob {
  const float* src;
  __nv_bfloat16 *hi, *mid;
  int rows, cols, ld, dst_ld, act;
};
struct SplitBatch {
  static constexpr int MAX_JOBS = 8;
  SplitJob jobs[MAX_JOBS];
  long first_chunk[MAX_JOBS + 1];
  int n;
  uint4* zero_ptr;      // optional: a region to clear in the same launch (the plan's zero-initialised results)
  size_t zero_vec16;    // its size in 16-byte units
};
void launch_split_batch(Context& ctx, const SplitJob* jobs, int n, cudaStream_t st, void* zero_ptr = nullptr,
                        size_t zero_bytes = 0);

void launch_fill_u32(Context& ctx, uint32_t* dst, uint32_t value, size_t n, cudaStream_t st);

enum GemmFlags {
  GEMM_ACCUMULATE = 1,  // C += alpha*acc  (reference `++=` on a tensor that already holds data)
  GEMM_BIAS = 2,        // + bias[col]     (row-broadcast add, reference dnn.nim:22-24)
  GEMM_RELU = 4,        // (raw entry point only) shorthand for epi = EPI_RELU with D = C
  GEMM_SPLIT_OUT = 8,   // additionally emit bf16 hi/mid planes of the final value
  // Dead-store elimination (planner): the fp32 form of C / D is consumed inside this epilogue only (fused stages,
  // column sums, operand planes) and no later kernel reads it - it is computed in registers but not stored.
  // (The epilogue stores of the dense step are L2-write bound: 6 MB per contraction, profiles/r02b_gemm_trace.txt.)
  GEMM_SKIP_C = 64,
  GEMM_SKIP_D = 128,
};

// Second epilogue stage: one reference kernel fused behind the contraction. v = value stored to C.
enum EpiMode {
  EPI_NONE = 0,
  EPI_RELU = 1,        // D = select(0 <= v, v, 0)                      (dnn.nim:26-27)
  EPI_LEAKY = 2,       // D = select(0 <= v, 1, leak) * v               (dnn.nim:29-30)
  EPI_MASK_RELU = 3,   // D = select(0 <= H, v, 0)                      (adjoint of relu, passes.nim:471-476)
  EPI_MASK_LEAKY = 4,  // D = v * select(0 <= H, 1, leak)               (adjoint of leakyRelu)
  EPI_SGD = 5,         // D += (0 - v) * rate                           (base.nim:37-38)
  EPI_SIGMOID = 6,     // D = 1 / (1 + exp(0 - v))                      (dnn.nim:32-33)
  EPI_TANH = 7,        // D = (e - f) / (e + f), e = exp(v), f = exp(0 - v)   (dnn.nim:35-40)
};

struct GemmArgs {
  // operands as bf16 (hi, mid) planes. K-major (default): A is stored [M, lda], B is stored [N, ldb]
  // with K contiguous. MN-major (a_mn / b_mn): A is stored [K, lda] with M contiguous, B is stored
  // [K, ldb] with N contiguous - i.e. the row-major planes of a [K, M] / [K, N] matrix are used as
  // they are, no transposed copy is needed (UMMA descriptor major bits 15/16).
  const __nv_bfloat16 *a_hi = nullptr, *a_mid = nullptr;
  const __nv_bfloat16 *b_hi = nullptr, *b_mid = nullptr;
  int lda = 0, ldb = 0;
  bool a_mn = false, b_mn = false;
  int M = 0, N = 0, K = 0;
  float* C = nullptr;  // [M, ldc] fp32 output
  int ldc = 0;
  const float* bias = nullptr;
  int epi = EPI_NONE;
  float epi_param = 0.0f;       // leak or rate
  float* D = nullptr;           // second-stage output (same layout as C)
  const float* H = nullptr;     // mask source (same layout as C)
  float* colsum = nullptr;      // [N] column sums of the final value, accumulated atomically
  float alpha = 1.0f;
  int flags = 0;
  __nv_bfloat16 *out_hi = nullptr, *out_mid = nullptr;  // GEMM_SPLIT_OUT: [M, ld_out] planes
  int ld_out = 0;
  int bn = 0;        // 0 = choose
  int cluster_k = 0; // cluster split-K factor: 0 = choose, 1 = off, 2/4/8 = CTAs per output tile
  int sm_budget = 0; // SMs this launch may plan for (0 = all): contractions that run concurrently share the machine
  // 2-CTA kernel, tail-wave split: zero-initialised workspace of gemm_2cta_workspace_bytes() owned by the caller
  // (exclusive to this launch site while it runs), or null = whole tiles only
  void* ws = nullptr;
  size_t ws_bytes = 0;
};

// 2-CTA (cta_group::2) variant for large K-major problems with a plain epilogue (gemm_tcgen05_2cta.cu)
bool gemm_2cta_eligible(const GemmArgs& a);
void launch_gemm_bf16x3_2cta(Context& ctx, const GemmArgs& a, cudaStream_t st);
size_t gemm_2cta_workspace_bytes(int sm_count);
void gemm_choose_config(int M, int N, int K, bool b_mn, int sm_count, int* bn, int* tiles);

void launch_gemm_bf16x3(Context& ctx, const GemmArgs& a, cudaStream_t st);
// Latency-optimised kernel for small problems (gemm_lat.cu): TMA-staged epilogue in the TMEM-native layout,
// push-based cluster split-K
bool gemm_lat_eligible(const GemmArgs& a);
bool launch_gemm_lat(Context& ctx, const GemmArgs& a, cudaStream_t st);   // false: not launched (does not fit)
void gemm_lat_plan(int M, int N, int K, bool b_mn, int sms, int max_ck, int* bn, int* ck);  // host-only
int gemm_planned_ctas(const GemmArgs& a, int sm_budget, int machine_sms);                  // host-only
void gemm_plan(int M, int N, int K, bool b_mn, int sms, int machine_sms, int max_ck, int* bn, int* ck);  // host-only

// Fused softmax + crossEntropy forward/adjoint row kernel (fused_rows.cu)
bool softmax_xent_supported(int64_t cols);
// colsum (zeroed by the caller, may be null): += column sums of DH; out_hi/out_mid (may be null): bf16 planes of DH
void launch_softmax_xent_rows(Context& ctx, const float* H, const float* Y, const float* DL, float* S, float* P, float* DP,
                              float* DH, float* DS, int rows, int cols, float* colsum, __nv_bfloat16* out_hi,
                              __nv_bfloat16* out_mid, int ld_out, cudaStream_t st);

// Direct fp32 conv2 kernels (conv2.cu): NHWC images, filters [F, KH, KW, C], valid, stride 1.
void launch_conv2_fwd(Context& ctx, const float* img, const float* w, float* out, int N, int H, int W, int C, int F,
                      int KH, int KW, bool accumulate, cudaStream_t st);
void launch_conv2_dw(Context& ctx, const float* img, const float* dout, float* dw, int N, int H, int W, int C, int F,
                     int KH, int KW, cudaStream_t st);
void launch_conv2_dimg(Context& ctx, const float* dout, const float* w, float* dimg, int N, int H, int W, int C, int F,
                       int KH, int KW, bool accumulate, cudaStream_t st);
bool conv2_dimg_supported(int KW);
// Tensor-core implicit-GEMM forward (conv2_tc.cu) for 3x3 filters on 1- or 3-channel images, F in {32, 64, 128}
bool conv2_fwd_tc_supported(const float* out, int C, int F, int KH, int KW);
bool conv2_dimg_tc_supported(const float* dout, int C, int F, int KH, int KW);
void launch_conv2_dimg_tc(Context& ctx, const float* dout, const float* w, float* dimg, int N, int H, int W, int C, int F,
                          int KH, int KW, bool accumulate, cudaStream_t st);
bool conv2_dw_tc_supported(const float* dout, int C, int F, int KH, int KW);
void launch_conv2_dw_tc(Context& ctx, const float* img, const float* dout, float* dw, int N, int H, int W, int C, int F,
                        int KH, int KW, cudaStream_t st);
void launch_conv2_fwd_tc(Context& ctx, const float* img, const float* w, float* out, int N, int H, int W, int C, int F,
                         int KH, int KW, bool accumulate, cudaStream_t st);

// Large operands whose stored orientation is MN-major are better served by one transposing split
// pass plus the K-major tensor-core path (measured on 4096^3: 0.330 ms vs 0.353 ms per GEMM).
inline bool prefer_transposed_copy(int64_t k_extent, int64_t mn_extent) {
  return k_extent * mn_extent >= (int64_t)1 << 22;
}

}  // namespace egb
