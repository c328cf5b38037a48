// C-ABI group 1 + 2: device runtime (mirror of exprgrad/runtimes/gpu.nim:25-52 as implemented for
// OpenCL in exprgrad/runtimes/cl.nim:83-207) and the raw operator entry points.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <sstream>

#include "abi_common.hpp"
#include "egb_internal.hpp"

namespace egb {

static thread_local std::string g_last_error;

void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

void set_last_error(const std::string& s) { g_last_error = s; }

void* Context::ensure_scratch(size_t bytes) {
  if (bytes > scratch_bytes) {
    if (scratch) {
      EGB_CUDA(cudaStreamSynchronize(stream));
      EGB_CUDA(cudaFree(scratch));
      scratch = nullptr;
      scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 2);
    EGB_CUDA(cudaMalloc(&scratch, want));
    scratch_bytes = want;
  }
  return scratch;
}

cudaEvent_t Context::get_event() {
  if (!event_pool.empty()) {
    cudaEvent_t e = event_pool.back();
    event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  EGB_CUDA(cudaEventCreate(&e));
  return e;
}

}  // namespace egb

using namespace egb;




static void require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    fail(EGB_ERR_GPU, "Unable to find device");  // cl.nim:97-98
  }
  if (device >= n) fail(EGB_ERR_GPU, "Device %d does not exist (%d devices)", device, n);
}

extern "C" {

const char* egb_last_error(void) { return g_last_error.c_str(); }
const char* egb_version(void) { return "egb200 0.1 sm_100a"; }

int egb_device_count(int* count) {
  EGB_TRY
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  EGB_CATCH
}

static int copy_str(const std::string& s, char* out, size_t cap) {
  if (cap == 0) return EGB_OK;
  size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
  memcpy(out, s.data(), n);
  out[n] = 0;
  return EGB_OK;
}

int egb_device_name(int device, char* out, size_t cap) {
  EGB_TRY
  require_device(device);
  cudaDeviceProp prop;
  EGB_CUDA(cudaGetDeviceProperties(&prop, device));
  copy_str(prop.name, out, cap);
  EGB_CATCH
}
int egb_device_vendor(int device, char* out, size_t cap) {
  EGB_TRY
  require_device(device);
  copy_str("NVIDIA Corporation", out, cap);
  EGB_CATCH
}
int egb_device_version(int device, char* out, size_t cap) {
  EGB_TRY
  require_device(device);
  cudaDeviceProp prop;
  EGB_CUDA(cudaGetDeviceProperties(&prop, device));
  int rt = 0;
  cudaRuntimeGetVersion(&rt);
  std::ostringstream ss;
  ss << "CUDA " << rt / 1000 << "." << (rt % 1000) / 10 << " sm_" << prop.major << prop.minor;
  copy_str(ss.str(), out, cap);
  EGB_CATCH
}
int egb_device_is_gpu(int device, int* is_gpu) {
  EGB_TRY
  require_device(device);
  *is_gpu = 1;
  EGB_CATCH
}

int egb_context_create(int device, egb_context** out) {
  EGB_TRY
  if (device < 0) device = 0;
  require_device(device);
  EGB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EGB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    fail(EGB_ERR_GPU, "egb200 kernels are built for sm_100a only; device %d is sm_%d%d", device, prop.major,
         prop.minor);
  egb_context* ctx = new egb_context();
  ctx->c.device = device;
  ctx->c.sm_count = prop.multiProcessorCount;
  // The context's stream carries the critical path of every launch plan; the side streams that hold the
  // parallel branches of a captured graph keep the default (lowest) priority, so when both compete for SMs the
  // critical path's CTAs are placed first. EGB_STREAM_PRIORITY=0 restores equal priorities (measurement knob).
  {
    int least = 0, greatest = 0;
    EGB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const char* pr = getenv("EGB_STREAM_PRIORITY");
    const int prio = (pr && atoi(pr) == 0) ? least : greatest;
    EGB_CUDA(cudaStreamCreateWithPriority(&ctx->c.stream, cudaStreamNonBlocking, prio));
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  EGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) fail(EGB_ERR_GPU, "cuTensorMapEncodeTiled not found in driver");
  ctx->c.encode_tiled = (PFN_encodeTiled)fn;
  if (const char* e = getenv("EGB_PDL")) ctx->c.pdl = atoi(e) != 0;
  if (getenv("EGB_GEMM_TRACE")) {
    EGB_CUDA(cudaHostAlloc((void**)&ctx->c.trace, (size_t)TRACE_SLOTS * TRACE_SLOT_WORDS * 8, cudaHostAllocMapped));
    memset(ctx->c.trace, 0, (size_t)TRACE_SLOTS * TRACE_SLOT_WORDS * 8);
  }
  *out = ctx;
  EGB_CATCH
}

int egb_context_destroy(egb_context* ctx) {
  EGB_TRY
  if (!ctx) return EGB_OK;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  if (ctx->c.trace) {
    if (FILE* f = fopen(getenv("EGB_GEMM_TRACE") ? getenv("EGB_GEMM_TRACE") : "/dev/null", "w")) {
      const unsigned long long n = ctx->c.trace[0] < TRACE_SLOTS - 1 ? ctx->c.trace[0] : TRACE_SLOTS - 1;
      for (unsigned long long i = 1; i <= n; ++i) {
        for (int w = 0; w < TRACE_SLOT_WORDS; ++w) fprintf(f, "%llu ", ctx->c.trace[i * TRACE_SLOT_WORDS + w]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    cudaFreeHost(ctx->c.trace);
  }
  if (ctx->c.scratch) cudaFree(ctx->c.scratch);
  for (auto st : ctx->c.aux_stream)
    if (st) cudaStreamDestroy(st);
  for (auto e : ctx->c.fork_events) cudaEventDestroy(e);
  for (auto& s : ctx->c.spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  for (auto e : ctx->c.event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->c.stream);
  delete ctx;
  EGB_CATCH
}

int egb_context_synchronize(egb_context* ctx) {
  EGB_TRY
  EGB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  EGB_CATCH
}

void* egb_context_stream(egb_context* ctx) { return (void*)ctx->c.stream; }

int egb_context_set_option(egb_context* ctx, const char* key, int64_t value) {
  EGB_TRY
  const std::string k(key);
  if (k == "pdl") ctx->c.pdl = value != 0;
  else fail(EGB_ERR_GPU, "unknown context option '%s'", key);
  EGB_CATCH
}

int egb_context_set_timing(egb_context* ctx, int enabled) {
  EGB_TRY
  Context& c = ctx->c;
  EGB_CUDA(cudaStreamSynchronize(c.stream));
  for (auto& s : c.spans) {
    c.event_pool.push_back(s.a);
    c.event_pool.push_back(s.b);
  }
  c.spans.clear();
  c.timing = enabled != 0;
  EGB_CATCH
}

int egb_context_kernel_time(egb_context* ctx, int kernel_class, double* total_ms, int64_t* launches) {
  EGB_TRY
  Context& c = ctx->c;
  EGB_CUDA(cudaStreamSynchronize(c.stream));
  double ms = 0;
  int64_t n = 0;
  for (auto& s : c.spans) {
    if (kernel_class >= 0 && s.cls != kernel_class) continue;
    float t = 0;
    EGB_CUDA(cudaEventElapsedTime(&t, s.a, s.b));
    ms += t;
    ++n;
  }
  *total_ms = ms;
  *launches = n;
  EGB_CATCH
}

int egb_event_create(egb_context* ctx, void** out) {
  EGB_TRY
  (void)ctx;
  cudaEvent_t e;
  EGB_CUDA(cudaEventCreate(&e));
  *out = (void*)e;
  EGB_CATCH
}
int egb_event_record(egb_context* ctx, void* ev) {
  EGB_TRY
  EGB_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->c.stream));
  EGB_CATCH
}
int egb_event_elapsed_ms(void* start, void* stop, double* ms) {
  EGB_TRY
  EGB_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  float t = 0;
  EGB_CUDA(cudaEventElapsedTime(&t, (cudaEvent_t)start, (cudaEvent_t)stop));
  *ms = t;
  EGB_CATCH
}
int egb_event_destroy(void* ev) {
  EGB_TRY
  if (ev) EGB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  EGB_CATCH
}

int egb_host_alloc(size_t bytes, void** out) {
  EGB_TRY
  *out = nullptr;
  if (bytes) EGB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  EGB_CATCH
}
int egb_host_free(void* p) {
  EGB_TRY
  if (p) EGB_CUDA(cudaFreeHost(p));
  EGB_CATCH
}
int64_t egb_context_launch_count(egb_context* ctx) { return (int64_t)ctx->c.launches; }

int egb_alloc_buffer(egb_context* ctx, size_t bytes, egb_buffer** out) {
  EGB_TRY
  egb_buffer* b = new egb_buffer();
  b->ctx = ctx;
  b->size = bytes;
  b->ptr = nullptr;
  if (bytes > 0) {
    cudaError_t e = cudaMalloc(&b->ptr, bytes);
    if (e != cudaSuccess) {
      delete b;
      fail(EGB_ERR_GPU, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
  }
  *out = b;
  EGB_CATCH
}

int egb_buffer_free(egb_buffer* buf) {
  EGB_TRY
  if (!buf) return EGB_OK;
  if (buf->ptr) {
    EGB_CUDA(cudaStreamSynchronize(buf->ctx->c.stream));
    EGB_CUDA(cudaFree(buf->ptr));
  }
  delete buf;
  EGB_CATCH
}

size_t egb_buffer_size(const egb_buffer* buf) { return buf->size; }
void* egb_buffer_device_ptr(egb_buffer* buf) { return buf->ptr; }

int egb_buffer_write(egb_buffer* buf, const void* data, size_t bytes) {
  EGB_TRY
  if (bytes != buf->size)
    fail(EGB_ERR_GPU, "Attempted to write %zu bytes, but the size of the buffer is %zu bytes", bytes, buf->size);
  if (bytes) {
    EGB_CUDA(cudaMemcpyAsync(buf->ptr, data, bytes, cudaMemcpyHostToDevice, buf->ctx->c.stream));
    EGB_CUDA(cudaStreamSynchronize(buf->ctx->c.stream));
  }
  EGB_CATCH
}

int egb_buffer_read_into(egb_buffer* buf, void* data, size_t bytes) {
  EGB_TRY
  if (bytes != buf->size) fail(EGB_ERR_GPU, "Buffer size is not equal to target size");
  if (bytes) {
    EGB_CUDA(cudaMemcpyAsync(data, buf->ptr, bytes, cudaMemcpyDeviceToHost, buf->ctx->c.stream));
    EGB_CUDA(cudaStreamSynchronize(buf->ctx->c.stream));
  }
  EGB_CATCH
}

int egb_buffer_fill(egb_buffer* buf, const void* value, size_t elem_size) {
  EGB_TRY
  if (buf->size == 0) return EGB_OK;
  if (elem_size == 0 || buf->size % elem_size != 0)
    fail(EGB_ERR_GPU, "Buffer size is not divisible by item type size");
  cudaStream_t st = buf->ctx->c.stream;
  bool all_same = true;
  const unsigned char* v = (const unsigned char*)value;
  for (size_t i = 1; i < elem_size; ++i) all_same = all_same && v[i] == v[0];
  if (all_same) {
    EGB_CUDA(cudaMemsetAsync(buf->ptr, v[0], buf->size, st));
  } else if (elem_size == 4) {
    uint32_t w;
    memcpy(&w, value, 4);
    launch_fill_u32(buf->ctx->c, (uint32_t*)buf->ptr, w, buf->size / 4, st);
  } else if (elem_size == 2) {
    fail(EGB_ERR_GPU, "fill: 2-byte non-uniform patterns are not supported");
  } else if (elem_size == 8) {
    // two interleaved 32-bit halves: stage on host for simplicity (rare path: float64 fill)
    std::vector<uint64_t> host(buf->size / 8);
    uint64_t w;
    memcpy(&w, value, 8);
    for (auto& x : host) x = w;
    EGB_CUDA(cudaMemcpyAsync(buf->ptr, host.data(), buf->size, cudaMemcpyHostToDevice, st));
    EGB_CUDA(cudaStreamSynchronize(st));
  } else {
    fail(EGB_ERR_GPU, "fill: unsupported element size %zu", elem_size);
  }
  EGB_CATCH
}

int egb_split_bf16(egb_context* ctx, const float* src, int64_t rows, int64_t cols, int64_t ld, int transpose,
                   void* hi, void* mid, int64_t dst_ld, int act) {
  EGB_TRY
  launch_split_bf16(ctx->c, src, (int)rows, (int)cols, (int)ld, transpose != 0, (__nv_bfloat16*)hi,
                    (__nv_bfloat16*)mid, (int)dst_ld, act, ctx->c.stream);
  EGB_CATCH
}

int egb_gemm_plan(int64_t M, int64_t N, int64_t K, int b_mn_major, int sms, int* bn, int* cluster_k) {
  EGB_TRY
  if (M <= 0 || N <= 0 || K <= 0 || sms <= 0) fail(EGB_ERR_GPU, "gemm plan: sizes must be positive");
  gemm_plan((int)M, (int)N, (int)K, b_mn_major != 0, sms, 148, 8, bn, cluster_k);
  EGB_CATCH
}

int egb_gemm_lat_plan(int64_t M, int64_t N, int64_t K, int b_mn_major, int sms, int* bn, int* cluster_k, int* ctas) {
  EGB_TRY
  if (M <= 0 || N <= 0 || K <= 0 || sms <= 0) fail(EGB_ERR_GPU, "gemm plan: sizes must be positive");
  gemm_lat_plan((int)M, (int)N, (int)K, b_mn_major != 0, sms, 4, bn, cluster_k);
  if (*cluster_k > 1 && *bn > 64) *bn = 64;
  *ctas = (int)(((M + 127) / 128) * ((N + *bn - 1) / *bn)) * (*cluster_k > 1 ? *cluster_k : 1);
  if (*cluster_k == 1 && *ctas > sms) *ctas = sms;
  EGB_CATCH
}

int egb_gemm_planes(egb_context* ctx, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_mid,
                    int64_t lda, const void* b_hi, const void* b_mid, int64_t ldb, float* C, int64_t ldc,
                    int flags, const float* bias, float alpha, int bn) {
  EGB_TRY
  GemmArgs g;
  g.a_hi = (const __nv_bfloat16*)a_hi; g.a_mid = (const __nv_bfloat16*)a_mid; g.lda = (int)lda;
  g.b_hi = (const __nv_bfloat16*)b_hi; g.b_mid = (const __nv_bfloat16*)b_mid; g.ldb = (int)ldb;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.a_mn = (flags & 16) != 0;
  g.b_mn = (flags & 32) != 0;
  g.C = C; g.ldc = (int)ldc; g.flags = flags & 3; g.bias = bias; g.alpha = alpha; g.bn = bn & 0xffff;
  g.cluster_k = (bn >> 16) & 0xff;  // probing: bn | cluster split-K factor << 16
  if (flags & 4) { g.epi = EPI_RELU; g.D = C; }
  launch_gemm_bf16x3(ctx->c, g, ctx->c.stream);
  EGB_CATCH
}

int egb_gemm_f32(egb_context* ctx, int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int flags,
                 const float* bias, float alpha) {
  EGB_TRY
  if (M <= 0 || N <= 0) return EGB_OK;
  if (K <= 0) fail(EGB_ERR_GPU, "gemm: K must be positive");
  Context& c = ctx->c;
  // bf16 planes of both operands. The tensor-core descriptors take either orientation (K-major or
  // MN-major), so small operands are split exactly as stored; large MN-major ones get a transposed copy.
  const bool a_copy = trans_a && prefer_transposed_copy(K, M);
  const bool b_copy = !trans_b && prefer_transposed_copy(K, N);
  const int64_t a_rows = trans_a ? K : M, a_cols = trans_a ? M : K;    // as stored
  const int64_t b_rows = trans_b ? N : K, b_cols = trans_b ? K : N;
  const int64_t a_prow = a_copy ? a_cols : a_rows, a_pcol = a_copy ? a_rows : a_cols;  // plane shape
  const int64_t b_prow = b_copy ? b_cols : b_rows, b_pcol = b_copy ? b_rows : b_cols;
  const int64_t a_ld = (a_pcol + 7) & ~int64_t(7), b_ld = (b_pcol + 7) & ~int64_t(7);
  const size_t a_plane = (size_t)a_prow * a_ld * 2, b_plane = (size_t)b_prow * b_ld * 2;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t planes_total = 2 * al(a_plane) + 2 * al(b_plane);
  const size_t tail_ws = al(gemm_2cta_workspace_bytes(c.sm_count));
  const size_t scratch_before = c.scratch_bytes;
  char* ws = (char*)c.ensure_scratch(planes_total + tail_ws);
  if (c.scratch_bytes != scratch_before) c.scratch_tail_ws_zeroed = false;   // a new allocation: flags are garbage
  __nv_bfloat16* a_hi = (__nv_bfloat16*)ws;
  __nv_bfloat16* a_mid = (__nv_bfloat16*)(ws + al(a_plane));
  __nv_bfloat16* b_hi = (__nv_bfloat16*)(ws + 2 * al(a_plane));
  __nv_bfloat16* b_mid = (__nv_bfloat16*)(ws + 2 * al(a_plane) + al(b_plane));
  launch_split_bf16(c, A, (int)a_rows, (int)a_cols, (int)lda, a_copy, a_hi, a_mid, (int)a_ld, 0, c.stream);
  launch_split_bf16(c, B, (int)b_rows, (int)b_cols, (int)ldb, b_copy, b_hi, b_mid, (int)b_ld, 0, c.stream);
  GemmArgs g;
  g.a_hi = a_hi; g.a_mid = a_mid; g.lda = (int)a_ld; g.a_mn = trans_a != 0 && !a_copy;  // stored [K, M]
  g.b_hi = b_hi; g.b_mid = b_mid; g.ldb = (int)b_ld; g.b_mn = trans_b == 0 && !b_copy;  // stored [K, N]
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.C = C; g.ldc = (int)ldc; g.flags = flags & 3; g.bias = bias; g.alpha = alpha;
  if (flags & 4) { g.epi = EPI_RELU; g.D = C; }
  if (gemm_2cta_eligible(g)) {
    // tail-wave split workspace behind the planes (its 4 KB of flags must start out as zero)
    g.ws = ws + planes_total;
    g.ws_bytes = tail_ws;
    if (!c.scratch_tail_ws_zeroed || c.scratch_tail_ws != g.ws) {
      EGB_CUDA(cudaMemsetAsync(g.ws, 0, 4096, c.stream));
      c.scratch_tail_ws_zeroed = true;
      c.scratch_tail_ws = g.ws;
    }
  }
  launch_gemm_bf16x3(c, g, c.stream);
  EGB_CATCH
}

}  // extern "C"
