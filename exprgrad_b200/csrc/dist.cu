// Data-parallel extension: one process per GPU, full parameter replica per rank, the
// batch dimension sharded by the caller; the only exchange step of the train target is one all-reduce
// (average) of the contiguous parameter-gradient bucket, placed between the last adjoint kernel and
// the first optimizer kernel (the reference has no multi-device path; its gradient kernels reduce over
// the batch index, exprgrad/passes.nim:519-549, so averaging the per-shard gradients of equal-size
// shards reproduces the global-batch step, base.nim:66-67 normalising by the local shape[0]).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy torch has already loaded when the host
// side uses torch.distributed for rendezvous, otherwise the system one), so libegb200.so carries no
// link-time dependency on it. The communicator runs on the context's stream and is captured into the
// plan's CUDA graph together with the compute kernels.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "abi_model.hpp"

namespace egb {

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[128];
};
typedef int (*PFN_ncclGetUniqueId)(NcclUniqueId*);
typedef int (*PFN_ncclCommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*PFN_ncclCommDestroy)(ncclComm_t);
typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*PFN_ncclAllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef const char* (*PFN_ncclGetErrorString)(int);
typedef int (*PFN_ncclGetVersion)(int*);
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4, kNcclInt8 = 0;

struct NcclApi {
  void* handle = nullptr;
  PFN_ncclGetUniqueId get_unique_id = nullptr;
  PFN_ncclCommInitRank comm_init_rank = nullptr;
  PFN_ncclCommDestroy comm_destroy = nullptr;
  PFN_ncclAllReduce all_reduce = nullptr;
  PFN_ncclAllGather all_gather = nullptr;
  PFN_ncclGetErrorString error_string = nullptr;
  PFN_ncclGetVersion get_version = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) fail(EGB_ERR_GPU, "unable to load libnccl.so.2: %s", dlerror());
  auto sym = [&](const char* name) {
    void* p = dlsym(api.handle, name);
    if (!p) fail(EGB_ERR_GPU, "libnccl is missing symbol %s", name);
    return p;
  };
  api.get_unique_id = (PFN_ncclGetUniqueId)sym("ncclGetUniqueId");
  api.comm_init_rank = (PFN_ncclCommInitRank)sym("ncclCommInitRank");
  api.comm_destroy = (PFN_ncclCommDestroy)sym("ncclCommDestroy");
  api.all_reduce = (PFN_ncclAllReduce)sym("ncclAllReduce");
  api.all_gather = (PFN_ncclAllGather)sym("ncclAllGather");
  api.error_string = (PFN_ncclGetErrorString)sym("ncclGetErrorString");
  api.get_version = (PFN_ncclGetVersion)sym("ncclGetVersion");
  return api;
}

void nccl_check(int rc, const char* what) {
  if (rc != 0) fail(EGB_ERR_GPU, "%s failed: %s", what, nccl().error_string(rc));
}

}  // namespace

struct CommHooks {
  Context* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  void all_reduce_avg(float* buf, size_t n, cudaStream_t st) {
    if (world <= 1 || n == 0) return;
    nccl_check(nccl().all_reduce(buf, buf, n, kNcclFloat32, kNcclAvg, comm, st), "ncclAllReduce");
  }
};

void comm_all_reduce_avg(CommHooks* c, float* buf, size_t n, cudaStream_t st) { c->all_reduce_avg(buf, n, st); }
int comm_world(CommHooks* c) { return c ? c->world : 1; }
int comm_rank(CommHooks* c) { return c ? c->rank : 0; }

// ---- peer windows for the fused exchange kernel (exchange.cu) ----------------------------------------------
// Every rank exports its plan arena (it contains the gradient bucket) and a small flag area as cudaIpc handles;
// one ncclAllGather of the 160-byte records distributes them, cudaIpcOpenMemHandle maps the peers' memory into
// this process (NVLink peer access is enabled lazily by the mapping). One process per GPU, one node.
namespace {
struct WindowRecord {
  cudaIpcMemHandle_t arena, flags;   // 64 bytes each
  uint64_t arena_bytes, bucket_off, bucket_bytes, magic;
};
}  // namespace

void comm_open_window(CommHooks* c, Context& ctx, char* arena, size_t arena_bytes, size_t bucket_off, size_t bucket_bytes,
                      PeerWindow& w) {
  const int world = c->world, rank = c->rank;
  if (world > EX_MAX_WORLD) fail(EGB_ERR_GPU, "peer exchange supports at most %d ranks", EX_MAX_WORLD);
  w.world = world;
  w.rank = rank;
  w.owner = c;
  EGB_CUDA(cudaMalloc((void**)&w.barrier_buf, 256));
  EGB_CUDA(cudaMemsetAsync(w.barrier_buf, 0, 256, ctx.stream));
  w.area_stride = (exchange_area_bytes(bucket_bytes, world) + 255) & ~(size_t)255;   // flags + inboxes of one exchange kernel
  const size_t fbytes = w.area_stride * EX_AREAS;
  EGB_CUDA(cudaMalloc((void**)&w.local_flags, fbytes));
  EGB_CUDA(cudaMemsetAsync(w.local_flags, 0, fbytes, ctx.stream));
  for (int r = 0; r < EX_MAX_WORLD; ++r) {
    w.arena[r] = nullptr;
    w.flags[r] = nullptr;
  }
  w.arena[rank] = arena;
  w.flags[rank] = w.local_flags;
  if (world > 1) {
    WindowRecord mine;
    memset(&mine, 0, sizeof(mine));
    EGB_CUDA(cudaIpcGetMemHandle(&mine.arena, arena));
    EGB_CUDA(cudaIpcGetMemHandle(&mine.flags, w.local_flags));
    mine.arena_bytes = arena_bytes;
    mine.bucket_off = bucket_off;
    mine.bucket_bytes = bucket_bytes;
    mine.magic = 0x45474258ull;  // "EGBX"
    char* dev = nullptr;
    EGB_CUDA(cudaMalloc((void**)&dev, sizeof(WindowRecord) * (size_t)(world + 1)));
    std::vector<WindowRecord> all((size_t)world);
    cudaError_t e = cudaMemcpyAsync(dev, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx.stream);
    int rc = 0;
    if (e == cudaSuccess)
      rc = nccl().all_gather(dev, dev + sizeof(WindowRecord), sizeof(WindowRecord), kNcclInt8, c->comm, ctx.stream);
    if (e == cudaSuccess && rc == 0)
      e = cudaMemcpyAsync(all.data(), dev + sizeof(WindowRecord), sizeof(WindowRecord) * (size_t)world, cudaMemcpyDeviceToHost,
                          ctx.stream);
    if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(ctx.stream);
    cudaFree(dev);
    if (rc != 0) nccl_check(rc, "ncclAllGather (peer window handles)");
    EGB_CUDA(e);
    for (int r = 0; r < world; ++r) {
      if (all[(size_t)r].magic != mine.magic || all[(size_t)r].bucket_off != mine.bucket_off ||
          all[(size_t)r].bucket_bytes != mine.bucket_bytes)
        fail(EGB_ERR_GPU, "data parallel: rank %d laid out its gradient bucket differently (offset %llu, %llu bytes; here %llu, %llu)"
                          " - every rank must build the same target with the same per-rank shapes",
             r, (unsigned long long)all[(size_t)r].bucket_off, (unsigned long long)all[(size_t)r].bucket_bytes,
             (unsigned long long)mine.bucket_off, (unsigned long long)mine.bucket_bytes);
      if (r == rank) continue;
      void *pa = nullptr, *pf = nullptr;
      EGB_CUDA(cudaIpcOpenMemHandle(&pa, all[(size_t)r].arena, cudaIpcMemLazyEnablePeerAccess));
      w.opened[w.nopened++] = pa;
      EGB_CUDA(cudaIpcOpenMemHandle(&pf, all[(size_t)r].flags, cudaIpcMemLazyEnablePeerAccess));
      w.opened[w.nopened++] = pf;
      w.arena[r] = (char*)pa;
      w.flags[r] = (uint32_t*)pf;
    }
  } else {
    EGB_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  w.mapped = true;
}

void comm_close_window(PeerWindow& w) {
  if (w.local_flags && getenv("EGB_EXCHANGE_TRACE")) {
    // phase stamps of the last exchange launch (CTA 0 and the last CTA): start, after A, B, C, D (exchange.cu)
    unsigned long long t[16];
    const size_t off = (size_t)(2 * EX_MAX_WORLD * EX_MAX_CTAS + EX_MAX_CTAS) * sizeof(uint32_t);
    for (int area = 0; area < EX_AREAS; ++area)
    if (cudaMemcpy(t, (char*)w.local_flags + area * w.area_stride + off, sizeof(t), cudaMemcpyDeviceToHost) == cudaSuccess && t[0])
      for (int q = 0; q < 2; ++q)
        fprintf(stderr, "egb exchange %d trace rank %d %s: push %.2f us, reduce + push of the average %.2f us, update %.2f us (start %llu)\n",
                area, w.rank, q == 0 ? "cta 0" : "last cta", (t[8 * q + 1] - t[8 * q]) / 1e3, (t[8 * q + 2] - t[8 * q + 1]) / 1e3,
                (t[8 * q + 3] - t[8 * q + 2]) / 1e3, t[8 * q]);
  }
  for (int i = 0; i < w.nopened; ++i)
    if (w.opened[i]) cudaIpcCloseMemHandle(w.opened[i]);
  w.nopened = 0;
  // An exported allocation must outlive every mapping of it: the ranks tear a data-parallel plan down together
  // (as they built it), so a tiny all-reduce serves as the barrier "everybody has closed my arena" before the
  // caller frees it. (Skipped when the communicator is already gone.)
  if (w.owner && w.world > 1 && w.owner->comm && w.barrier_buf) {
    cudaStream_t st = w.owner->ctx->stream;
    if (nccl().all_reduce(w.barrier_buf, w.barrier_buf, 1, kNcclFloat32, kNcclSum, w.owner->comm, st) == 0)
      cudaStreamSynchronize(st);
  }
  if (w.barrier_buf) cudaFree(w.barrier_buf);
  w.barrier_buf = nullptr;
  if (w.local_flags) cudaFree(w.local_flags);
  w.local_flags = nullptr;
  w.mapped = false;
}

}  // namespace egb

using namespace egb;

struct egb_comm {
  CommHooks h;
};

extern "C" {

int egb_comm_unique_id(void* out, size_t cap) {
  EGB_TRY
  if (cap < 128) fail(EGB_ERR_GPU, "unique id buffer must hold 128 bytes");
  NcclUniqueId id;
  nccl_check(nccl().get_unique_id(&id), "ncclGetUniqueId");
  memcpy(out, &id, 128);
  EGB_CATCH
}

int egb_comm_create(egb_context* ctx, const void* unique_id, int rank, int world, egb_comm** out) {
  EGB_TRY
  if (world < 1 || rank < 0 || rank >= world) fail(EGB_ERR_GPU, "invalid rank %d of %d", rank, world);
  EGB_CUDA(cudaSetDevice(ctx->c.device));
  auto c = new egb_comm();
  c->h.ctx = &ctx->c;
  c->h.rank = rank;
  c->h.world = world;
  if (world > 1) {
    NcclUniqueId id;
    memcpy(&id, unique_id, 128);
    int rc = nccl().comm_init_rank(&c->h.comm, world, id, rank);
    if (rc != 0) {
      delete c;
      nccl_check(rc, "ncclCommInitRank");
    }
  }
  *out = c;
  EGB_CATCH
}

int egb_comm_destroy(egb_comm* comm) {
  EGB_TRY
  if (!comm) return EGB_OK;
  if (comm->h.comm) {
    cudaStreamSynchronize(comm->h.ctx->stream);
    nccl().comm_destroy(comm->h.comm);
  }
  delete comm;
  EGB_CATCH
}

int egb_comm_info(egb_comm* comm, int* rank, int* world, int* nccl_version) {
  EGB_TRY
  if (rank) *rank = comm->h.rank;
  if (world) *world = comm->h.world;
  if (nccl_version) {
    *nccl_version = 0;
    if (comm->h.world > 1) nccl().get_version(nccl_version);
  }
  EGB_CATCH
}

int egb_comm_allreduce_avg_f32(egb_comm* comm, float* device_buf, size_t n) {
  EGB_TRY
  comm->h.all_reduce_avg(device_buf, n, comm->h.ctx->stream);
  EGB_CATCH
}

}  // extern "C"

extern "C" int egb_model_set_data_parallel(egb_model* model, egb_comm* comm) {
  EGB_TRY
  Model& m = *model->m;
  EGB_CUDA(cudaStreamSynchronize(model->ctx->c.stream));
  m.plans.clear();  // plans are rebuilt with (or without) the gradient-bucket all-reduce node
  m.last_plan = nullptr;
  m.comm = comm ? &comm->h : nullptr;
  EGB_CATCH
}
