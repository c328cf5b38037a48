// Host runtime of the B200 backend: what exprgrad/model.nim does for a CompileGpu target
// (model.nim:302-383, 392-454) plus what its JIT'd `target_<name>` function does (launch every
// kernel of the target, llvmgen.nim:455-500, 518-563) - without a JIT: every kernel of a compiled
// target is lowered once per input-shape signature into a launch plan (tcgen05 contraction nodes,
// generic loop-nest nodes, fused nodes), the plan is captured into a CUDA graph, and call/apply/fit
// replay it.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "egb_internal.hpp"
#include "interp.hpp"
#include "lower.hpp"
#include "pattern.hpp"
#include "program.hpp"

namespace egb {

// One launch of a specialised streaming map kernel (eltwise_stream.cu): out (+)= f(in0 [, in1]; p0..p3)
struct EltLaunch {
  int kind = 0;        // EltKind (pattern.hpp)
  int nreads = 0;
  float* out = nullptr;
  const float* in[2] = {nullptr, nullptr};
  float p[4] = {0, 0, 0, 0};
  int64_t n = 0, row = 0;
  bool accumulate = false;
  // fused adam (kind ELT_ADAM_FUSED): out = parameter, in[0] = gradient; first / second moment caches and the
  // literals of the three reference kernels (m: p[0..1], v: p2[0..1], step: p3[0..3])
  float *adam_m = nullptr, *adam_v = nullptr;
  float p2[2] = {0, 0}, p3[4] = {0, 0, 0, 0};
};
constexpr int ELT_ADAM_FUSED = 1000;   // node-level fusion of adam-m + adam-v + adam-step (not a single-kernel pattern)
bool eltwise_stream_supported(const EltLaunch& e);
void launch_eltwise_stream(Context& ctx, const EltLaunch& e, cudaStream_t st);

// Data-parallel gradient exchange through peer memory (exchange.cu, dist.cu)
constexpr int EX_MAX_WORLD = 8;    // ranks of one NVSwitch domain
constexpr int EX_MAX_CTAS = 128;   // CTAs of the exchange kernel (= flag slots per rank)
constexpr int EX_MAX_SEG = 32;
constexpr int EX_AREAS = 2;        // exchange kernels per plan (early part of the bucket / last gradient)     // parameters whose gradientDescent update is fused into the exchange
struct ExchangeSeg {
  long long off = 0, len = 0;      // floats, relative to the bucket
  float* param = nullptr;
  float rate = 0.0f;
  int pad = 0;
};
struct ExchangeParams {
  float* bucket[EX_MAX_WORLD];     // every rank's gradient bucket (own: local pointer, others: IPC mappings)
  uint32_t* flags[EX_MAX_WORLD];   // every rank's flag area: ready[world][ctas], done[world][ctas], epoch[ctas]
  ExchangeSeg seg[EX_MAX_SEG];
  long long n = 0;                 // floats per bucket (multiple of 4)
  int rank = 0, world = 1, nseg = 0;
  int ctas = 0;                    // CTAs of this exchange (0 = default); the same on every rank
  int backoff_ns = 0;              // sleep between two looks at a line that has not arrived yet
};
size_t exchange_flag_bytes();
size_t exchange_area_bytes(size_t bucket_bytes, int world);
void launch_exchange(Context& ctx, const ExchangeParams& p, cudaStream_t st);

// Peer mappings of one plan's arena and flag area on every rank (cudaIpc handles exchanged once per plan)
struct CommHooks;
struct PeerWindow {
  bool mapped = false;
  CommHooks* owner = nullptr;          // communicator the window was opened on (teardown barrier)
  float* barrier_buf = nullptr;
  int world = 1, rank = 0;
  char* arena[EX_MAX_WORLD] = {};      // base of rank r's arena as seen from this process
  uint32_t* flags[EX_MAX_WORLD] = {};
  uint32_t* local_flags = nullptr;     // this rank's flag + inbox areas (owned): EX_AREAS exchanges per plan
  size_t area_stride = 0;              // bytes between the areas of two exchange nodes
  void* opened[2 * EX_MAX_WORLD] = {}; // IPC mappings to close
  int nopened = 0;
};

// The classification head as one row kernel (head_rows.cu): skinny contraction -> softmax + crossEntropy rows ->
// skinny adjoint contraction with its fused stages
struct HeadParams {
  // forward contraction z[rows, cols] = a[rows, Kin] * w + bias_in  (operand planes of the absorbed contraction)
  const __nv_bfloat16 *a_hi = nullptr, *a_mid = nullptr;
  int lda = 0, Kin = 0;
  const __nv_bfloat16 *win_hi = nullptr, *win_mid = nullptr;
  int win_ld = 0, win_mn = 0;
  const float* bias_in = nullptr;
  float* Z = nullptr;
  // softmax + crossEntropy rows (fused_rows.cu)
  const float *Y = nullptr, *DL = nullptr;
  float *S = nullptr, *P = nullptr, *DP = nullptr, *DH = nullptr, *DS = nullptr;
  float* colsum_dz = nullptr;
  __nv_bfloat16 *dh_hi = nullptr, *dh_mid = nullptr;
  int dh_ld = 0;
  int rows = 0, cols = 0;
  // adjoint contraction g[rows, Nout] = dz[rows, cols] * w' and its fused stages (GemmArgs of the absorbed contraction)
  const __nv_bfloat16 *wout_hi = nullptr, *wout_mid = nullptr;
  int wout_ld = 0, wout_mn = 0, Nout = 0;
  const float* bias_out = nullptr;
  int epi = EPI_NONE;
  float epi_param = 0.0f;
  const float* Hm = nullptr;
  float *C = nullptr, *D = nullptr;
  int ldc = 0, flags = 0;
  float* colsum_out = nullptr;
  __nv_bfloat16 *out_hi = nullptr, *out_mid = nullptr;
  int ld_out = 0;
  float* tables = nullptr;   // fp32 tables of both weight operands (head_tables_kernel), plan-owned
  int w_early = 0;    // the tables were complete before the predecessor kernel started: fetch them before griddepcontrol.wait
  unsigned long long* trace = nullptr;   // debug timeline (Context::trace) or null
  int trace_index = 0;
};
bool head_rows_supported(const HeadParams& p);
size_t head_table_floats(const HeadParams& p);
void launch_head_tables(Context& ctx, const HeadParams& p, cudaStream_t st);
void launch_head_rows(Context& ctx, const HeadParams& p, cudaStream_t st);

struct DevTensor {
  void* ptr = nullptr;
  size_t bytes = 0;
  std::vector<int64_t> shape;
  bool owned = false;
  int64_t len() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct Node {
  enum Kind { INTERP, GEMM, SPLIT, MEMSET, RANDOM, ALLREDUCE, CONV, ROWCHAIN, SOFTMAX_XENT, ELTWISE, EXCHANGE, HEAD, HEADPREP } kind = INTERP;
  std::string label;
  // INTERP
  IpProgram ip;
  int pb = 256, rb = 1, points_fast = 1;
  bool strict = false;
  bool uses_epoch = false;
  int rsplit = 1;  // long reduction split over blockIdx.y (partials meet through atomicAdd)
  int kernel_index = -1;  // index into target.kernels (for re-lowering when the epoch changes)
  // GEMM
  GemmArgs gemm;
  // SPLIT
  const float* split_src = nullptr;
  int split_rows = 0, split_cols = 0, split_ld = 0, split_dst_ld = 0, split_act = 0;
  bool split_transpose = false;
  __nv_bfloat16 *split_hi = nullptr, *split_mid = nullptr;
  std::vector<SplitJob> split_jobs;  // non-empty: several row splits merged into this node (one launch)
  // ROWCHAIN: several row-local INTERP programs executed by one launch
  IpProgram* chain_progs = nullptr;  // device array (owned by the plan)
  int chain_n = 0, chain_slots = 0;
  int64_t chain_rows = 0;
  // SOFTMAX_XENT: H, Y, DL -> S, P, DP, DH, DS (fused_rows.cu)
  const float *sx_h = nullptr, *sx_y = nullptr, *sx_dl = nullptr;
  float *sx_s = nullptr, *sx_p = nullptr, *sx_dp = nullptr, *sx_dh = nullptr, *sx_ds = nullptr;
  int sx_rows = 0, sx_cols = 0;
  float* sx_colsum = nullptr;                                   // fused bias-gradient column sum of DH (zeroed first)
  __nv_bfloat16 *sx_out_hi = nullptr, *sx_out_mid = nullptr;    // fused operand planes of DH
  int sx_ld_out = 0;
  // HEAD: contraction + softmax/crossEntropy rows + adjoint contraction in one launch (head_rows.cu)
  HeadParams head;
  // ELTWISE: one of the specialised streaming map kernels (eltwise_stream.cu)
  EltLaunch elt;
  // EXCHANGE: peer-memory gradient exchange + fused gradientDescent (exchange.cu)
  ExchangeParams exchange;
  // CONV
  ConvPattern conv;
  const float *conv_a = nullptr, *conv_b = nullptr;
  float* conv_out = nullptr;
  bool conv_accumulate = false;
  // dependency bookkeeping for concurrent scheduling inside the CUDA graph: buffers read / written
  // (tensor id, or -(2*tensor + orientation + 1) for a bf16 operand-plane pair) and the resulting level
  std::vector<int64_t> reads, writes;
  int level = 0;
  // MEMSET / RANDOM / ALLREDUCE
  void* ptr = nullptr;
  size_t bytes = 0;
  float lo = 0, hi = 1;
  int tensor = 0;
};

struct Model;

struct KernelInfo {
  bool is_gemm = false;
  GemmPattern gemm;
  bool overwrite = false;  // first writer of a zero-initialised result that it covers completely
  bool is_conv = false;
  ConvPattern conv;
  // Epilogue fusion: kernels that follow a contraction and only post-process its output element by
  // element (bias add, relu / leakyRelu, their adjoint masks, bias-gradient column sums, the SGD
  // update) run inside the contraction's epilogue instead of as separate launches.
  int absorbed_by = -1;       // >= 0: this kernel runs inside the epilogue of that contraction kernel
  std::vector<int> absorbed;  // (contraction) kernels fused behind this one
  int bias_tensor = 0, d_tensor = 0, h_tensor = 0, colsum_tensor = 0;
  int epi = EPI_NONE;
  float epi_param = 0.0f;
  int final_tensor = 0;       // tensor that holds the final epilogue value (C or D)
  bool emit_planes = false;   // the epilogue also writes bf16 operand planes of the final value
  int bn = 0, tiles = 0;  // tile width and tile count of a contraction
  bool in_exchange = false;  // a gradientDescent update that runs inside the data-parallel exchange kernel
};

struct Plan {
  std::string target_name;
  const Target* target = nullptr;
  ShapeTable shapes;
  std::vector<std::pair<int, std::vector<int64_t>>> input_sig;
  std::map<int, DevTensor> tensors;   // plan-owned: inputs, results, random tensors (views into the arena)
  std::map<int, const void*> bound;   // input tensors currently bound to caller-owned device memory
  char* arena = nullptr;
  size_t arena_bytes = 0;
  size_t zero_bytes = 0;              // leading part of the arena that must be zeroed before every run
  std::vector<KernelInfo> info;       // one per target kernel
  size_t plane_off = 0, plane_bytes = 0;  // bf16 operand-plane region of the arena
  size_t bucket_off = 0, bucket_bytes = 0;  // contiguous parameter-gradient bucket (data parallel)
  std::vector<std::pair<size_t, size_t>> bucket_segments;  // (arena offset, bytes), in order of readiness
  int bucket_before_kernel = -1;            // the all-reduce runs right before this target kernel
  PeerWindow window;                        // data parallel: peer mappings for the fused exchange kernel
  std::vector<ExchangeSeg> exchange_segs;   // gradientDescent updates fused into it (param pointers filled at build)
  std::vector<int> exchange_seg_param;      // their parameter tensor ids
  std::vector<Node> nodes;
  std::set<int> unmaterialized;       // result tensors whose fp32 form the fused plan never stores (dead stores)
  std::vector<void*> chain_bufs;      // device copies of row-chain programs
  std::vector<std::string> notes;     // planner log: why a fusion / fast path was not taken (describe_plan prints it)
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_valid = false;
  int64_t epoch_built = -1;
  uint64_t runs = 0;
  uint64_t last_used = 0;             // Model::use_clock stamp of the last get_plan hit (LRU eviction)
  size_t launches_per_run = 0;
  ~Plan();
};

struct CommHooks;  // data-parallel extension (dist.cu)
void comm_all_reduce_avg(CommHooks* c, float* buf, size_t n, cudaStream_t st);
int comm_world(CommHooks* c);
int comm_rank(CommHooks* c);
// Exchange cudaIpc handles of (arena, flag area) with every rank and map the peers' memory. Collective: every
// rank calls it at the same point (plan construction of the same target). Checks that all ranks laid the bucket
// out identically.
void comm_open_window(CommHooks* c, Context& ctx, char* arena, size_t arena_bytes, size_t bucket_off, size_t bucket_bytes,
                      PeerWindow& w);
void comm_close_window(PeerWindow& w);

struct Model {
  Context* ctx = nullptr;
  std::shared_ptr<Program> prog;
  std::map<int, DevTensor> state;  // params + caches, persistent in HBM (model.nim:37-38)
  int64_t epoch = 0;
  uint64_t seed = 0;
  uint64_t rng_counter = 0;
  bool strict = false;      // bit-exact mode: sequential accumulation everywhere, no tensor cores
  bool use_graphs = true;
  bool fuse = true;         // epilogue fusion of contraction + elementwise / column-sum / SGD kernels
  bool concurrent = true;   // independent plan nodes run on parallel branches of the CUDA graph
  bool rowchain = true;     // runs of small row-local kernels execute in one launch
  bool eltwise = true;      // fixed elementwise / optimizer forms run on the specialised streaming kernels
  bool headfuse = true;     // skinny contractions around the softmax + crossEntropy rows join the row kernel (head_rows.cu)
  bool keep_intermediates = false;  // store every fp32 intermediate even when only a fused epilogue consumes it
  // data parallel: exchange the gradient bucket with the fused peer-memory kernel (exchange.cu); off = the
  // ncclAllReduce(avg) + separate optimizer kernels of round 1 (kept for comparison and as the > 8 rank path)
  bool dp_peer = true;
  // cluster split-K: contractions with few output tiles spread each tile's reduction over the CTAs of a
  // thread-block cluster (partial tiles meet through distributed shared memory, gemm_tcgen05.cu).
  // (An earlier global-memory variant - RED.ADD partial tiles + last-arriver epilogue - lost on the dense
  // step, 115.6 vs 111.4 us, and was removed.)
  bool splitk = true;
  // Plan cache: one plan (arena, operand planes, CUDA graph) per target x input-shape signature. Bounded: at most
  // `max_plans_per_target` signatures per target stay resident, the least recently used one is dropped first
  // (the reference frees and reallocates when a shape changes, model.nim:311-317); an allocation failure evicts
  // every other plan before giving up.
  std::vector<std::unique_ptr<Plan>> plans;
  int max_plans_per_target = 4;
  uint64_t use_clock = 0;
  void evict_plan(size_t index);
  Plan* last_plan = nullptr;
  CommHooks* comm = nullptr;
  ~Model();

  Plan& get_plan(const std::string& target, const std::vector<int>& ids,
                 const std::vector<std::vector<int64_t>>& shapes);
  void build_nodes(Plan& plan);  // (re)lower every kernel of the plan against the current pointers / epoch
  void run(Plan& plan);
};

std::unique_ptr<Model> new_model(Context& ctx, std::shared_ptr<Program> prog, uint64_t seed);

// kernels (host launchers)
void launch_interp(Context& ctx, const IpProgram& prog, int pb, int rb, int points_fast, bool strict,
                   cudaStream_t st, int rsplit = 1);
int interp_reduction_splits(const IpProgram& prog, int pb, int rb, bool strict, int sm_count);
void launch_interp_rowchain(Context& ctx, const IpProgram* dev_progs, int nprogs, int max_slots, int64_t rows,
                            cudaStream_t st);
void launch_fill_uniform(Context& ctx, float* dst, size_t n, float lo, float hi, uint64_t seed, uint64_t counter,
                         uint64_t tensor, cudaStream_t st);

// passes.cpp helpers reused by the planner
int eval_index_instrs(const std::vector<Instr>& instrs, const ShapeTable& shapes, std::map<int, int64_t>& regs,
                      int64_t epoch);
int64_t eval_linear(const LinearIndex& li, const std::map<int, int64_t>& regs);

}  // namespace egb
