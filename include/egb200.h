/* egb200.h - C ABI of libegb200.so, the B200 (sm_100a) execution backend for exprgrad's
 * compiled hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns an int status:
 * 0 = ok, non-zero = one of EGB_ERR_* with a human-readable message in egb_last_error()
 * (thread-local). The Nim side turns a non-zero status into `raise GpuError(msg: ...)` the way
 * exprgrad/runtimes/cl.nim:41-43 does for OpenCL status codes (RuntimeError / ShapeError for the
 * matching EGB_ERR_* codes, exprgrad/model.nim:358-359, 395-396; exprgrad/passes.nim:1393-1403).
 *
 * Three groups of functions, from the bottom up:
 *   1. device runtime   - one-to-one with the backend-neutral stub list in
 *                         exprgrad/runtimes/gpu.nim:25-52 (what `cl.nim` implements for OpenCL).
 *   2. operator kernels - the hand-written CUDA kernels, callable on raw device buffers
 *                         (what a `GpuKernel` launch resolves to).
 *   3. program / model  - the replacement of the JIT'd `target_<name>(model*)` function plus the
 *                         CompileGpu branches of exprgrad/model.nim:302-383, 392-454: the Nim side
 *                         hands over its `Program` (tensors, targets, structured kernels as in
 *                         exprgrad/ir.nim:211-270) once; call/apply/fit then run entirely on device.
 *
 * Threading / sync contract (same as the reference's single in-order command queue,
 * exprgrad/runtimes/cl.nim:92, 114-131, 199-207): one CUDA stream per context; kernel launches and
 * fills are asynchronous; write/read are blocking and are the sync points; a context is used from
 * one host thread at a time.
 */
#ifndef EGB200_H
#define EGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------- */
#define EGB_OK 0
#define EGB_ERR_GPU 1       /* GpuError       (exprgrad/runtimes/cl.nim:18)  */
#define EGB_ERR_RUNTIME 2   /* RuntimeError   (exprgrad/ir.nim:27)           */
#define EGB_ERR_SHAPE 3     /* ShapeError     (exprgrad/ir.nim:28)           */
#define EGB_ERR_PARSER 4    /* ParserError    (exprgrad/ir.nim:21)           */
#define EGB_ERR_GRADIENT 5  /* GradientError  (exprgrad/ir.nim:23)           */
#define EGB_ERR_GENERATOR 6 /* GeneratorError (exprgrad/ir.nim:24)           */
#define EGB_ERR_VALUE 7     /* ValueError (linear-system solver, passes.nim:1262-1296) */

#define EGB_MAX_RANK 8

typedef struct egb_context egb_context;
typedef struct egb_buffer egb_buffer;
typedef struct egb_kernel egb_kernel;
typedef struct egb_program egb_program;
typedef struct egb_model egb_model;
typedef struct egb_comm egb_comm;

/* Message of the last failing call on this thread. Never NULL. */
const char* egb_last_error(void);
/* "egb200 <version> sm_100a" */
const char* egb_version(void);

/* ---- 1. device runtime (exprgrad/runtimes/gpu.nim:25-52, cl.nim:83-207) ---------------------- */

/* listDevices (gpu.nim:32, cl.nim:63-65): number of CUDA devices; 0 (and EGB_OK) when none. */
int egb_device_count(int* count);
/* name/vendor/version/isGpu (gpu.nim:33-36, cl.nim:74-81). Strings are copied into caller buffers. */
int egb_device_name(int device, char* out, size_t cap);
int egb_device_vendor(int device, char* out, size_t cap);
int egb_device_version(int device, char* out, size_t cap);
int egb_device_is_gpu(int device, int* is_gpu);

/* newGpuContext(device) / newGpuContext() (gpu.nim:37-38, cl.nim:83-99); device < 0 = first. */
int egb_context_create(int device, egb_context** out);
int egb_context_destroy(egb_context* ctx);
/* Block until everything queued on the context's stream has finished. */
int egb_context_synchronize(egb_context* ctx);
/* The context's cudaStream_t (for callers that want to time with CUDA events). */
void* egb_context_stream(egb_context* ctx);
/* Number of CUDA kernels launched through this context so far (graph replays count their nodes). */
int64_t egb_context_launch_count(egb_context* ctx);

/* Context options: "pdl" 0/1 - programmatic dependent launch of consecutive kernels (default 1; the
 * environment variable EGB_PDL overrides the default). */
int egb_context_set_option(egb_context* ctx, const char* key, int64_t value);
/* Device timing for benchmarks. set_timing(1) clears the record and brackets every subsequent kernel
 * launch with CUDA events on the context's stream; kernel_time sums the device time of the launches
 * of one kernel class (EGB_KC_*, -1 = all) recorded since. */
#define EGB_KC_GEMM 0
#define EGB_KC_SPLIT 1
#define EGB_KC_FILL 2
#define EGB_KC_INTERP 3
#define EGB_KC_REDUCE 4
#define EGB_KC_ELTWISE 5
#define EGB_KC_CONV 6       /* conv2 forward */
#define EGB_KC_OTHER 7
#define EGB_KC_CONV_DW 8    /* conv2 d_filters */
#define EGB_KC_CONV_DIMG 9  /* conv2 d_images */
#define EGB_KC_EXCHANGE 10  /* data-parallel gradient exchange (+ fused optimizer update) */
int egb_context_set_timing(egb_context* ctx, int enabled);
int egb_context_kernel_time(egb_context* ctx, int kernel_class, double* total_ms, int64_t* launches);
/* CUDA events on the context's stream (cudaEvent_t behind void*). */
int egb_event_create(egb_context* ctx, void** out);
int egb_event_record(egb_context* ctx, void* event);
int egb_event_elapsed_ms(void* start, void* stop, double* ms); /* waits for `stop` */
int egb_event_destroy(void* event);
/* Page-locked host memory for Tensor[T] storage that is copied to/from the device
 * (exprgrad/tensors.nim:36-49 allocates from the host heap; pinned memory makes write/readInto run
 * at full PCIe rate). */
int egb_host_alloc(size_t bytes, void** out);
int egb_host_free(void* p);

/* allocBuffer(ctx, size) (gpu.nim:39, cl.nim:101-106) and dealloc(buffer) (cl.nim:108-109). */
int egb_alloc_buffer(egb_context* ctx, size_t bytes, egb_buffer** out);
int egb_buffer_free(egb_buffer* buf);
size_t egb_buffer_size(const egb_buffer* buf);
void* egb_buffer_device_ptr(egb_buffer* buf);
/* write(buffer, data, size) (gpu.nim:40, cl.nim:111-116): blocking H2D; size must equal the buffer
 * size, otherwise EGB_ERR_GPU "Attempted to write N bytes, but the size of the buffer is M bytes". */
int egb_buffer_write(egb_buffer* buf, const void* data, size_t bytes);
/* fill[T](buffer, value) (gpu.nim:42, cl.nim:122-126): asynchronous pattern fill, elem_size in {1,2,4,8}. */
int egb_buffer_fill(egb_buffer* buf, const void* value, size_t elem_size);
/* readInto(buffer, ptr) (gpu.nim:43-44, cl.nim:128-138): blocking D2H; bytes must equal buffer size
 * ("Buffer size is not equal to target size"). */
int egb_buffer_read_into(egb_buffer* buf, void* data, size_t bytes);

/* compile(ctx, name, source) / compile(ctx, GpuKernelSource) (gpu.nim:46-47, cl.nim:149-179). `source`
 * is not OpenCL C: it is the text of a compiled program (same grammar as egb_program_parse) and `name`
 * selects the target whose kernels this GpuKernel launches; build problems are reported like the
 * reference's "Failed to build program: <log>" (cl.nim:163-171). */
int egb_compile(egb_context* ctx, const char* name, const char* source, egb_kernel** out);
int egb_kernel_free(egb_kernel* k);
/* Number of tensor arguments (= the target's tensors, in `target.tensors` order). */
int egb_kernel_arg_count(egb_kernel* k, int* count);
/* arg(kernel, index, buffer) (gpu.nim:49, cl.nim:186-188): binds tensor argument `index`. */
int egb_kernel_arg_buffer(egb_kernel* k, int index, egb_buffer* buf);
/* The reference hands shapes to its OpenCL kernels as captured Index registers (llvmgen.nim:474-486);
 * here the shape of tensor argument `index` is attached directly. */
int egb_kernel_arg_shape(egb_kernel* k, int index, int rank, const int64_t* dims);
/* arg[T](kernel, index, value) (gpu.nim:48, cl.nim:181-184): only index -1 (the epoch) is accepted. */
int egb_kernel_arg_index(egb_kernel* k, int index, int64_t value);
/* run(kernel, groupSize, localSize) (gpu.nim:50, cl.nim:190-207): asynchronous; the launch geometry is
 * chosen by the library, the arguments are validated like the reference ("Group size must have at least
 * one dimension"). Every kernel accumulates into its destination (`+=`, clgen.nim:116-125), so results
 * must be zero-filled by the caller exactly as model.nim:302-318 does. */
int egb_kernel_run(egb_kernel* k, int work_dims, const int64_t* group_size, const int64_t* local_size);

/* ---- 2. operator kernels on raw device pointers ---------------------------------------------- */

/* C[M,N] (+)= alpha * op(A)[M,K] * op(B)[K,N] (+ bias[n]) (relu). fp32 in/out, row-major.
 * trans_a: A is stored [K,M]; trans_b: B is stored [N,K]. flags: 1 accumulate into C, 2 add bias,
 * 4 relu. Replaces the contraction loop nests of llvmgen.nim:277-297 for kernels shaped like
 * exprgrad/layers/base.nim:27-28 and their adjoints (passes.nim:519-549). tcgen05 bf16x3. */
int egb_gemm_f32(egb_context* ctx, int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int flags,
                 const float* bias, float alpha);
/* The tile width (multiple of 32; 64 for an MN-major B) and cluster split-K factor (1, 2, 4 or 8 CTAs per output
 * tile, reduced through distributed shared memory) the library would pick for an M x N x K contraction that may
 * plan for `sms` SMs. Pure host arithmetic (no device needed): exposed so that the planner's cost model can be
 * checked without a GPU. */
int egb_gemm_plan(int64_t M, int64_t N, int64_t K, int b_mn_major, int sms, int* bn, int* cluster_k);
/* The same for the latency kernel that serves small contractions (csrc/gemm_lat.cu: split factors 1, 2, 4; a split
 * tile is at most 64 columns wide), plus the number of CTAs the launch will occupy. Host arithmetic only. */
int egb_gemm_lat_plan(int64_t M, int64_t N, int64_t K, int b_mn_major, int sms, int* bn, int* cluster_k, int* ctas);
/* Same contraction on operands already split into bf16 (hi, mid) planes. By default both are K-major
 * (A stored [M, lda], B stored [N, ldb], K contiguous); flags bit 16: A is MN-major (stored [K, lda], M
 * contiguous), bit 32: B is MN-major (stored [K, ldb], N contiguous). bn = 0 lets the library choose the
 * N tile; bits 16-23 of bn force the cluster split-K factor (probing / tests). */
int egb_gemm_planes(egb_context* ctx, int64_t M, int64_t N, int64_t K, const void* a_hi, const void* a_mid,
                    int64_t lda, const void* b_hi, const void* b_mid, int64_t ldb, float* C, int64_t ldc,
                    int flags, const float* bias, float alpha, int bn);
/* fp32 [rows, cols] -> bf16 hi/mid planes ([cols, rows] when transpose != 0). act: 0 none, 1 relu. */
int egb_split_bf16(egb_context* ctx, const float* src, int64_t rows, int64_t cols, int64_t ld, int transpose,
                   void* hi, void* mid, int64_t dst_ld, int act);

/* ---- 3. program / model --------------------------------------------------------------------- */

/* Parse the text form of an exprgrad `Program` (exprgrad/ir.nim:263-270): either the SOURCE program
 * that `toProgram` builds from the Fun graphs (exprgrad/parser.nim:404-417; stage 0) or the program
 * after exprgrad's own semantics-defining passes (model.nim:46-58; stage 1). The grammar is documented
 * in exprgrad_b200/csrc/program.cpp; INTEGRATION.md shows the Nim serialiser. */
int egb_program_parse(const char* text, size_t len, egb_program** out);
/* program.compile() up to and including sortShapeConstraints (exprgrad/model.nim:46-58): dead-code
 * elimination, index folding, read dedup, shape constraints, reverse-mode autodiff (`generate`,
 * passes.nim:558-698), dead-kernel elimination, loop bounds, independent loops, loop order.
 * No-op for a stage-1 program. ShapeError / GradientError are reported as EGB_ERR_SHAPE / _GRADIENT. */
int egb_program_compile(egb_program* program);
/* Text form of the current program state (used by the tests to compare the native passes with the
 * oracle's). Two-call protocol: *needed receives the size including the terminating NUL. */
int egb_program_serialize(egb_program* program, char* buf, size_t cap, size_t* needed);
/* One line per kernel of a target: tensors read/written and the canonical text of the kernel
 * (iteration space, accesses, value expression) that the planner matches fused kernels against. */
int egb_program_describe(egb_program* program, const char* target, char* buf, size_t cap, size_t* needed);
/* Which device kernel family runs each IR kernel of `target` at the given input shapes: "contraction",
 * "conv2 ...", "eltwise <form>" (the specialised streaming kernels that replace clgen.nim:74-190 for the fixed
 * forms of layers/base.nim and layers/dnn.nim) or "generic" (the loop-nest kernel). Host only. */
int egb_program_classify(egb_program* program, const char* target, int n_args, const char* const* names,
                         const int* ranks, const int64_t* dims, char* buf, size_t cap, size_t* needed);
/* The device program of every IR kernel of `target` at the given input shapes, one JSON object per line (host only):
 * loops (independent ones first), flattened affine accesses with tensor ids in place of addresses, the register
 * program and its literal pool - what the generic loop-nest kernel interprets for a kernel no specialised device kernel
 * takes (the role of the OpenCL source exprgrad/clgen.nim:74-257 prints). `strict` as in egb_model_set_option. The
 * tests execute these programs with an independent interpreter to check the lowering without a device. */
int egb_program_lower_dump(egb_program* program, const char* target, int n_args, const char* const* names,
                           const int* ranks, const int64_t* dims, int strict, int64_t epoch, char* buf, size_t cap,
                           size_t* needed);
int egb_program_free(egb_program* program);
int egb_program_tensor_count(egb_program* program, int* count);
/* kind: 0 result, 1 input, 2 param, 3 cache, 4 random (exprgrad/ir.nim:222-233). dims has room for
 * EGB_MAX_RANK entries; -1 = dynamic. */
int egb_program_tensor_info(egb_program* program, int tensor_id, int* kind, int* rank, int64_t* dims, char* name,
                            size_t name_cap);
int egb_program_target_output(egb_program* program, const char* target, int* tensor_id);
/* inferShapes (exprgrad/passes.nim:1386-1436): host-only, integer-exact. Inputs are given by name with
 * their shapes flattened into `dims` (ranks[i] entries each). Returns the shape of `tensor_id`
 * (0 = the target's output). Needs no GPU. */
int egb_program_infer_shapes(egb_program* program, const char* target, int n_args, const char* const* names,
                             const int* ranks, const int64_t* dims, int tensor_id, int* out_rank,
                             int64_t* out_dims);

/* newModel (exprgrad/model.nim:232-251): compiles the program if needed, allocates every parameter
 * in HBM initialised U(initRange) from `seed` and every cache zero-filled. The model keeps the
 * program alive. float64 programs are rejected (EGB_ERR_GENERATOR). */
int egb_model_create(egb_context* ctx, egb_program* program, uint64_t seed, egb_model** out);
int egb_model_free(egb_model* model);
/* Options: "strict" 1 = bit-exact mode (every kernel runs on the generic loop-nest kernel with the
 * reference's sequential accumulation order; no tensor cores), "graphs" 0 = launch eagerly instead
 * of replaying a CUDA graph, "epoch" = set model.epoch (exprgrad/model.nim:39).
 * Planner switches (default 1; an execution detail each - results stay within the parity bar): "fuse" epilogue
 * fusion behind contractions, "head" the classification head (logits contraction + softmax/crossEntropy rows +
 * first adjoint contraction) as one launch, "rowchain" runs of small row-local kernels in one launch, "eltwise"
 * specialised streaming elementwise / optimizer kernels, "concurrent" independent plan nodes on parallel graph
 * branches, "splitk" cluster split-K, "keep_intermediates" (default 0) store every fp32 intermediate,
 * "dp_peer" data parallel: fused peer-memory exchange (1) or in-graph ncclAllReduce + optimizer kernels (0). */
int egb_model_set_option(egb_model* model, const char* key, int64_t value);
int egb_model_epoch(egb_model* model, int64_t* epoch);
/* model.params / model.caches are public in the reference (exprgrad/model.nim:37-38): blocking copies
 * of one state tensor (or, for read, of any tensor of the most recent call) to / from host memory. */
int egb_model_write_tensor(egb_model* model, int tensor_id, const void* host, size_t bytes);
int egb_model_read_tensor(egb_model* model, int tensor_id, void* host, size_t bytes);
int egb_model_tensor_shape(egb_model* model, int tensor_id, int* rank, int64_t* dims);
int egb_model_tensor_device_ptr(egb_model* model, int tensor_id, void** ptr);
/* Model.call / Model.apply (exprgrad/model.nim:392-411). Inputs are passed by name; data[i] is a
 * host pointer (copied H2D on the context's stream) or, when on_device[i] != 0, a device pointer
 * that is used in place. Shapes are inferred once per distinct input-shape signature and the launch
 * plan is cached; the launches are asynchronous. out_rank/out_dims (may be NULL = apply) receive the
 * output tensor's shape (out_rank = -1 if the target has no output). Errors: unknown target/input ->
 * EGB_ERR_RUNTIME (model.nim:358-359, 395-396), shape problems -> EGB_ERR_SHAPE. */
int egb_model_call(egb_model* model, const char* target, int n_args, const char* const* names,
                   const void* const* data, const int* ranks, const int64_t* dims, const int* on_device,
                   int* out_rank, int64_t* out_dims);
/* Model.call as the reference defines it (exprgrad/model.nim:392-406): run the target and return its
 * output in HOST memory (`out_host`, `out_bytes` must equal the output's size - use
 * egb_program_infer_shapes to size it). Blocking. When the target is a single plain contraction of host
 * operands (benchmarks/matmul/matmul_gpu.nim:28-75) the call is streamed in row blocks so that the H2D
 * copies, the tensor-core work and the D2H copy overlap; every other target runs call + readOutput. */
int egb_model_call_read(egb_model* model, const char* target, int n_args, const char* const* names,
                        const void* const* data, const int* ranks, const int64_t* dims, const int* on_device,
                        void* out_host, size_t out_bytes, int* out_rank, int64_t* out_dims);
/* readOutput (exprgrad/model.nim:370-376): blocking D2H of the last call's output tensor. */
int egb_model_read_output(egb_model* model, void* dst, size_t bytes);
/* Model.fit (exprgrad/model.nim:413-454): shapes inferred once with dim 0 = batch_size, epoch += 1,
 * one plan replay per full batch (trailing partial batch dropped). The host data set is uploaded once per
 * call (double-buffered chunks of whole batches when it exceeds the staging budget) and every batch is a
 * device-side slice of it - the device form of viewFirst (exprgrad/tensors.nim:290-297); no host-to-device
 * copy happens inside the batch loop. EGB_ERR_SHAPE if an argument has fewer rows than the batches cover.
 * Blocking. */
int egb_model_fit(egb_model* model, const char* target, int n_args, const char* const* names,
                  const void* const* data, const int* ranks, const int64_t* dims, int64_t batch_size,
                  int64_t* batches_run);
/* Number of launch plans (one per target x input-shape signature, model.nim:347-355 re-allocates per call
 * instead) the model currently caches; bounded by the "plan_cache" option (least recently used evicted). */
int egb_model_plan_count(egb_model* model, int* count);
/* Human-readable node list of the most recent call's launch plan. */
int egb_model_describe_plan(egb_model* model, char* buf, size_t cap, size_t* needed);

/* ---- 4. data parallel (no counterpart in the reference, which is single-device: cl.nim:95-99) ---- */

/* One process per GPU. Rank 0 creates a 128-byte NCCL unique id and the host side distributes it
 * (torch.distributed / MPI / a file); every rank then creates its communicator on its own context. */
int egb_comm_unique_id(void* out, size_t cap);
int egb_comm_create(egb_context* ctx, const void* unique_id, int rank, int world, egb_comm** out);
int egb_comm_destroy(egb_comm* comm);
int egb_comm_info(egb_comm* comm, int* rank, int* world, int* nccl_version);
/* In-place average of n floats across all ranks on the context's stream (asynchronous). */
int egb_comm_allreduce_avg_f32(egb_comm* comm, float* device_buf, size_t n);
/* Make every target that has a backward pass data parallel: the gradients of all parameters are laid
 * out as one contiguous bucket and averaged across ranks between the last adjoint kernel and the first
 * optimizer kernel (where exprgrad/parser.nim:757-766 places the optimizer effects), inside the same CUDA
 * graph: by the library's own exchange kernel over NVLink / NVSwitch peer memory, fused with the
 * gradientDescent updates (up to 8 ranks of one node; cudaIpc mappings are exchanged when a plan is built,
 * which is a collective call - every rank builds the same target with the same per-rank shapes), or with
 * option "dp_peer" 0 by one ncclAllReduce per bucket segment + the optimizer kernels. comm = NULL turns it
 * off again. With equal-size batch shards this reproduces the reference's global-batch step. */
int egb_model_set_data_parallel(egb_model* model, egb_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* EGB200_H */
