// C-ABI group 3: program + model. Replaces, for CompileGpu targets, exprgrad's
//   compile[T]            exprgrad/model.nim:270-273  (egb_program_parse/compile + egb_model_create)
//   Model.call/apply      exprgrad/model.nim:392-411  (egb_model_call [+ egb_model_read_output])
//   Model.fit             exprgrad/model.nim:413-454  (egb_model_fit)
//   model.params/caches   exprgrad/model.nim:37-38    (egb_model_write_tensor / read_tensor)
//   inferShapes           exprgrad/passes.nim:1386-1436 (egb_program_infer_shapes)
#include <stdlib.h>
#include <string.h>

#include "abi_model.hpp"

using namespace egb;

namespace {

int copy_out(const std::string& s, char* buf, size_t cap, size_t* needed) {
  if (needed) *needed = s.size() + 1;
  if (buf && cap) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return EGB_OK;
}

struct Args {
  std::vector<int> ids;
  std::vector<std::vector<int64_t>> shapes;
};

Args resolve_args(const Program& prog, int n_args, const char* const* names, const int* ranks, const int64_t* dims) {
  Args a;
  size_t off = 0;
  for (int i = 0; i < n_args; ++i) {
    auto it = prog.inputs.find(names[i]);
    if (it == prog.inputs.end()) fail(EGB_ERR_RUNTIME, "%s is not an input to the model", names[i]);
    a.ids.push_back(it->second);
    std::vector<int64_t> shape(dims + off, dims + off + ranks[i]);
    off += ranks[i];
    a.shapes.push_back(shape);
  }
  return a;
}

int64_t shape_len(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

}  // namespace

extern "C" {

int egb_program_parse(const char* text, size_t len, egb_program** out) {
  EGB_TRY
  auto p = new egb_program();
  try {
    p->p = parse_program(std::string(text, len));
  } catch (...) {
    delete p;
    throw;
  }
  *out = p;
  EGB_CATCH
}

int egb_program_compile(egb_program* p) {
  EGB_TRY
  compile_program(*p->p);
  EGB_CATCH
}

int egb_program_serialize(egb_program* p, char* buf, size_t cap, size_t* needed) {
  EGB_TRY
  copy_out(serialize_program(*p->p), buf, cap, needed);
  EGB_CATCH
}

int egb_program_describe(egb_program* p, const char* target, char* buf, size_t cap, size_t* needed) {
  EGB_TRY
  Target* t = p->p->find_target(target);
  if (!t) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  std::string s;
  for (size_t i = 0; i < t->kernels.size(); ++i) {
    const Kernel& k = *t->kernels[i];
    s += std::to_string(i) + ": T" + std::to_string(k.write.tensor) + " <-";
    for (auto& r : k.reads) s += " T" + std::to_string(r.tensor);
    s += " | " + describe_kernel(k) + "\n";
  }
  copy_out(s, buf, cap, needed);
  EGB_CATCH
}

// Which device kernel family runs each IR kernel of a target at the given input shapes (host only: shape
// inference + the structural matchers the planner uses). One line per kernel:
//   "<index>: contraction M N K transA transB" | "conv2 forward|d_filters|d_images" |
//   "eltwise <form> n=<elements> [row=<len>]" | "generic"
int egb_program_classify(egb_program* p, const char* target, int n_args, const char* const* names, const int* ranks,
                         const int64_t* dims, char* buf, size_t cap, size_t* needed) {
  EGB_TRY
  Program& prog = *p->p;
  if (!prog.compiled) compile_program(prog);
  Target* t = prog.find_target(target);
  if (!t) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  Args a = resolve_args(prog, n_args, names, ranks, dims);
  ShapeTable in;
  for (size_t i = 0; i < a.ids.size(); ++i) in[a.ids[i]] = a.shapes[i];
  ShapeTable shapes = infer_shapes(prog, *t, in);
  for (int id : prog.params) shapes[id] = prog.tdef(id).shape;
  for (int id : prog.caches) shapes[id] = prog.tdef(id).shape;
  std::string s;
  for (size_t i = 0; i < t->kernels.size(); ++i) {
    const Kernel& k = *t->kernels[i];
    GemmPattern g;
    ConvPattern cv;
    EltSpec es;
    s += std::to_string(i) + ": ";
    if (match_gemm(k, shapes, g)) {
      s += "contraction " + std::to_string(g.M) + " " + std::to_string(g.N) + " " + std::to_string(g.K) + " " +
           (g.trans_a ? "T" : "N") + (g.trans_b ? "T" : "N");
    } else if (match_conv2(k, shapes, cv)) {
      static const char* kinds[] = {"forward", "d_filters", "d_images"};
      s += std::string("conv2 ") + kinds[(int)cv.kind];
    } else if (match_eltwise(k, shapes, es)) {
      s += std::string("eltwise ") + elt_kind_name(es.kind) + " n=" + std::to_string(es.n);
      if (es.kind == ELT_BIAS_ROW) s += " row=" + std::to_string(es.row);
    } else {
      s += "generic";
    }
    s += "\n";
  }
  copy_out(s, buf, cap, needed);
  EGB_CATCH
}

// The device program of every IR kernel of a target, as text (host only: shape inference + lower.cpp). What the generic
// loop-nest kernel (interp.cu) would execute is DATA - loops, flattened affine accesses, a register program, literals -
// so it can be checked without a device: tests/test_lowering_cpu.py runs these programs with an independent
// interpreter and compares the results with the oracle. Tensor bases are written as tensor ids. One JSON object per line.
int egb_program_lower_dump(egb_program* p, const char* target, int n_args, const char* const* names, const int* ranks,
                           const int64_t* dims, int strict, int64_t epoch, char* buf, size_t cap, size_t* needed) {
  EGB_TRY
  Program& prog = *p->p;
  if (!prog.compiled) compile_program(prog);
  Target* t = prog.find_target(target);
  if (!t) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  Args a = resolve_args(prog, n_args, names, ranks, dims);
  ShapeTable in;
  for (size_t i = 0; i < a.ids.size(); ++i) in[a.ids[i]] = a.shapes[i];
  ShapeTable shapes = infer_shapes(prog, *t, in);
  for (int id : prog.params) shapes[id] = prog.tdef(id).shape;
  for (int id : prog.caches) shapes[id] = prog.tdef(id).shape;
  std::map<int, void*> ptrs;
  for (auto& kv : shapes) ptrs[kv.first] = reinterpret_cast<void*>((uintptr_t)kv.first << 32);
  auto num = [](int64_t v) { return std::to_string((long long)v); };
  auto op_text = [&](const IpTensorOp& op) {
    std::string s = "{\"tensor\":" + num((int64_t)(op.base >> 32)) + ",\"offset\":" + num(op.offset) + ",\"dst\":" + num(op.dst) +
                    ",\"terms\":[";
    for (int q = 0; q < op.nterms; ++q) s += std::string(q ? "," : "") + "[" + num(op.slot[q]) + "," + num(op.coef[q]) + "]";
    return s + "]}";
  };
  auto instrs_text = [&](const IpInstr* ins, int n) {
    std::string s = "[";
    for (int q = 0; q < n; ++q)
      s += std::string(q ? "," : "") + "[" + num(ins[q].op) + "," + num(ins[q].dst) + "," + num(ins[q].a) + "," + num(ins[q].b) +
           "," + num(ins[q].c) + "," + num(ins[q].imm) + "]";
    return s + "]";
  };
  std::string s;
  for (size_t ki = 0; ki < t->kernels.size(); ++ki) {
    const Kernel& k = *t->kernels[ki];
    if (k.is_generator()) fail(EGB_ERR_GENERATOR, "program still contains generator kernels; compile it first");
    // every kernel accumulates into its (zero-initialised) result, like the reference before InstrOverwrite
    Lowered lw = lower_kernel(k, shapes, ptrs, epoch, strict != 0, false, 148);
    const IpProgram& ip = lw.ip;
    s += "{\"kernel\":" + num((int64_t)ki) + ",\"npar\":" + num(ip.npar) + ",\"npoints\":" + num(ip.npoints) + ",\"nred\":" + num(ip.nred) +
         ",\"accumulate\":" + num(ip.accumulate) + ",\"scatter\":" + num(ip.scatter) + ",\"nslots\":" + num(ip.nslots) + ",\"loops\":[";
    for (int q = 0; q < ip.nloops; ++q)
      s += std::string(q ? "," : "") + "[" + num(ip.loops[q].start) + "," + num(ip.loops[q].step) + "," + num(ip.loops[q].count) + "," +
           num(ip.loops[q].slot) + "]";
    s += "],\"lits\":[";
    for (int q = 0; q < ip.nlits; ++q) s += std::string(q ? "," : "") + "[" + num(ip.lit_slot[q]) + ",\"" + std::to_string((unsigned long long)ip.lits[q]) + "\"]";
    s += "],\"array_table\":[";
    for (int q = 0; q < 64; ++q) s += std::string(q ? "," : "") + num(ip.array_table[q]);
    s += "],\"index_instrs\":" + instrs_text(ip.index_instrs, ip.nindex_instrs) + ",\"instrs\":" + instrs_text(ip.instrs, ip.ninstrs) +
         ",\"reads\":[";
    for (int q = 0; q < ip.nreads; ++q) s += std::string(q ? "," : "") + op_text(ip.reads[q]);
    s += "],\"write\":" + op_text(ip.write);
    // what the planner's matchers make of this kernel (the same order as the planner: contraction, conv2, map form):
    // the operands and constants a specialised device kernel would be launched with
    GemmPattern g;
    ConvPattern cv;
    EltSpec es;
    if (match_gemm(k, shapes, g)) {
      s += ",\"gemm\":{\"a\":" + num(g.a_tensor) + ",\"b\":" + num(g.b_tensor) + ",\"c\":" + num(g.c_tensor) + ",\"ta\":" + num(g.trans_a) +
           ",\"tb\":" + num(g.trans_b) + ",\"M\":" + num(g.M) + ",\"N\":" + num(g.N) + ",\"K\":" + num(g.K) + ",\"lda\":" + num(g.lda) +
           ",\"ldb\":" + num(g.ldb) + ",\"ldc\":" + num(g.ldc) + "}";
    } else if (match_conv2(k, shapes, cv)) {
      s += ",\"conv2\":{\"kind\":" + num((int)cv.kind) + ",\"img\":" + num(cv.img_tensor) + ",\"fil\":" + num(cv.fil_tensor) + ",\"out\":" +
           num(cv.out_tensor) + ",\"N\":" + num(cv.N) + ",\"H\":" + num(cv.H) + ",\"W\":" + num(cv.W) + ",\"C\":" + num(cv.C) + ",\"F\":" +
           num(cv.F) + ",\"KH\":" + num(cv.KH) + ",\"KW\":" + num(cv.KW) + "}";
    } else if (match_eltwise(k, shapes, es)) {
      char lit[160];
      snprintf(lit, sizeof(lit), "[%.17g,%.17g,%.17g,%.17g]", es.lit[0], es.lit[1], es.lit[2], es.lit[3]);
      s += std::string(",\"eltwise\":{\"form\":\"") + elt_kind_name(es.kind) + "\",\"n\":" + num(es.n) + ",\"row\":" + num(es.row) +
           ",\"uses_epoch\":" + num(es.uses_epoch) + ",\"lit\":" + lit + ",\"reads\":[";
      for (int q = 0; q < es.nreads; ++q)
        s += std::string(q ? "," : "") + "{\"tensor\":" + num(es.read_tensor[q]) + ",\"row\":" + num(es.row_read[q]) + ",\"scalar\":" +
             num(es.scalar_read[q]) + ",\"offset\":" + num(es.scalar_offset[q]) + "}";
      s += "]}";
    }
    s += "}\n";
  }
  copy_out(s, buf, cap, needed);
  EGB_CATCH
}

int egb_program_free(egb_program* p) {
  EGB_TRY
  delete p;
  EGB_CATCH
}

int egb_program_tensor_count(egb_program* p, int* count) {
  EGB_TRY
  *count = (int)p->p->tensors.size();
  EGB_CATCH
}

int egb_program_tensor_info(egb_program* p, int tensor_id, int* kind, int* rank, int64_t* dims, char* name,
                            size_t name_cap) {
  EGB_TRY
  if (tensor_id < 1 || tensor_id > (int)p->p->tensors.size()) fail(EGB_ERR_RUNTIME, "no tensor with id %d", tensor_id);
  const TensorDef& t = p->p->tdef(tensor_id);
  if (kind) *kind = (int)t.kind;
  if (rank) *rank = (int)t.shape.size();
  if (dims)
    for (size_t i = 0; i < t.shape.size() && i < EGB_MAX_RANK; ++i) dims[i] = t.shape[i];
  copy_out(t.name, name, name_cap, nullptr);
  EGB_CATCH
}

int egb_program_target_output(egb_program* p, const char* target, int* tensor_id) {
  EGB_TRY
  Target* t = p->p->find_target(target);
  if (!t) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  *tensor_id = t->output;
  EGB_CATCH
}

int egb_program_infer_shapes(egb_program* p, const char* target, int n_args, const char* const* names,
                             const int* ranks, const int64_t* dims, int tensor_id, int* out_rank,
                             int64_t* out_dims) {
  EGB_TRY
  Program& prog = *p->p;
  if (!prog.compiled) compile_program(prog);
  Target* t = prog.find_target(target);
  if (!t) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  Args a = resolve_args(prog, n_args, names, ranks, dims);
  ShapeTable in;
  for (size_t i = 0; i < a.ids.size(); ++i) in[a.ids[i]] = a.shapes[i];
  ShapeTable res = infer_shapes(prog, *t, in);
  const int id = tensor_id ? tensor_id : t->output;
  if (!id) {  // the target has no output tensor
    *out_rank = -1;
    return EGB_OK;
  }
  auto it = res.find(id);
  if (it == res.end()) fail(EGB_ERR_SHAPE, "Missing shape for tensor%d", id - 1);
  if (it->second.size() > EGB_MAX_RANK) fail(EGB_ERR_SHAPE, "rank %zu exceeds EGB_MAX_RANK", it->second.size());
  *out_rank = (int)it->second.size();
  for (size_t i = 0; i < it->second.size(); ++i) out_dims[i] = it->second[i];
  EGB_CATCH
}

// ---- compile / arg / run: the per-kernel launch interface of exprgrad/runtimes/gpu.nim:46-50 --------
// (cl.nim:149-207). `source` is not OpenCL C but the text of a compiled program with one target
// (the kernels the JIT'd host code would launch one by one); arguments are the target's tensors in
// `target.tensors` order. The reference passes shapes to its OpenCL kernels as captured Index registers
// (llvmgen.nim:474-486); here they are attached per tensor with egb_kernel_arg_shape.
struct egb_kernel {
  egb_context* ctx = nullptr;
  std::shared_ptr<Program> prog;
  Target* target = nullptr;
  std::map<int, void*> ptrs;   // tensor id -> device pointer
  ShapeTable shapes;
  int64_t epoch = 0;
};

int egb_compile(egb_context* ctx, const char* name, const char* source, egb_kernel** out) {
  EGB_TRY
  auto k = std::make_unique<egb_kernel>();
  k->ctx = ctx;
  try {
    k->prog = parse_program(source);
  } catch (const Error& e) {
    fail(EGB_ERR_GPU, "Failed to build program: %s", e.what());  // cl.nim:163-171
  }
  if (!k->prog->compiled) compile_program(*k->prog);
  if (k->prog->f64) fail(EGB_ERR_GPU, "Failed to build program: float64 kernels are not supported");
  k->target = k->prog->find_target(name);
  if (!k->target) fail(EGB_ERR_GPU, "Failed to build program: no target named %s", name);
  *out = k.release();
  EGB_CATCH
}

int egb_kernel_free(egb_kernel* k) {
  EGB_TRY
  delete k;
  EGB_CATCH
}

int egb_kernel_arg_count(egb_kernel* k, int* count) {
  EGB_TRY
  *count = (int)k->target->tensors.size();
  EGB_CATCH
}

static int kernel_tensor(egb_kernel* k, int index) {
  if (index < 0 || index >= (int)k->target->tensors.size())
    fail(EGB_ERR_GPU, "kernel argument %d out of range (the kernel has %zu tensor arguments)", index,
         k->target->tensors.size());
  return k->target->tensors[index];
}

int egb_kernel_arg_buffer(egb_kernel* k, int index, egb_buffer* buf) {
  EGB_TRY
  k->ptrs[kernel_tensor(k, index)] = buf->ptr;
  EGB_CATCH
}

int egb_kernel_arg_shape(egb_kernel* k, int index, int rank, const int64_t* dims) {
  EGB_TRY
  k->shapes[kernel_tensor(k, index)] = std::vector<int64_t>(dims, dims + rank);
  EGB_CATCH
}

int egb_kernel_arg_index(egb_kernel* k, int index, int64_t value) {
  EGB_TRY
  if (index != -1) fail(EGB_ERR_GPU, "index arguments other than the epoch (-1) are derived from the tensor shapes");
  k->epoch = value;
  EGB_CATCH
}

int egb_kernel_run(egb_kernel* k, int work_dims, const int64_t* group_size, const int64_t* local_size) {
  EGB_TRY
  (void)group_size;
  (void)local_size;
  if (work_dims <= 0) fail(EGB_ERR_GPU, "Group size must have at least one dimension");  // cl.nim:191-192
  Context& c = k->ctx->c;
  EGB_CUDA(cudaSetDevice(c.device));
  for (int id : k->target->tensors) {
    if (!k->ptrs.count(id)) fail(EGB_ERR_GPU, "kernel argument for tensor%d has not been set", id - 1);
    if (!k->shapes.count(id)) fail(EGB_ERR_GPU, "shape of tensor%d has not been set", id - 1);
  }
  // launch geometry is the library's; every kernel accumulates into its destination like the
  // reference's generated `+=` (clgen.nim:116-125) - the caller zero-fills results (model.nim:302-318)
  for (auto& kp : k->target->kernels) {
    const Kernel& kern = *kp;
    GemmPattern g;
    ConvPattern cv;
    if (match_gemm(kern, k->shapes, g)) {
      int st = egb_gemm_f32(k->ctx, g.trans_a, g.trans_b, g.M, g.N, g.K, (const float*)k->ptrs[g.a_tensor], g.lda,
                            (const float*)k->ptrs[g.b_tensor], g.ldb, (float*)k->ptrs[g.c_tensor], g.ldc, 1, nullptr, 1.0f);
      if (st != EGB_OK) return st;
    } else if (match_conv2(kern, k->shapes, cv) && (cv.kind != ConvPattern::D_IMAGES || conv2_dimg_supported(cv.KW))) {
      const float* img = (const float*)k->ptrs[cv.img_tensor];
      const float* fil = (const float*)k->ptrs[cv.fil_tensor];
      const float* out_t = (const float*)k->ptrs[cv.out_tensor];
      if (cv.kind == ConvPattern::FORWARD)
        launch_conv2_fwd(c, img, fil, (float*)out_t, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, true, c.stream);
      else if (cv.kind == ConvPattern::D_FILTERS)
        launch_conv2_dw(c, img, out_t, (float*)fil, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, c.stream);
      else
        launch_conv2_dimg(c, out_t, fil, (float*)img, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, true, c.stream);
    } else {
      Lowered lw = lower_kernel(kern, k->shapes, k->ptrs, k->epoch, false, false, c.sm_count);
      launch_interp(c, lw.ip, lw.pb, lw.rb, lw.points_fast, false, c.stream);
    }
  }
  EGB_CATCH
}

int egb_model_create(egb_context* ctx, egb_program* program, uint64_t seed, egb_model** out) {
  EGB_TRY
  EGB_CUDA(cudaSetDevice(ctx->c.device));
  auto m = new egb_model();
  m->ctx = ctx;
  try {
    m->m = new_model(ctx->c, program->p, seed);
  } catch (...) {
    delete m;
    throw;
  }
  *out = m;
  EGB_CATCH
}

int egb_model_free(egb_model* m) {
  EGB_TRY
  delete m;
  EGB_CATCH
}

int egb_model_set_option(egb_model* m, const char* key, int64_t value) {
  EGB_TRY
  std::string k(key);
  if (k == "strict") {
    if (m->m->strict != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->strict = value != 0;
  } else if (k == "graphs") {
    m->m->use_graphs = value != 0;
  } else if (k == "concurrent") {
    if (m->m->concurrent != (value != 0)) {
      for (auto& p : m->m->plans) p->graph_valid = false;
    }
    m->m->concurrent = value != 0;
  } else if (k == "splitk") {
    if (m->m->splitk != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->splitk = value != 0;
  } else if (k == "rowchain") {
    if (m->m->rowchain != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->rowchain = value != 0;
  } else if (k == "keep_intermediates") {
    if (m->m->keep_intermediates != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->keep_intermediates = value != 0;
  } else if (k == "dp_peer") {
    if (m->m->dp_peer != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->dp_peer = value != 0;
  } else if (k == "head") {
    if (m->m->headfuse != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->headfuse = value != 0;
  } else if (k == "eltwise") {
    if (m->m->eltwise != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->eltwise = value != 0;
  } else if (k == "fuse") {
    if (m->m->fuse != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->fuse = value != 0;
  } else if (k == "plan_cache") {
    if (value < 1) fail(EGB_ERR_RUNTIME, "plan_cache must be at least 1");
    m->m->max_plans_per_target = (int)value;
  } else if (k == "epoch") {
    m->m->epoch = value;
  } else {
    fail(EGB_ERR_RUNTIME, "unknown model option '%s'", key);
  }
  EGB_CATCH
}

int egb_model_epoch(egb_model* m, int64_t* epoch) {
  EGB_TRY
  *epoch = m->m->epoch;
  EGB_CATCH
}

static DevTensor* find_tensor(egb_model* m, int tensor_id) {
  auto s = m->m->state.find(tensor_id);
  if (s != m->m->state.end()) return &s->second;
  if (m->m->last_plan) {
    if (m->m->last_plan->unmaterialized.count(tensor_id))
      fail(EGB_ERR_RUNTIME, "tensor%d is consumed inside a fused contraction epilogue and never stored by this plan; set the "
                            "model option keep_intermediates=1 to materialise every intermediate", tensor_id - 1);
    auto t = m->m->last_plan->tensors.find(tensor_id);
    if (t != m->m->last_plan->tensors.end()) return &t->second;
  }
  fail(EGB_ERR_RUNTIME, "tensor%d has no device storage (not a parameter/cache and not part of the last call)",
       tensor_id - 1);
}

int egb_model_write_tensor(egb_model* m, int tensor_id, const void* host, size_t bytes) {
  EGB_TRY
  DevTensor* t = find_tensor(m, tensor_id);
  if (bytes != t->bytes)
    fail(EGB_ERR_GPU, "Attempted to write %zu bytes, but the size of the buffer is %zu bytes", bytes, t->bytes);
  cudaStream_t st = m->ctx->c.stream;
  if (bytes) {
    EGB_CUDA(cudaMemcpyAsync(t->ptr, host, bytes, cudaMemcpyHostToDevice, st));
    EGB_CUDA(cudaStreamSynchronize(st));
  }
  EGB_CATCH
}

int egb_model_read_tensor(egb_model* m, int tensor_id, void* host, size_t bytes) {
  EGB_TRY
  DevTensor* t = find_tensor(m, tensor_id);
  if (bytes != t->bytes) fail(EGB_ERR_GPU, "Buffer size is not equal to target size");
  cudaStream_t st = m->ctx->c.stream;
  if (bytes) {
    EGB_CUDA(cudaMemcpyAsync(host, t->ptr, bytes, cudaMemcpyDeviceToHost, st));
    EGB_CUDA(cudaStreamSynchronize(st));
  }
  EGB_CATCH
}

int egb_model_tensor_shape(egb_model* m, int tensor_id, int* rank, int64_t* dims) {
  EGB_TRY
  DevTensor* t = find_tensor(m, tensor_id);
  *rank = (int)t->shape.size();
  for (size_t i = 0; i < t->shape.size() && i < EGB_MAX_RANK; ++i) dims[i] = t->shape[i];
  EGB_CATCH
}

int egb_model_tensor_device_ptr(egb_model* m, int tensor_id, void** ptr) {
  EGB_TRY
  *ptr = find_tensor(m, tensor_id)->ptr;
  EGB_CATCH
}

int egb_model_call(egb_model* m, const char* target, int n_args, const char* const* names, const void* const* data,
                   const int* ranks, const int64_t* dims, const int* on_device, int* out_rank, int64_t* out_dims) {
  EGB_TRY
  Model& model = *m->m;
  Context& c = m->ctx->c;
  EGB_CUDA(cudaSetDevice(c.device));
  if (!model.prog->find_target(target)) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  Args a = resolve_args(*model.prog, n_args, names, ranks, dims);
  Plan& plan = model.get_plan(target, a.ids, a.shapes);
  bool rebind = false;
  for (int i = 0; i < n_args; ++i) {
    const int id = a.ids[i];
    auto t = plan.tensors.find(id);
    if (t == plan.tensors.end()) continue;  // the target does not use this input
    const bool dev = on_device && on_device[i];
    const void* want = dev ? data[i] : nullptr;
    auto b = plan.bound.find(id);
    const void* have = b == plan.bound.end() ? nullptr : b->second;
    if (want != have) {
      plan.bound[id] = want;
      rebind = true;
    }
    if (!dev && t->second.bytes)
      EGB_CUDA(cudaMemcpyAsync(t->second.ptr, data[i], t->second.bytes, cudaMemcpyHostToDevice, c.stream));
  }
  if (rebind) model.build_nodes(plan);
  model.run(plan);
  const int out = plan.target->output;
  if (out_rank) {
    *out_rank = 0;
    if (out) {
      const auto& shape = plan.shapes.at(out);
      if (shape.size() > EGB_MAX_RANK) fail(EGB_ERR_SHAPE, "rank %zu exceeds EGB_MAX_RANK", shape.size());
      *out_rank = (int)shape.size();
      for (size_t i = 0; i < shape.size(); ++i) out_dims[i] = shape[i];
    } else {
      *out_rank = -1;
    }
  }
  EGB_CATCH
}

// ---- Model.call with a host destination (model.nim:392-406: run the target, then readOutput) -------
// A target that is one plain contraction of host operands (benchmarks/matmul/matmul_gpu.nim:28-75 times
// exactly this: upload, multiply, read back) is limited by the PCIe copies, not by the tensor cores. Its
// rows are independent (`y` is a parallel loop of c[y,x] ++= a[y,it]*b[it,x]), so the call is streamed:
// B is uploaded first, then A in row blocks; each block is split into bf16 planes, multiplied and copied
// back while the next block is still arriving - H2D, tensor cores and D2H overlap on three streams.
namespace {

struct SplitOp {   // one operand split of the plan (a SPLIT node or one job of a merged one)
  const float* split_src = nullptr;
  __nv_bfloat16 *split_hi = nullptr, *split_mid = nullptr;
  int split_rows = 0, split_cols = 0, split_ld = 0, split_dst_ld = 0, split_act = 0;
  bool split_transpose = false;
};

struct RowStream {
  SplitOp sa, sb;
  SplitOp* split_a = nullptr;
  SplitOp* split_b = nullptr;
  Node* gemm = nullptr;
  Node* memset = nullptr;
  int a_id = 0, b_id = 0;
};

bool find_row_stream(Plan& plan, const Args& a, const int* on_device, RowStream& rs) {
  std::vector<SplitOp> splits;
  for (auto& n : plan.nodes) {
    if (n.kind == Node::SPLIT && n.bytes) return false;   // the split launch also clears results: not streamable
    if (n.kind == Node::SPLIT && !n.split_jobs.empty()) {
      for (auto& j : n.split_jobs) splits.push_back(SplitOp{j.src, j.hi, j.mid, j.rows, j.cols, j.ld, j.dst_ld, j.act, false});
    } else if (n.kind == Node::SPLIT) {
      splits.push_back(SplitOp{n.split_src, n.split_hi, n.split_mid, n.split_rows, n.split_cols, n.split_ld, n.split_dst_ld,
                               n.split_act, n.split_transpose});
    } else if (n.kind == Node::GEMM && !rs.gemm) rs.gemm = &n;
    else if (n.kind == Node::MEMSET && !rs.memset) rs.memset = &n;
    else return false;
  }
  if (!rs.gemm || splits.size() != 2) return false;
  const GemmArgs& g = rs.gemm->gemm;
  const int out = plan.target->output;
  if (!out || g.flags != 0 || g.epi != EPI_NONE || g.bias || g.colsum || g.a_mn || g.alpha != 1.0f) return false;
  auto ot = plan.tensors.find(out);
  if (ot == plan.tensors.end() || g.C != ot->second.ptr || g.ldc != g.N) return false;
  for (auto& s : splits) {
    if (s.split_hi == g.a_hi) { rs.sa = s; rs.split_a = &rs.sa; }
    else if (s.split_hi == g.b_hi) { rs.sb = s; rs.split_b = &rs.sb; }
  }
  if (!rs.split_a || !rs.split_b || rs.split_a->split_transpose || rs.split_a->split_act || rs.split_a->split_rows != g.M) return false;
  // both operands must be inputs of this call; A has to come from the host
  for (size_t i = 0; i < a.ids.size(); ++i) {
    auto t = plan.tensors.find(a.ids[i]);
    if (t == plan.tensors.end()) continue;
    const bool dev = on_device && on_device[i];
    if (!dev && t->second.ptr == rs.split_a->split_src) rs.a_id = a.ids[i];
    if (t->second.ptr == rs.split_b->split_src || (dev && plan.bound[a.ids[i]] == rs.split_b->split_src)) rs.b_id = a.ids[i];
  }
  return rs.a_id != 0 && rs.b_id != 0 && (int64_t)g.M * g.K >= ((int64_t)1 << 20);
}

struct StreamEvents {
  std::vector<cudaEvent_t> ev;
  cudaEvent_t get(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e;
      EGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    return ev[i];
  }
};

void run_row_stream(Model& model, Context& c, Plan& plan, const RowStream& rs, const Args& a, const void* const* data,
                    const int* on_device, char* out_host) {
  // events belong to the device they were created on: one pool per device (a process may open several contexts)
  static thread_local std::map<int, StreamEvents> events_by_device;
  StreamEvents& events = events_by_device[c.device];
  for (auto& st : c.aux_stream)
    if (!st) EGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaStream_t h2d = c.aux_stream[0], d2h = c.aux_stream[1], comp = c.stream;
  const GemmArgs& g = rs.gemm->gemm;
  const SplitOp& sa = *rs.split_a;
  // everything queued on the main stream so far (earlier calls) is ordered before the copies
  cudaEvent_t e_start = events.get(0);
  EGB_CUDA(cudaEventRecord(e_start, comp));
  EGB_CUDA(cudaStreamWaitEvent(h2d, e_start, 0));
  EGB_CUDA(cudaStreamWaitEvent(d2h, e_start, 0));
  const char* a_host = nullptr;
  for (size_t i = 0; i < a.ids.size(); ++i) {
    auto t = plan.tensors.find(a.ids[i]);
    if (t == plan.tensors.end()) continue;
    if (a.ids[i] == rs.a_id) {
      a_host = (const char*)data[i];
      continue;
    }
    if (!(on_device && on_device[i]) && t->second.bytes)
      EGB_CUDA(cudaMemcpyAsync(t->second.ptr, data[i], t->second.bytes, cudaMemcpyHostToDevice, h2d));
  }
  cudaEvent_t e_b = events.get(1);
  EGB_CUDA(cudaEventRecord(e_b, h2d));
  EGB_CUDA(cudaStreamWaitEvent(comp, e_b, 0));
  if (rs.memset) EGB_CUDA(cudaMemsetAsync(rs.memset->ptr, 0, rs.memset->bytes, comp));
  {
    const SplitOp& sb = *rs.split_b;
    launch_split_bf16(c, sb.split_src, sb.split_rows, sb.split_cols, sb.split_ld, sb.split_transpose, sb.split_hi,
                      sb.split_mid, sb.split_dst_ld, sb.split_act, comp);
  }
  // row blocks of about 2 MiB of A (multiples of the 128-row MMA tile)
  const size_t row_bytes = (size_t)sa.split_cols * 4;
  static const char* blk_env = getenv("EGB_STREAM_BLOCK_KIB");
  const size_t blk_bytes = blk_env ? (size_t)atol(blk_env) << 10 : (size_t)2 << 20;
  int64_t rows_per = (int64_t)(blk_bytes / std::max<size_t>(row_bytes, 1)) / 128 * 128;
  if (rows_per < 128) rows_per = 128;
  size_t ei = 2;
  for (int64_t r0 = 0; r0 < g.M; r0 += rows_per) {
    const int rc = (int)std::min<int64_t>(rows_per, g.M - r0);
    char* a_dev = (char*)const_cast<float*>(sa.split_src) + (size_t)r0 * sa.split_ld * 4;
    if ((size_t)sa.split_ld * 4 == row_bytes)
      EGB_CUDA(cudaMemcpyAsync(a_dev, a_host + (size_t)r0 * row_bytes, (size_t)rc * row_bytes, cudaMemcpyHostToDevice, h2d));
    else
      EGB_CUDA(cudaMemcpy2DAsync(a_dev, (size_t)sa.split_ld * 4, a_host + (size_t)r0 * row_bytes, row_bytes, row_bytes, rc,
                                 cudaMemcpyHostToDevice, h2d));
    cudaEvent_t e_a = events.get(ei++);
    EGB_CUDA(cudaEventRecord(e_a, h2d));
    EGB_CUDA(cudaStreamWaitEvent(comp, e_a, 0));
    launch_split_bf16(c, sa.split_src + (size_t)r0 * sa.split_ld, rc, sa.split_cols, sa.split_ld, false,
                      sa.split_hi + (size_t)r0 * sa.split_dst_ld, sa.split_mid + (size_t)r0 * sa.split_dst_ld, sa.split_dst_ld, 0, comp);
    GemmArgs gc = g;
    gc.a_hi = g.a_hi + (size_t)r0 * g.lda;
    gc.a_mid = g.a_mid + (size_t)r0 * g.lda;
    gc.M = rc;
    gc.C = g.C + (size_t)r0 * g.ldc;
    gc.bn = 0;
    launch_gemm_bf16x3(c, gc, comp);
    cudaEvent_t e_c = events.get(ei++);
    EGB_CUDA(cudaEventRecord(e_c, comp));
    EGB_CUDA(cudaStreamWaitEvent(d2h, e_c, 0));
    EGB_CUDA(cudaMemcpyAsync(out_host + (size_t)r0 * g.ldc * 4, gc.C, (size_t)rc * g.ldc * 4, cudaMemcpyDeviceToHost, d2h));
  }
  cudaEvent_t e_done = events.get(ei++);
  EGB_CUDA(cudaEventRecord(e_done, d2h));
  EGB_CUDA(cudaStreamWaitEvent(comp, e_done, 0));
  EGB_CUDA(cudaStreamSynchronize(comp));
  plan.runs++;
  model.last_plan = &plan;
}

}  // namespace

int egb_model_call_read(egb_model* m, const char* target, int n_args, const char* const* names, const void* const* data,
                        const int* ranks, const int64_t* dims, const int* on_device, void* out_host, size_t out_bytes,
                        int* out_rank, int64_t* out_dims) {
  EGB_TRY
  Model& model = *m->m;
  Context& c = m->ctx->c;
  EGB_CUDA(cudaSetDevice(c.device));
  if (!model.prog->find_target(target)) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  Args a = resolve_args(*model.prog, n_args, names, ranks, dims);
  Plan& plan = model.get_plan(target, a.ids, a.shapes);
  const int out = plan.target->output;
  if (!out) fail(EGB_ERR_RUNTIME, "target %s has no output tensor", target);
  const auto& shape = plan.shapes.at(out);
  if (shape.size() > EGB_MAX_RANK) fail(EGB_ERR_SHAPE, "rank %zu exceeds EGB_MAX_RANK", shape.size());
  if (out_rank) {
    *out_rank = (int)shape.size();
    for (size_t i = 0; i < shape.size(); ++i) out_dims[i] = shape[i];
  }
  if ((size_t)shape_len(shape) * 4 != out_bytes) fail(EGB_ERR_GPU, "Buffer size is not equal to target size");  // cl.nim:134-135
  bool rebind = false;
  for (int i = 0; i < n_args; ++i) {
    const int id = a.ids[i];
    if (plan.tensors.find(id) == plan.tensors.end()) continue;
    const void* want = (on_device && on_device[i]) ? data[i] : nullptr;
    auto b = plan.bound.find(id);
    const void* have = b == plan.bound.end() ? nullptr : b->second;
    if (want != have) {
      plan.bound[id] = want;
      rebind = true;
    }
  }
  if (rebind) model.build_nodes(plan);
  RowStream rs;
  if (!model.strict && !c.timing && find_row_stream(plan, a, on_device, rs)) {
    run_row_stream(model, c, plan, rs, a, data, on_device, (char*)out_host);
  } else {
    for (int i = 0; i < n_args; ++i) {
      auto t = plan.tensors.find(a.ids[i]);
      if (t == plan.tensors.end()) continue;
      if (!(on_device && on_device[i]) && t->second.bytes)
        EGB_CUDA(cudaMemcpyAsync(t->second.ptr, data[i], t->second.bytes, cudaMemcpyHostToDevice, c.stream));
    }
    model.run(plan);
    DevTensor* t = find_tensor(m, out);
    if (out_bytes) EGB_CUDA(cudaMemcpyAsync(out_host, t->ptr, out_bytes, cudaMemcpyDeviceToHost, c.stream));
    EGB_CUDA(cudaStreamSynchronize(c.stream));
  }
  EGB_CATCH
}

int egb_model_read_output(egb_model* m, void* dst, size_t bytes) {
  EGB_TRY
  Model& model = *m->m;
  if (!model.last_plan || !model.last_plan->target->output) fail(EGB_ERR_RUNTIME, "the last call has no output tensor");
  const int out = model.last_plan->target->output;
  DevTensor* t = find_tensor(m, out);
  if (bytes != t->bytes) fail(EGB_ERR_GPU, "Buffer size is not equal to target size");
  cudaStream_t st = m->ctx->c.stream;
  if (bytes) EGB_CUDA(cudaMemcpyAsync(dst, t->ptr, bytes, cudaMemcpyDeviceToHost, st));
  EGB_CUDA(cudaStreamSynchronize(st));
  EGB_CATCH
}

int egb_model_fit(egb_model* m, const char* target, int n_args, const char* const* names, const void* const* data,
                  const int* ranks, const int64_t* dims, int64_t batch_size, int64_t* batches_run) {
  EGB_TRY
  Model& model = *m->m;
  Context& c = m->ctx->c;
  EGB_CUDA(cudaSetDevice(c.device));
  if (n_args == 0)
    fail(EGB_ERR_RUNTIME,
         "Model.fit requires at least one input tensor. Use Model.apply instead if the target has zero inputs.");
  if (!model.prog->find_target(target)) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target);
  if (batch_size <= 0) fail(EGB_ERR_RUNTIME, "batch size must be positive");
  Args a = resolve_args(*model.prog, n_args, names, ranks, dims);
  std::vector<int64_t> totals, row_bytes;
  for (auto& s : a.shapes) {
    if (s.empty()) fail(EGB_ERR_SHAPE, "Model.fit inputs need at least one dimension");
    totals.push_back(s[0]);
    s[0] = batch_size;
    row_bytes.push_back(shape_len(s) / batch_size * 4);
  }
  const int64_t batch_count = totals[0] / batch_size;
  // every argument is sliced with viewFirst(b * batchSize, batchSize) (model.nim:443, tensors.nim:290-297): an
  // argument with fewer rows than the batches cover would be read past its end
  for (int i = 0; i < n_args; ++i)
    if (totals[i] < batch_count * batch_size)
      fail(EGB_ERR_SHAPE, "Model.fit: input %s has %lld rows, but %lld batches of %lld need %lld", names[i],
           (long long)totals[i], (long long)batch_count, (long long)batch_size, (long long)(batch_count * batch_size));
  Plan& plan = model.get_plan(target, a.ids, a.shapes);
  bool rebind = false;
  for (auto& kv : plan.bound)
    if (kv.second) {
      kv.second = nullptr;
      rebind = true;
    }
  if (rebind) model.build_nodes(plan);
  model.epoch += 1;
  // The data set is uploaded ONCE per fit (in chunks of whole batches when it does not fit the staging budget,
  // double buffered: the upload of chunk c+1 overlaps the training steps of chunk c) and every batch is a
  // device-side slice of it - the device equivalent of the reference's zero-copy viewFirst. A batch reaches the
  // plan's input tensors through a device-to-device copy on the compute stream (~1 us for a 3 MB batch), so the
  // captured CUDA graph and its pointers stay valid; nothing crosses PCIe inside the batch loop.
  std::vector<size_t> batch_bytes(n_args, 0), arg_off(n_args, 0);
  size_t bytes_per_batch = 0;
  for (int i = 0; i < n_args; ++i) {
    auto t = plan.tensors.find(a.ids[i]);
    if (t == plan.tensors.end() || !t->second.bytes) continue;
    batch_bytes[i] = (size_t)batch_size * row_bytes[i];
    arg_off[i] = bytes_per_batch;
    bytes_per_batch += (batch_bytes[i] + 255) / 256 * 256;
  }
  if (batch_count > 0 && bytes_per_batch > 0) {
    size_t free_b = 0, total_b = 0;
    EGB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const char* budget_env = getenv("EGB_FIT_STAGE_MIB");  // (tests: forces chunked staging)
    size_t budget = budget_env ? (size_t)atol(budget_env) << 20 : std::min<size_t>(free_b / 4, (size_t)4 << 30);
    int64_t per_chunk = std::max<int64_t>(1, (int64_t)(budget / 2 / bytes_per_batch));
    const bool single = per_chunk >= batch_count;
    if (single) per_chunk = batch_count;
    const int nbuf = single ? 1 : 2;
    const size_t chunk_bytes = (size_t)per_chunk * bytes_per_batch;
    char* stage[2] = {nullptr, nullptr};
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    for (auto& st : c.aux_stream)
      if (!st) EGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaStream_t h2d = c.aux_stream[0], comp = c.stream;
    auto cleanup = [&]() {
      cudaStreamSynchronize(h2d);
      cudaStreamSynchronize(comp);
      for (int q = 0; q < 2; ++q) {
        if (stage[q]) cudaFree(stage[q]);
        if (up[q]) cudaEventDestroy(up[q]);
        if (done[q]) cudaEventDestroy(done[q]);
      }
    };
    try {
      for (int q = 0; q < nbuf; ++q) {
        EGB_CUDA(cudaMalloc((void**)&stage[q], chunk_bytes));
        EGB_CUDA(cudaEventCreateWithFlags(&up[q], cudaEventDisableTiming));
        EGB_CUDA(cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming));
      }
      // earlier work on the compute stream (parameter writes ...) is ordered before the first upload
      EGB_CUDA(cudaEventRecord(done[0], comp));
      EGB_CUDA(cudaStreamWaitEvent(h2d, done[0], 0));
      int64_t chunk_index = 0;
      for (int64_t b0 = 0; b0 < batch_count; b0 += per_chunk, ++chunk_index) {
        const int q = (int)(chunk_index % nbuf);
        const int64_t nb = std::min<int64_t>(per_chunk, batch_count - b0);
        if (chunk_index >= nbuf) EGB_CUDA(cudaStreamWaitEvent(h2d, done[q], 0));  // steps of chunk c-2 are done with it
        for (int i = 0; i < n_args; ++i) {
          if (!batch_bytes[i]) continue;
          // layout inside a chunk: [arg][batch] - one contiguous upload per argument
          const char* src = (const char*)data[i] + (size_t)b0 * batch_bytes[i];
          EGB_CUDA(cudaMemcpyAsync(stage[q] + arg_off[i] * (size_t)per_chunk, src, (size_t)nb * batch_bytes[i],
                                   cudaMemcpyHostToDevice, h2d));
        }
        EGB_CUDA(cudaEventRecord(up[q], h2d));
        EGB_CUDA(cudaStreamWaitEvent(comp, up[q], 0));
        for (int64_t b = 0; b < nb; ++b) {
          for (int i = 0; i < n_args; ++i) {
            if (!batch_bytes[i]) continue;
            auto t = plan.tensors.find(a.ids[i]);
            EGB_CUDA(cudaMemcpyAsync(t->second.ptr, stage[q] + arg_off[i] * (size_t)per_chunk + (size_t)b * batch_bytes[i],
                                     batch_bytes[i], cudaMemcpyDeviceToDevice, comp));
          }
          model.run(plan);
        }
        EGB_CUDA(cudaEventRecord(done[q], comp));
      }
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  } else {
    for (int64_t b = 0; b < batch_count; ++b) model.run(plan);
  }
  if (batches_run) *batches_run = batch_count;
  EGB_CATCH
}

int egb_model_plan_count(egb_model* m, int* count) {
  EGB_TRY
  *count = (int)m->m->plans.size();
  EGB_CATCH
}

int egb_model_describe_plan(egb_model* m, char* buf, size_t cap, size_t* needed) {
  EGB_TRY
  Model& model = *m->m;
  std::string s;
  if (model.last_plan) {
    Plan& p = *model.last_plan;
    s += "target " + p.target_name + ": " + std::to_string(p.nodes.size()) + " nodes, arena " +
         std::to_string(p.arena_bytes) + " bytes (zeroed per run: " + std::to_string(p.zero_bytes) + "), graph " +
         (p.graph_valid ? "yes" : "no") + "\n";
    static const char* kinds[] = {"interp", "gemm", "split", "memset", "random", "allreduce", "conv", "rowchain", "softmax_xent", "eltwise", "exchange", "head", "headprep"};
    auto hits = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
      for (auto x : a)
        for (auto y : b)
          if (x == y) return true;
      return false;
    };
    int node_index = 0;
    for (auto& n : p.nodes) {
      s += std::string("  #") + std::to_string(node_index) + " L" + std::to_string(n.level) + " " + kinds[n.kind] + " " + n.label;
      s += " <-";
      for (int j = 0; j < node_index; ++j) {
        const Node& a2 = p.nodes[j];
        if (hits(a2.writes, n.reads) || hits(a2.writes, n.writes) || hits(a2.reads, n.writes)) s += " #" + std::to_string(j);
      }
      ++node_index;
      if (n.kind == Node::INTERP) {
        static const char* paths[] = {"", " 4wide-eltwise", " 4wide-stream-reduce", " 4wide-point-reduce"};
        s += " points=" + std::to_string(n.ip.npoints) + " red=" + std::to_string(n.ip.nred) + " loops=" +
             std::to_string(n.ip.nloops);
        if (n.ip.vec4 && !n.strict) s += paths[n.ip.vec4 & 3];
        else s += " pb=" + std::to_string(n.pb) + " rb=" + std::to_string(n.rb) + " rsplit=" + std::to_string(n.rsplit);
        s += std::string(n.ip.accumulate ? " +=" : " =") + (n.ip.scatter ? " scatter" : "");
      }
      if (n.kind == Node::GEMM)
        s += " M=" + std::to_string(n.gemm.M) + " N=" + std::to_string(n.gemm.N) + " K=" + std::to_string(n.gemm.K) +
             ((n.gemm.flags & GEMM_ACCUMULATE) ? " +=" : " =");
      s += "\n";
    }
    for (auto& note : p.notes) s += "  note: " + note + "\n";
  }
  copy_out(s, buf, cap, needed);
  EGB_CATCH
}

}  // extern "C"
