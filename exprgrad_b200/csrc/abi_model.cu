// C-ABI group 3: program + model. Replaces, for CompileGpu targets, exprgrad's
//   compile[T]            exprgrad/model.nim:270-273  (egb_program_parse/compile + egb_model_create)
//   Model.call/apply      exprgrad/model.nim:392-411  (egb_model_call [+ egb_model_read_output])
//   Model.fit             exprgrad/model.nim:413-454  (egb_model_fit)
//   model.params/caches   exprgrad/model.nim:37-38    (egb_model_write_tensor / read_tensor)
//   inferShapes           exprgrad/passes.nim:1386-1436 (egb_program_infer_shapes)
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
  } else if (k == "fuse") {
    if (m->m->fuse != (value != 0)) {
      EGB_CUDA(cudaStreamSynchronize(m->ctx->c.stream));
      m->m->plans.clear();
      m->m->last_plan = nullptr;
    }
    m->m->fuse = value != 0;
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
  Plan& plan = model.get_plan(target, a.ids, a.shapes);
  bool rebind = false;
  for (auto& kv : plan.bound)
    if (kv.second) {
      kv.second = nullptr;
      rebind = true;
    }
  if (rebind) model.build_nodes(plan);
  model.epoch += 1;
  for (int64_t b = 0; b < batch_count; ++b) {
    for (int i = 0; i < n_args; ++i) {
      auto t = plan.tensors.find(a.ids[i]);
      if (t == plan.tensors.end() || !t->second.bytes) continue;
      const char* src = (const char*)data[i] + (size_t)b * batch_size * row_bytes[i];
      EGB_CUDA(cudaMemcpyAsync(t->second.ptr, src, t->second.bytes, cudaMemcpyHostToDevice, c.stream));
    }
    model.run(plan);
  }
  if (batches_run) *batches_run = batch_count;
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
    static const char* kinds[] = {"interp", "gemm", "split", "memset", "random", "allreduce", "conv"};
    for (auto& n : p.nodes) {
      s += std::string("  ") + kinds[n.kind] + " " + n.label;
      if (n.kind == Node::INTERP)
        s += " points=" + std::to_string(n.ip.npoints) + " red=" + std::to_string(n.ip.nred) + " pb=" +
             std::to_string(n.pb) + " rb=" + std::to_string(n.rb) + (n.ip.accumulate ? " +=" : " =") +
             (n.ip.scatter ? " scatter" : "");
      if (n.kind == Node::GEMM)
        s += " M=" + std::to_string(n.gemm.M) + " N=" + std::to_string(n.gemm.N) + " K=" + std::to_string(n.gemm.K) +
             ((n.gemm.flags & GEMM_ACCUMULATE) ? " +=" : " =");
      s += "\n";
    }
  }
  copy_out(s, buf, cap, needed);
  EGB_CATCH
}

}  // extern "C"
