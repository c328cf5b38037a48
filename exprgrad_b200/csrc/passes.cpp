// Semantics-defining middle-end passes and run-time shape inference of the B200 backend.
//
// exprgrad's own pipeline (exprgrad/model.nim:46-77) mixes passes that decide WHAT is computed
// (autodiff, dead-kernel elimination, loop bounds, shape constraints) with CPU/OpenCL scheduling
// (reorder/fuse/tile/cache/LICM). This file implements the first group so that a program handed
// over in source form (stage 0) can be prepared natively; a program that the Nim side has already
// run through passes.nim (stage 1) skips it. Shape inference (passes.nim:1386-1436) runs on every
// call in the reference and is integer-exact here.
//
//   fold_linear_indices      passes.nim:195-253      dead_code_elim     passes.nim:268-329
//   deduplicate_reads        passes.nim:352-381      shape constraints  passes.nim:1040-1117
//   derive / generate        passes.nim:383-698      dead_kernel_elim   passes.nim:331-350
//   infer_loop_bounds        passes.nim:1001-1038    independent loops  passes.nim:1774-1781
//   collect_tensors          passes.nim:936-967      sort constraints   passes.nim:1119-1221
//   solve / eval / infer_shapes  passes.nim:1252-1436
#include <math.h>

#include <algorithm>
#include <functional>
#include <set>

#include "egb_internal.hpp"
#include "program.hpp"

namespace egb {

namespace {

Instr mk(Op op, std::vector<int> args, int res) {
  Instr i;
  i.op = op;
  i.args = std::move(args);
  i.res = res;
  return i;
}

// ------------------------------------------------------------------ folding / DCE / dedup

// Symbolic evaluation of an index's setup instructions into one LinearIndex over loop iterators
// and opaque (non-linear) registers.
LinearIndex fold_setup(const LinearIndex& index, const Kernel& k) {
  std::map<int, LinearIndex> regs;
  for (auto& l : k.loops) regs[l.iter] = LinearIndex::from_reg(l.iter);
  auto get = [&](int r) -> const LinearIndex& {
    auto it = regs.find(r);
    if (it == regs.end()) fail(EGB_ERR_PARSER, "index expression uses undefined register %d", r);
    return it->second;
  };
  for (auto& ins : index.setup) {
    switch (ins.op) {
      case Op::Index: regs[ins.res] = LinearIndex::from_const(ins.index); break;
      case Op::Add: regs[ins.res] = get(ins.args[0]).plus(get(ins.args[1])); break;
      case Op::Sub: regs[ins.res] = get(ins.args[0]).minus(get(ins.args[1])); break;
      case Op::Negate: regs[ins.res] = get(ins.args[0]).scaled(-1); break;
      case Op::Mul: {
        const LinearIndex& a = get(ins.args[0]);
        const LinearIndex& b = get(ins.args[1]);
        if (a.factors.empty()) regs[ins.res] = b.scaled(a.constant);
        else if (b.factors.empty()) regs[ins.res] = a.scaled(b.constant);
        else regs[ins.res] = LinearIndex::from_reg(ins.res);
        break;
      }
      default: regs[ins.res] = LinearIndex::from_reg(ins.res); break;
    }
  }
  LinearIndex total;
  total.constant = index.constant;
  for (auto& kv : index.factors) total = total.plus(get(kv.first).scaled(kv.second));
  total.setup.clear();
  std::set<int> used;
  for (auto& kv : total.factors) used.insert(kv.first);
  std::vector<Instr> kept;
  for (auto it = index.setup.rbegin(); it != index.setup.rend(); ++it) {
    if (used.count(it->res)) {
      kept.push_back(*it);
      for (int a : it->args) used.insert(a);
    }
  }
  std::reverse(kept.begin(), kept.end());
  total.setup = kept;
  return total;
}

void fold_linear_indices(Kernel& k) {
  for (auto& l : k.loops) {
    l.start = fold_setup(l.start, k);
    l.stop = fold_setup(l.stop, k);
  }
  for (auto& r : k.reads)
    for (auto& d : r.dims) d = fold_setup(d, k);
  for (auto& d : k.write.dims) d = fold_setup(d, k);
}

std::vector<Instr> dce_instrs(const std::vector<Instr>& instrs, std::set<int>& used) {
  std::vector<Instr> out;
  for (auto it = instrs.rbegin(); it != instrs.rend(); ++it) {
    if (it->res != 0 && used.count(it->res)) {
      for (int a : it->args) used.insert(a);
      out.push_back(*it);
    }
  }
  std::reverse(out.begin(), out.end());
  return out;
}

void dce_index(LinearIndex& idx, std::set<int>& used) {
  for (auto& kv : idx.factors) used.insert(kv.first);
  idx.setup = dce_instrs(idx.setup, used);
}

void dead_code_elim(Kernel& k) {
  if (k.is_generator()) return;
  std::set<int> used;
  if (k.write.data) used.insert(k.write.data);
  for (auto& d : k.write.dims) dce_index(d, used);
  k.instrs = dce_instrs(k.instrs, used);
  std::vector<TensorOp> kept;
  for (auto& r : k.reads) {
    if (used.count(r.data)) {
      for (auto& d : r.dims) dce_index(d, used);
      kept.push_back(r);
    }
  }
  k.reads = kept;
  for (auto it = k.loops.rbegin(); it != k.loops.rend(); ++it) {
    dce_index(it->start, used);
    dce_index(it->stop, used);
  }
}

bool same_op_without_data(const TensorOp& a, const TensorOp& b) {
  if (a.tensor != b.tensor || a.is_raw != b.is_raw || a.dims.size() != b.dims.size()) return false;
  for (size_t i = 0; i < a.dims.size(); ++i)
    if (!a.dims[i].same_as(b.dims[i])) return false;
  return true;
}

void deduplicate_reads(Kernel& k) {
  std::map<int, int> subs;
  std::vector<TensorOp> kept;
  for (auto& r : k.reads) {
    bool dup = false;
    for (auto& u : kept) {
      if (same_op_without_data(u, r)) {
        subs[r.data] = u.data;
        dup = true;
        break;
      }
    }
    if (!dup) kept.push_back(r);
  }
  k.reads = kept;
  auto sub = [&](int& r) {
    auto it = subs.find(r);
    if (it != subs.end()) r = it->second;
  };
  for (auto& ins : k.instrs)
    for (int& a : ins.args) sub(a);
  sub(k.res);
  sub(k.write.data);
}

void for_all_kernels(Kernel& k, const std::function<void(Kernel&)>& fn) {
  fn(k);
  if (k.custom_grad)
    for (auto& g : k.custom_grad->kernels) fn(*g);
}

// ------------------------------------------------------------------ shape constraints

// Per dim keep, for every distinct factor table, the largest constant (passes.nim:1040-1057).
std::vector<LinearIndex> simplify_max_index(const std::vector<LinearIndex>& indices) {
  std::vector<LinearIndex> complex_, simple;
  for (auto& idx : indices) {
    if (!idx.setup.empty()) {
      complex_.push_back(idx);
      continue;
    }
    bool found = false;
    for (auto& s : simple) {
      if (s.factors == idx.factors) {
        s.constant = std::max(s.constant, idx.constant);
        found = true;
        break;
      }
    }
    if (!found) {
      LinearIndex li;
      li.factors = idx.factors;
      li.constant = idx.constant;
      simple.push_back(li);
    }
  }
  complex_.insert(complex_.end(), simple.begin(), simple.end());
  return complex_;
}

void kernel_shape_constraints(const Kernel& k, std::vector<ShapeConstraint>& out) {
  if (k.write.is_raw) {
    if (k.reads.size() == 1) {
      ShapeConstraint sc;
      sc.kind = ShapeKind::Copy;
      sc.dest = k.write.tensor;
      sc.priority = PRIO_INFERRED;
      sc.src = k.reads[0].tensor;
      out.push_back(sc);
    }
  } else {
    ShapeConstraint lin;
    lin.kind = ShapeKind::Linear;
    lin.dest = k.write.tensor;
    lin.priority = PRIO_INFERRED;
    for (auto& op : k.reads) {
      if (op.is_raw) continue;
      size_t pos = 0;
      for (; pos < lin.reads.size(); ++pos)
        if (lin.reads[pos].first == op.tensor) break;
      if (pos == lin.reads.size())
        lin.reads.emplace_back(op.tensor, std::vector<std::vector<LinearIndex>>(op.dims.size()));
      auto& dims = lin.reads[pos].second;
      if (dims.size() < op.dims.size()) dims.resize(op.dims.size());
      for (size_t i = 0; i < op.dims.size(); ++i) dims[i].push_back(op.dims[i]);
    }
    lin.write = k.write.dims;
    for (auto& rd : lin.reads)
      for (auto& dim : rd.second) dim = simplify_max_index(dim);
    out.push_back(lin);
  }
  auto rank_of = [&](const TensorOp& op) {
    if (op.is_raw) return;
    ShapeConstraint sc;
    sc.kind = ShapeKind::Rank;
    sc.dest = op.tensor;
    sc.priority = PRIO_CONDITION;
    sc.rank = (int)op.dims.size();
    out.push_back(sc);
  };
  for (auto& r : k.reads) rank_of(r);
  rank_of(k.write);
}

void infer_shape_constraints(Program& prog) {
  for (auto& t : prog.targets) {
    for (int tid : prog.caches) {
      ShapeConstraint sc;
      sc.kind = ShapeKind::Copy;
      sc.dest = tid;
      sc.priority = PRIO_INFERRED;
      sc.src = prog.tdef(tid).cache;
      t->shapes.push_back(sc);
    }
    for (auto& k : t->kernels)
      if (!k->is_generator()) kernel_shape_constraints(*k, t->shapes);
  }
}

// ------------------------------------------------------------------ autodiff

// Reverse sweep over one expression (passes.nim:383-517): one adjoint register per primal register,
// a register used twice receives the sum of both contributions.
std::vector<Instr> derive_instrs(const std::vector<Instr>& instrs, Kernel& k, std::map<int, int>& grad) {
  std::vector<Instr> out;
  auto emit = [&](Op op, std::vector<int> args) {
    int r = k.alloc_reg();
    out.push_back(mk(op, std::move(args), r));
    return r;
  };
  auto lit = [&](double v) {
    int r = k.alloc_reg();
    Instr i = mk(Op::Scalar, {}, r);
    i.scalar = v;
    out.push_back(i);
    return r;
  };
  for (auto it = instrs.rbegin(); it != instrs.rend(); ++it) {
    const Instr& ins = *it;
    auto gi = grad.find(ins.res);
    if (gi == grad.end()) continue;
    const int g = gi->second;
    const std::vector<int>& a = ins.args;
    std::vector<int> ga;
    switch (ins.op) {
      case Op::Add: ga = {g, g}; break;
      case Op::Sub: ga = {g, emit(Op::Negate, {g})}; break;
      case Op::Mul: {
        int g0 = emit(Op::Mul, {g, a[1]});
        int g1 = emit(Op::Mul, {g, a[0]});
        ga = {g0, g1};
        break;
      }
      case Op::Div: {
        int grad_a = emit(Op::Div, {g, a[1]});
        int sq_y = emit(Op::Mul, {a[1], a[1]});
        int div_g = emit(Op::Div, {g, sq_y});
        int neg_x = emit(Op::Negate, {a[0]});
        ga = {grad_a, emit(Op::Mul, {neg_x, div_g})};
        break;
      }
      case Op::Negate: ga = {emit(Op::Negate, {g})}; break;
      case Op::Ln: case Op::Log10: case Op::Log2: {
        int den = a[0];
        if (ins.op != Op::Ln) {
          int f = lit(ins.op == Op::Log10 ? log(10.0) : log(2.0));
          den = emit(Op::Mul, {a[0], f});
        }
        ga = {emit(Op::Div, {g, den})};
        break;
      }
      case Op::Log: {
        int log_y = emit(Op::Ln, {a[1]});
        int mul = emit(Op::Mul, {a[0], log_y});
        int gx = emit(Op::Div, {g, mul});
        int log_x = emit(Op::Ln, {a[0]});
        int neg_log_x = emit(Op::Negate, {log_x});
        int log_y_sq = emit(Op::Mul, {log_y, log_y});
        int den = emit(Op::Mul, {a[1], log_y_sq});
        int num = emit(Op::Mul, {g, neg_log_x});
        ga = {gx, emit(Op::Div, {num, den})};
        break;
      }
      case Op::Exp: ga = {emit(Op::Mul, {g, ins.res})}; break;
      case Op::Sin: {
        int c = emit(Op::Cos, {a[0]});
        ga = {emit(Op::Mul, {c, g})};
        break;
      }
      case Op::Cos: {
        int s = emit(Op::Sin, {a[0]});
        int ns = emit(Op::Negate, {s});
        ga = {emit(Op::Mul, {ns, g})};
        break;
      }
      case Op::Select: {
        int zero = lit(0.0);
        int g1 = emit(Op::Select, {a[0], g, zero});
        int g2 = emit(Op::Select, {a[0], zero, g});
        ga = {0, g1, g2};
        break;
      }
      case Op::Sqrt: {
        int two = lit(2.0);
        int den = emit(Op::Mul, {two, ins.res});
        ga = {emit(Op::Div, {g, den})};
        break;
      }
      case Op::Pow: {
        int one = lit(1.0);
        int new_exp = emit(Op::Sub, {a[1], one});
        int p = emit(Op::Pow, {a[0], new_exp});
        int pf = emit(Op::Mul, {a[1], p});
        int g_base = emit(Op::Mul, {g, pf});
        int lg = emit(Op::Ln, {a[0]});
        int prod = emit(Op::Mul, {ins.res, lg});
        ga = {g_base, emit(Op::Mul, {g, prod})};
        break;
      }
      case Op::ToScalar: case Op::ToIndex: ga = {0}; break;
      default: break;  // no rule: fine for literals (no arguments), an error otherwise
    }
    if (ga.size() != a.size())
      fail(EGB_ERR_GRADIENT, "Unable to derive %s", op_name(ins.op));
    for (size_t i = 0; i < a.size(); ++i) {
      if (ga[i] == 0) continue;
      auto f = grad.find(a[i]);
      if (f != grad.end()) f->second = emit(Op::Add, {f->second, ga[i]});
      else grad[a[i]] = ga[i];
    }
  }
  return out;
}

// One adjoint kernel per read that receives gradient (passes.nim:519-549).
std::vector<std::shared_ptr<Kernel>> derive_kernel(const Kernel& k, const std::map<int, int>& grad_tensors) {
  auto base = k.clone();
  std::map<int, int> grad_regs;
  const int write_grad = base->alloc_reg();
  TensorOp gread;
  gread.tensor = grad_tensors.at(k.write.tensor);
  gread.is_raw = k.write.is_raw;
  gread.dims = k.write.dims;
  gread.data = write_grad;
  base->reads.push_back(gread);
  grad_regs[k.write.data] = write_grad;
  auto extra = derive_instrs(k.instrs, *base, grad_regs);
  base->instrs.insert(base->instrs.end(), extra.begin(), extra.end());
  std::vector<std::shared_ptr<Kernel>> out;
  for (auto& read : k.reads) {
    auto gi = grad_regs.find(read.data);
    if (gi == grad_regs.end()) continue;
    auto gk = base->clone();
    gk->res = gi->second;
    gk->write.tensor = grad_tensors.at(read.tensor);
    gk->write.is_raw = read.is_raw;
    gk->write.dims = read.dims;
    gk->write.data = gi->second;
    dead_code_elim(*gk);
    out.push_back(gk);
  }
  return out;
}

// `dst{i} = 1.0` (seed) or `dst{i} = src{i}` (reshape) over len(src) (passes.nim:574-600, 643-673).
std::shared_ptr<Kernel> iota_kernel(int len_tensor, int write_tensor, int read_tensor, double lit) {
  auto k = std::make_shared<Kernel>();
  k->nregs = 3;
  Loop l;
  l.iter = 2;
  l.has_bounds = true;
  l.start = LinearIndex::from_const(0);
  Instr len = mk(Op::Len, {}, 3);
  len.tensor = len_tensor;
  l.stop.setup.push_back(len);
  l.stop.factors[3] = 1;
  l.step = 1;
  k->loops.push_back(l);
  if (read_tensor) {
    TensorOp r;
    r.tensor = read_tensor;
    r.is_raw = true;
    r.dims.push_back(LinearIndex::from_reg(2));
    r.data = 1;
    k->reads.push_back(r);
  } else {
    Instr s = mk(Op::Scalar, {}, 1);
    s.scalar = lit;
    k->instrs.push_back(s);
  }
  k->res = 1;
  k->write.tensor = write_tensor;
  k->write.is_raw = true;
  k->write.dims.push_back(LinearIndex::from_reg(2));
  k->write.data = 1;
  return k;
}

ShapeConstraint copy_constraint(int dest, int src, int prio) {
  ShapeConstraint sc;
  sc.kind = ShapeKind::Copy;
  sc.dest = dest;
  sc.src = src;
  sc.priority = prio;
  return sc;
}

void generate(Program& prog) {
  for (auto& tp : prog.targets) {
    Target& target = *tp;
    size_t it = 0;
    while (it < target.kernels.size()) {
      std::shared_ptr<Kernel> kernel = target.kernels[it];
      if (kernel->gen == GenKind::Backwards) {
        std::map<int, int> grad_tensors;
        std::vector<std::shared_ptr<Kernel>> grad_kernels;
        const int loss = kernel->gen_tensor;
        TensorDef td;
        td.kind = TensorKind::Result;
        const int grad_loss = prog.alloc_tensor(td);
        grad_kernels.push_back(iota_kernel(loss, grad_loss, 0, 1.0));
        target.shapes.push_back(copy_constraint(grad_loss, loss, PRIO_INFERRED));
        grad_tensors[loss] = grad_loss;
        for (size_t j = it + 1; j < target.kernels.size(); ++j) {
          Kernel& k2 = *target.kernels[j];
          if (k2.gen == GenKind::Gradient) {
            grad_tensors[k2.gen_tensor] = k2.write.tensor;
            target.shapes.push_back(copy_constraint(k2.write.tensor, k2.gen_tensor, PRIO_INFERRED));
          }
        }
        for (size_t j = it; j-- > 0;) {
          Kernel& k2 = *target.kernels[j];
          for (auto& read : k2.reads) {
            if (!grad_tensors.count(read.tensor)) {
              TensorDef g;
              g.kind = TensorKind::Result;
              const int gt = prog.alloc_tensor(g);
              target.shapes.push_back(copy_constraint(gt, read.tensor, PRIO_INFERRED));
              grad_tensors[read.tensor] = gt;
            }
          }
          if (k2.custom_grad) {
            std::map<int, int> subs = k2.custom_grad->subs;
            for (auto& kv : k2.custom_grad->tensors) {
              int t = kv.first;
              auto s = k2.custom_grad->subs.find(t);
              if (s != k2.custom_grad->subs.end()) t = s->second;
              subs[kv.second] = grad_tensors.at(t);
            }
            for (auto ck = k2.custom_grad->kernels.rbegin(); ck != k2.custom_grad->kernels.rend(); ++ck) {
              auto c = (*ck)->clone();
              c->substitute_tensors(subs);
              grad_kernels.push_back(c);
            }
          } else {
            if (k2.is_generator()) continue;
            if (!grad_tensors.count(k2.write.tensor)) continue;
            auto ks = derive_kernel(k2, grad_tensors);
            grad_kernels.insert(grad_kernels.end(), ks.begin(), ks.end());
          }
        }
        for (auto& gt : grad_tensors) prog.grad_tensors[target.name][gt.first] = gt.second;  // (several backwards() per target merge)
        target.kernels.erase(target.kernels.begin() + it);
        target.kernels.insert(target.kernels.begin() + it, grad_kernels.begin(), grad_kernels.end());
        it += grad_kernels.size();
      } else if (kernel->gen == GenKind::Gradient) {
        target.kernels.erase(target.kernels.begin() + it);
      } else if (kernel->gen == GenKind::Reshape) {
        const int src = kernel->gen_tensor, dst = kernel->write.tensor;
        target.kernels[it] = iota_kernel(src, dst, src, 0.0);
        ShapeConstraint sc;
        sc.kind = ShapeKind::Dims;
        sc.dest = dst;
        sc.priority = PRIO_INFERRED;
        int64_t prod = 1;
        for (auto s : kernel->reshape)
          if (s >= 0) prod *= s;
        for (auto s : kernel->reshape) {
          if (s >= 0) {
            sc.dims.push_back(LinearIndex::from_const(s));
          } else {
            LinearIndex li;
            Instr len = mk(Op::Len, {}, 1);
            len.tensor = src;
            Instr c = mk(Op::Index, {}, 2);
            c.index = prod;
            li.setup = {len, c, mk(Op::IndexDiv, {1, 2}, 3)};
            li.factors[3] = 1;
            sc.dims.push_back(li);
          }
        }
        target.shapes.push_back(sc);
        ++it;
      } else {
        ++it;
      }
    }
  }
}

void dead_kernel_elim(Program& prog) {
  for (auto& t : prog.targets) {
    std::set<int> used;
    for (size_t i = 0; i < prog.tensors.size(); ++i)
      if (prog.tensors[i].kind != TensorKind::Result) used.insert((int)i + 1);
    if (t->output) used.insert(t->output);
    std::vector<std::shared_ptr<Kernel>> kept;
    for (auto it = t->kernels.rbegin(); it != t->kernels.rend(); ++it) {
      if (used.count((*it)->write.tensor)) {
        for (auto& r : (*it)->reads) used.insert(r.tensor);
        kept.push_back(*it);
      }
    }
    std::reverse(kept.begin(), kept.end());
    t->kernels = kept;
  }
}

// ------------------------------------------------------------------ loops

void infer_loop_bounds(Program& prog) {
  for (auto& t : prog.targets) {
    for (auto& kp : t->kernels) {
      Kernel& k = *kp;
      auto visit = [&](const TensorOp& op) {
        for (size_t d = 0; d < op.dims.size(); ++d) {
          const int reg = op.dims[d].only_register();
          if (!reg) continue;
          for (auto& loop : k.loops) {
            if (loop.iter != reg || loop.has_bounds) continue;
            loop.has_bounds = true;
            loop.start = LinearIndex::from_const(0);
            const int size = k.alloc_reg();
            Instr s = mk(op.is_raw ? Op::Len : Op::Shape, {}, size);
            s.tensor = op.tensor;
            s.dim = op.is_raw ? 0 : (int)d;
            loop.stop = LinearIndex();
            loop.stop.setup.push_back(s);
            loop.stop.factors[size] = 1;
            loop.step = 1;
          }
        }
      };
      for (auto& r : k.reads) visit(r);
      visit(k.write);
    }
  }
}

void identify_independent(Program& prog) {
  for (auto& t : prog.targets) {
    for (auto& kp : t->kernels) {
      std::set<int> indep;
      for (auto& d : kp->write.dims)
        if (int r = d.only_register()) indep.insert(r);
      for (auto& l : kp->loops)
        if (indep.count(l.iter)) l.mode = 1;
    }
  }
}

// Greedy topological order on "dim i-1 register -> dim i register" edges, reads weigh 10, the write 1
// (passes.nim:700-745). It fixes the nesting - and therefore the fp accumulation order - of reductions.
void reorder_loops(Kernel& k) {
  const int n = (int)k.loops.size();
  std::map<int, int> loop_of;
  for (int i = 0; i < n; ++i) loop_of[k.loops[i].iter] = i;
  std::vector<std::vector<std::pair<int, int>>> graph(n);  // (target, weight)
  auto visit = [&](const TensorOp& op, int weight) {
    for (size_t i = 1; i < op.dims.size(); ++i)
      for (auto& ra : op.dims[i - 1].factors)
        for (auto& rb : op.dims[i].factors)
          if (loop_of.count(ra.first) && loop_of.count(rb.first))
            graph[loop_of[ra.first]].emplace_back(loop_of[rb.first], weight);
  };
  for (auto& r : k.reads) visit(r, 10);
  visit(k.write, 1);
  std::vector<int> scores(n, 0);
  for (auto& edges : graph)
    for (auto& e : edges) scores[e.first] += e.second;
  std::vector<bool> closed(n, false);
  std::vector<Loop> order;
  for (int step = 0; step < n; ++step) {
    int best = -1;
    for (int i = 0; i < n; ++i)
      if (!closed[i] && (best < 0 || scores[i] < scores[best])) best = i;
    closed[best] = true;
    order.push_back(k.loops[best]);
    for (auto& e : graph[best]) scores[e.first] -= e.second;
  }
  k.loops = order;
}

// ------------------------------------------------------------------ target bookkeeping

void add_unique(std::vector<int>& v, int x) {
  if (x && std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x);
}

void collect_tensors(Program& prog) {
  for (auto& t : prog.targets) {
    std::vector<int> ts;
    for (auto& kp : t->kernels) {
      for (auto& r : kp->reads) add_unique(ts, r.tensor);
      add_unique(ts, kp->write.tensor);
      for (auto& l : kp->loops) {
        for (auto& i : l.start.setup) add_unique(ts, i.tensor);
        for (auto& i : l.stop.setup) add_unique(ts, i.tensor);
      }
      for (auto& i : kp->instrs) add_unique(ts, i.tensor);
    }
    t->tensors = ts;
  }
}

bool underconstrained(const ShapeConstraint& sc) {
  if (sc.kind == ShapeKind::Rank) return sc.rank > 0;
  if (sc.kind != ShapeKind::Linear) return false;
  std::set<int> defined;
  for (auto& rd : sc.reads)
    for (auto& dim : rd.second)
      for (auto& idx : dim)
        for (auto& kv : idx.factors) defined.insert(kv.first);
  for (auto& d : sc.write)
    for (auto& kv : d.factors)
      if (!defined.count(kv.first)) return true;
  return false;
}

std::vector<int> constraint_deps(const ShapeConstraint& sc) {
  std::vector<int> deps;
  if (sc.kind == ShapeKind::Dims) {
    for (auto& d : sc.dims)
      for (auto& i : d.setup)
        if (i.tensor) deps.push_back(i.tensor);
  } else if (sc.kind == ShapeKind::Linear) {
    for (auto& rd : sc.reads) deps.push_back(rd.first);
  } else if (sc.kind == ShapeKind::Copy) {
    deps.push_back(sc.src);
  }
  return deps;
}

void sort_shape_constraints(Program& prog) {
  for (auto& tp : prog.targets) {
    Target& target = *tp;
    std::map<int, ShapeConstraint> chosen;
    std::vector<ShapeConstraint> conditions;
    for (auto& sc : target.shapes) {
      auto it = chosen.find(sc.dest);
      if (it == chosen.end()) chosen[sc.dest] = sc;  // first wins on ties (passes.nim:1182-1185)
      else if (it->second.priority < sc.priority) it->second = sc;
      if (sc.priority == PRIO_CONDITION) conditions.push_back(sc);
    }
    for (auto& cond : conditions) {
      auto ci = chosen.find(cond.dest);
      if (ci == chosen.end()) continue;
      ShapeConstraint sc = ci->second;
      while (sc.kind == ShapeKind::Copy && chosen.count(sc.src) && prog.tdef(sc.dest).shape.empty())
        sc = chosen[sc.src];
      if (sc.kind == ShapeKind::Copy && prog.tdef(sc.dest).shape.empty()) {
        chosen[sc.src] = cond;
      } else {
        int rank = -1;
        if (!prog.tdef(sc.dest).shape.empty()) rank = (int)prog.tdef(sc.dest).shape.size();
        else if (sc.kind == ShapeKind::Dims) rank = (int)sc.dims.size();
        else if (sc.kind == ShapeKind::Linear) rank = (int)sc.write.size();
        else if (sc.kind == ShapeKind::Rank) rank = sc.rank;
        if (cond.rank != rank)
          fail(EGB_ERR_SHAPE, "A condition requires that tensor%d has rank %d, but it has rank %d", cond.dest - 1,
               cond.rank, rank);
      }
    }
    std::vector<ShapeConstraint> order;
    std::set<int> closed;
    std::function<void(int)> visit = [&](int tid) {
      const TensorKind kind = prog.tdef(tid).kind;
      if ((kind == TensorKind::Result || kind == TensorKind::Cache || kind == TensorKind::Random) &&
          !closed.count(tid)) {
        closed.insert(tid);
        auto it = chosen.find(tid);
        if (it == chosen.end())
          fail(EGB_ERR_SHAPE, "tensor%d (%s) requires shape", tid - 1, prog.tdef(tid).name.c_str());
        const ShapeConstraint sc = it->second;
        if (underconstrained(sc)) fail(EGB_ERR_SHAPE, "Shape for tensor%d is underconstrained", tid - 1);
        for (int dep : constraint_deps(sc)) visit(dep);
        order.push_back(sc);
      }
    };
    for (int tid : target.tensors) visit(tid);
    target.shapes = order;
  }
}

}  // namespace

// ------------------------------------------------------------------ pipeline

void compile_program(Program& prog) {
  if (prog.compiled) return;
  prog.params.clear();
  prog.caches.clear();
  prog.inputs.clear();
  for (size_t i = 0; i < prog.tensors.size(); ++i) {  // makeTensorLookups (passes.nim:1745-1758)
    const TensorDef& t = prog.tensors[i];
    if (t.kind == TensorKind::Param) prog.params.push_back((int)i + 1);
    else if (t.kind == TensorKind::Input) prog.inputs[t.name] = (int)i + 1;
    else if (t.kind == TensorKind::Cache) prog.caches.push_back((int)i + 1);
  }
  for (auto& t : prog.targets)
    for (auto& k : t->kernels)
      for_all_kernels(*k, [](Kernel& x) {
        dead_code_elim(x);
        fold_linear_indices(x);
        deduplicate_reads(x);
      });
  infer_shape_constraints(prog);
  generate(prog);
  dead_kernel_elim(prog);
  infer_loop_bounds(prog);
  identify_independent(prog);
  dead_kernel_elim(prog);
  collect_tensors(prog);
  sort_shape_constraints(prog);
  // static shapes of caches (subset of inferStaticShapes, passes.nim:1444-1514)
  for (int tid : prog.caches) {
    const TensorDef& src = prog.tdef(prog.tdef(tid).cache);
    bool known = !src.shape.empty();
    for (auto s : src.shape) known = known && s >= 0;
    if (!known)
      fail(EGB_ERR_SHAPE, "Shape of cache \"%s\" must be inferred at compile time", prog.tdef(tid).name.c_str());
    prog.tdef(tid).shape = src.shape;
  }
  for (auto& t : prog.targets)
    for (auto& k : t->kernels) reorder_loops(*k);
  prog.compiled = true;
}

// ------------------------------------------------------------------ run-time shape inference

namespace {

struct Frac {  // exact rational over int64 (passes.nim uses Rational[int])
  int64_t n = 0, d = 1;
  static int64_t gcd(int64_t a, int64_t b) {
    a = a < 0 ? -a : a;
    b = b < 0 ? -b : b;
    while (b) {
      int64_t t = a % b;
      a = b;
      b = t;
    }
    return a ? a : 1;
  }
  Frac() {}
  Frac(int64_t num, int64_t den = 1) {
    if (den < 0) {
      num = -num;
      den = -den;
    }
    int64_t g = gcd(num, den);
    n = num / g;
    d = den / g;
  }
  Frac operator-(const Frac& o) const { return Frac(n * o.d - o.n * d, d * o.d); }
  Frac operator*(const Frac& o) const { return Frac(n * o.n, d * o.d); }
  Frac operator/(const Frac& o) const { return Frac(n * o.d, d * o.n); }
  bool operator==(const Frac& o) const { return n == o.n && d == o.d; }
};

int64_t trunc_div(int64_t a, int64_t b) { return a / b; }  // C++ '/' truncates toward zero like Nim `div`
int64_t trunc_mod(int64_t a, int64_t b) { return a % b; }

// Exact solve of `index = 0` equations: first n distinct (normalised) rows, fraction-free
// elimination with partial pivoting, rational back-substitution (passes.nim:1252-1323).
std::map<int, Frac> solve(const std::vector<LinearIndex>& eqs) {
  std::vector<int> order;
  std::map<int, int> indices;
  for (auto& eq : eqs)
    for (auto& kv : eq.factors)
      if (!indices.count(kv.first)) {
        indices[kv.first] = (int)order.size();
        order.push_back(kv.first);
      }
  const int n = (int)order.size();
  std::map<int, Frac> result;
  if (n == 0) return result;
  if ((int)eqs.size() < n) fail(EGB_ERR_VALUE, "Underconstrained linear system");
  std::vector<std::vector<int64_t>> rows;
  std::vector<std::vector<Frac>> known;
  for (auto& eq : eqs) {
    if (eq.factors.empty()) {
      if (eq.constant != 0) fail(EGB_ERR_VALUE, "No solution");
      continue;
    }
    std::vector<int64_t> row(n + 1, 0);
    for (auto& kv : eq.factors) row[indices[kv.first]] = kv.second;
    row[n] = -eq.constant;
    int64_t first = 0;
    std::vector<Frac> norm;
    for (auto v : row) {
      if (first == 0) first = v;
      norm.push_back(first == 0 ? Frac(0) : Frac(v, first));
    }
    bool dup = false;
    for (auto& kn : known)
      if (kn == norm) {
        dup = true;
        break;
      }
    if (dup) continue;
    known.push_back(norm);
    rows.push_back(row);
    if ((int)rows.size() >= n) break;
  }
  if ((int)rows.size() < n) fail(EGB_ERR_VALUE, "Underconstrained linear system");
  auto& m = rows;
  for (int pivot = 0; pivot < n; ++pivot) {
    int max_row = pivot;
    for (int y = pivot + 1; y < n; ++y)
      if (llabs(m[y][pivot]) > llabs(m[max_row][pivot])) max_row = y;
    if (max_row != pivot) std::swap(m[max_row], m[pivot]);
    const int64_t tgt = m[pivot][pivot];
    for (int y = pivot + 1; y < n; ++y) {
      const int64_t cur = m[y][pivot];
      if (cur != 0)
        for (int x = 0; x <= n; ++x) m[y][x] = m[y][x] * tgt - m[pivot][x] * cur;
    }
  }
  std::vector<Frac> sol(n);
  for (int y = n - 1; y >= 0; --y) {
    Frac s(m[y][n]);
    for (int x = y + 1; x < n; ++x) s = s - sol[x] * Frac(m[y][x]);
    if (m[y][y] == 0) fail(EGB_ERR_VALUE, "singular shape system");
    sol[y] = s / Frac(m[y][y]);
  }
  for (int i = 0; i < n; ++i) result[order[i]] = sol[i];
  return result;
}

bool matches(const std::vector<int64_t>& stat, const std::vector<int64_t>& shape) {
  if (stat.empty()) return true;
  if (stat.size() != shape.size()) return false;
  for (size_t i = 0; i < stat.size(); ++i)
    if (stat[i] >= 0 && stat[i] != shape[i]) return false;
  return true;
}

std::string shape_str(const std::vector<int64_t>& s) {
  std::string out = "[";
  for (size_t i = 0; i < s.size(); ++i) out += (i ? ", " : "") + std::to_string(s[i]);
  return out + "]";
}

}  // namespace

// Mini evaluator for shape expressions (passes.nim:1328-1374). Returns 0 ok, 1 dynamic shape,
// 2 invalid instruction, 3 missing register.
int eval_index_instrs(const std::vector<Instr>& instrs, const ShapeTable& shapes, std::map<int, int64_t>& regs,
                      int64_t epoch) {
  for (auto& ins : instrs) {
    for (int a : ins.args)
      if (!regs.count(a)) return 3;
    const std::vector<int64_t>* shape = nullptr;
    if (ins.tensor) {
      auto it = shapes.find(ins.tensor);
      if (it == shapes.end()) return 3;
      shape = &it->second;
    }
    auto arg = [&](int i) { return regs[ins.args[i]]; };
    switch (ins.op) {
      case Op::Shape: {
        if (shape->empty()) return 1;
        const int64_t rank = (int64_t)shape->size();
        const int64_t d = ins.dim < 0 ? rank + ins.dim : ins.dim;
        if (d < 0 || d >= rank) return 1;
        if ((*shape)[d] < 0) return 1;
        regs[ins.res] = (*shape)[d];
        break;
      }
      case Op::Len: {
        if (shape->empty()) return 1;
        int64_t p = 1;
        for (auto s : *shape) {
          if (s < 0) return 1;
          p *= s;
        }
        regs[ins.res] = p;
        break;
      }
      case Op::ShapeLen: regs[ins.res] = (int64_t)shape->size(); break;
      case Op::Index: regs[ins.res] = ins.index; break;
      case Op::Add: regs[ins.res] = arg(0) + arg(1); break;
      case Op::Sub: regs[ins.res] = arg(0) - arg(1); break;
      case Op::Mul: regs[ins.res] = arg(0) * arg(1); break;
      case Op::IndexDiv:
        if (arg(1) == 0) return 2;
        regs[ins.res] = trunc_div(arg(0), arg(1));
        break;
      case Op::Mod:
        if (arg(1) == 0) return 2;
        regs[ins.res] = trunc_mod(arg(0), arg(1));
        break;
      case Op::Wrap: {
        if (arg(1) == 0) return 2;
        int64_t v = trunc_mod(arg(0), arg(1));
        regs[ins.res] = v < 0 ? v + arg(1) : v;
        break;
      }
      case Op::Negate: regs[ins.res] = -arg(0); break;
      case Op::Epoch:
        if (epoch < 0) return 2;
        regs[ins.res] = epoch;
        break;
      default: return 2;
    }
  }
  return 0;
}

int64_t eval_linear(const LinearIndex& li, const std::map<int, int64_t>& regs) {
  int64_t v = li.constant;
  for (auto& kv : li.factors) {
    auto it = regs.find(kv.first);
    if (it == regs.end()) fail(EGB_ERR_SHAPE, "Unable to evaluate all instructions.");
    v += kv.second * it->second;
  }
  return v;
}

ShapeTable infer_shapes(const Program& prog, const Target& target, const ShapeTable& inputs) {
  ShapeTable res;
  for (auto& kv : inputs) {
    res[kv.first] = kv.second;
    const auto& stat = prog.tdef(kv.first).shape;
    if (!matches(stat, kv.second))
      fail(EGB_ERR_SHAPE, "Given shape for tensor%d is %s, but its static shape is %s", kv.first - 1,
           shape_str(kv.second).c_str(), shape_str(stat).c_str());
  }
  for (int tid : prog.params) res[tid] = prog.tdef(tid).shape;
  for (auto& sc : target.shapes) {
    for (int dep : constraint_deps(sc))
      if (!res.count(dep))
        fail(EGB_ERR_SHAPE, "Missing shape for tensor%d, maybe you forgot to pass an input to the model?", dep - 1);
    switch (sc.kind) {
      case ShapeKind::Rank: res[sc.dest] = std::vector<int64_t>(sc.rank, 0); break;
      case ShapeKind::Dims: {
        std::vector<int64_t> sizes;
        for (auto& idx : sc.dims) {
          std::map<int, int64_t> regs;
          const int st = eval_index_instrs(idx.setup, res, regs, -1);
          if (st == 1) fail(EGB_ERR_SHAPE, "Not all shapes are known.");
          if (st == 2) fail(EGB_ERR_SHAPE, "Invalid instruction in tensor shape");
          if (st == 3) fail(EGB_ERR_SHAPE, "Unable to evaluate all instructions.");
          sizes.push_back(eval_linear(idx, regs));
        }
        res[sc.dest] = sizes;
        break;
      }
      case ShapeKind::Copy: res[sc.dest] = res[sc.src]; break;
      case ShapeKind::Linear: {
        std::vector<LinearIndex> eqs;
        for (auto& rd : sc.reads) {
          const auto& shape = res[rd.first];
          if (rd.second.size() != shape.size())
            fail(EGB_ERR_SHAPE, "tensor%d is read with %zu indices but has rank %zu", rd.first - 1, rd.second.size(),
                 shape.size());
          for (size_t d = 0; d < rd.second.size(); ++d) {
            if (rd.second[d].size() != 1)
              fail(EGB_ERR_SHAPE, "tensor%d: dimension %zu is indexed in more than one way", rd.first - 1, d);
            eqs.push_back(rd.second[d][0].minus(LinearIndex::from_const(shape[d] - 1)));
          }
        }
        std::map<int, int64_t> max_values;
        for (auto& kv : solve(eqs)) max_values[kv.first] = trunc_div(kv.second.n, kv.second.d);
        std::vector<int64_t> sizes;
        for (auto& idx : sc.write) sizes.push_back(eval_linear(idx, max_values) + 1);
        res[sc.dest] = sizes;
        break;
      }
    }
  }
  return res;
}

}  // namespace egb
