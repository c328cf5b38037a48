// Lowering of one compiled exprgrad kernel (structured form: loops, reads, expression, write -
// exprgrad/ir.nim:211-220) into the device program of the generic loop-nest kernel (interp.hpp),
// plus recognition of the contraction pattern that goes to the tcgen05 GEMM instead.
//
// All shape-dependent quantities (Shape/Len/ShapeLen/Epoch instructions, loop bounds, tensor strides)
// are evaluated here on the host, so the device program only contains per-element arithmetic.
// Constant folding follows the reference's propagateConstants (exprgrad/passes.nim:1614-1706):
// scalar constants fold in float64 and are rounded to T once, with the identities x+0, x*1, x*0 -> 0,
// 0/x -> 0, x/1; everything else is evaluated at run time in fp32 exactly as written.
#include <string.h>

#include <algorithm>

#include "lower.hpp"

namespace egb {

namespace {

enum Ty { T_SCALAR, T_INDEX, T_BOOL, T_ARRAY };
enum VKind { V_CONST, V_HOST, V_DEV };

struct Val {
  VKind kind = V_DEV;
  Ty ty = T_SCALAR;
  double f = 0;   // V_CONST scalar
  int64_t i = 0;  // V_CONST index/bool, V_HOST index
  int slot = -1;  // V_DEV
  uint32_t deps = 0;
  std::vector<int> items;  // T_ARRAY: item registers (for a selected row of a nested array: all rows, flattened)
  int row_len = 0;         // > 0: this is a dynamically selected row of a nested array ...
  int row_slot = -1;       // ... whose row index lives in this slot
};

struct Lowerer {
  const Kernel& k;
  const ShapeTable& shapes;
  int64_t epoch;
  IpProgram ip;
  std::map<int, Val> env;
  std::map<std::pair<int, uint64_t>, int> lit_slots;
  int nslots = 0;
  bool uses_epoch = false;
  std::vector<IpInstr>* sink = nullptr;
  std::vector<IpInstr> index_instrs, body_instrs;
  int array_table_len = 0;

  Lowerer(const Kernel& k_, const ShapeTable& s, int64_t e) : k(k_), shapes(s), epoch(e) { memset(&ip, 0, sizeof(ip)); }

  int new_slot() {
    if (nslots >= IP_MAX_SLOTS) fail(EGB_ERR_GENERATOR, "kernel needs more than %d registers", IP_MAX_SLOTS);
    return nslots++;
  }

  int literal(Ty ty, uint64_t bits) {
    auto key = std::make_pair((int)ty, bits);
    auto it = lit_slots.find(key);
    if (it != lit_slots.end()) return it->second;
    if (ip.nlits >= 48) fail(EGB_ERR_GENERATOR, "kernel uses too many literals");
    const int slot = new_slot();
    ip.lits[ip.nlits] = bits;
    ip.lit_slot[ip.nlits] = (uint8_t)slot;
    ip.nlits++;
    lit_slots[key] = slot;
    return slot;
  }

  int slot_of(const Val& v) {
    if (v.kind == V_DEV) return v.slot;
    if (v.ty == T_SCALAR) {
      const float f = (float)v.f;
      uint32_t b;
      memcpy(&b, &f, 4);
      return literal(T_SCALAR, b);
    }
    if (v.ty == T_ARRAY) fail(EGB_ERR_GENERATOR, "array value used as a scalar");
    return literal(T_INDEX, (uint64_t)v.i);
  }

  const Val& get(int reg) {
    auto it = env.find(reg);
    if (it == env.end()) fail(EGB_ERR_GENERATOR, "register %d is used before it is defined", reg);
    return it->second;
  }

  static Val cscalar(double f) { Val v; v.kind = V_CONST; v.ty = T_SCALAR; v.f = f; return v; }
  static Val cindex(int64_t i) { Val v; v.kind = V_CONST; v.ty = T_INDEX; v.i = i; return v; }
  static Val cbool(bool b) { Val v; v.kind = V_CONST; v.ty = T_BOOL; v.i = b ? 1 : 0; return v; }
  static Val host(int64_t i) { Val v; v.kind = V_HOST; v.ty = T_INDEX; v.i = i; return v; }

  static bool is_zero(const Val& v) {
    return v.kind == V_CONST && v.ty != T_ARRAY && (v.ty == T_SCALAR ? v.f == 0.0 : v.i == 0);
  }
  static bool is_one(const Val& v) {
    return v.kind == V_CONST && v.ty != T_ARRAY && (v.ty == T_SCALAR ? v.f == 1.0 : v.i == 1);
  }

  Val emit(uint8_t op, Ty ty, const std::vector<const Val*>& args, int imm = 0) {
    IpInstr in;
    memset(&in, 0, sizeof(in));
    in.op = op;
    Val r;
    r.kind = V_DEV;
    r.ty = ty;
    uint8_t* dst[3] = {&in.a, &in.b, &in.c};
    for (size_t i = 0; i < args.size() && i < 3; ++i) {
      *dst[i] = (uint8_t)slot_of(*args[i]);
      r.deps |= args[i]->deps;
    }
    r.slot = new_slot();
    in.dst = (uint8_t)r.slot;
    in.imm = (uint16_t)imm;
    sink->push_back(in);
    return r;
  }

  const std::vector<int64_t>& shape_of(int tensor) {
    auto it = shapes.find(tensor);
    if (it == shapes.end()) fail(EGB_ERR_SHAPE, "Missing shape for tensor%d", tensor - 1);
    return it->second;
  }

  void lower_instr(const Instr& ins) {
    const std::vector<int>& a = ins.args;
    std::vector<const Val*> v;
    for (int r : a) v.push_back(&get(r));
    Val res;
    switch (ins.op) {
      case Op::Scalar: res = cscalar(ins.scalar); break;
      case Op::Index: res = cindex(ins.index); break;
      case Op::Boolean: res = cbool(ins.index != 0); break;
      case Op::Add: case Op::Sub: case Op::Mul: case Op::Div: case Op::IndexDiv: case Op::Mod: {
        const Op op = ins.op;
        if (op == Op::Add && is_zero(*v[0])) { res = *v[1]; break; }
        if ((op == Op::Add || op == Op::Sub) && is_zero(*v[1])) { res = *v[0]; break; }
        if (op == Op::Mul) {
          if (is_zero(*v[0])) { res = *v[0]; break; }
          if (is_zero(*v[1])) { res = *v[1]; break; }
          if (is_one(*v[0])) { res = *v[1]; break; }
          if (is_one(*v[1])) { res = *v[0]; break; }
        }
        if ((op == Op::Div || op == Op::IndexDiv) && (is_zero(*v[0]) || is_one(*v[1]))) { res = *v[0]; break; }
        if (op == Op::Mod && is_zero(*v[0])) { res = *v[0]; break; }
        const Ty ty = v[0]->ty;
        const bool both_known = v[0]->kind != V_DEV && v[1]->kind != V_DEV;
        if (both_known && ty == T_SCALAR && v[0]->kind == V_CONST && v[1]->kind == V_CONST) {
          const double x = v[0]->f, y = v[1]->f;
          double r = 0;
          if (op == Op::Add) r = x + y;
          else if (op == Op::Sub) r = x - y;
          else if (op == Op::Mul) r = x * y;
          else if (op == Op::Div) r = x / y;
          else fail(EGB_ERR_GENERATOR, "integer operation on scalar operands");
          res = cscalar(r);
          break;
        }
        if (both_known && ty != T_SCALAR) {
          const int64_t x = v[0]->i, y = v[1]->i;
          int64_t r = 0;
          if (op == Op::Add) r = x + y;
          else if (op == Op::Sub) r = x - y;
          else if (op == Op::Mul) r = x * y;
          else {
            if (y == 0) fail(EGB_ERR_GENERATOR, "division by zero in an index expression");
            r = (op == Op::Mod) ? x % y : x / y;
          }
          res = (v[0]->kind == V_CONST && v[1]->kind == V_CONST) ? cindex(r) : host(r);
          break;
        }
        uint8_t code;
        if (ty == T_SCALAR) {
          code = op == Op::Add ? IP_FADD : op == Op::Sub ? IP_FSUB : op == Op::Mul ? IP_FMUL : IP_FDIV;
          if (op == Op::IndexDiv || op == Op::Mod) fail(EGB_ERR_GENERATOR, "integer operation on scalar operands");
        } else {
          code = op == Op::Add ? IP_IADD : op == Op::Sub ? IP_ISUB : op == Op::Mul ? IP_IMUL
                 : op == Op::Mod ? IP_IMOD : IP_IDIV;
        }
        res = emit(code, ty, {v[0], v[1]});
        break;
      }
      case Op::Eq: case Op::Lt: case Op::Le: {
        const Ty ty = v[0]->ty;
        if (ins.op == Op::Eq && v[0]->kind == V_DEV && v[1]->kind == V_DEV && v[0]->slot == v[1]->slot) {
          res = cbool(true);
          break;
        }
        if (v[0]->kind == V_CONST && v[1]->kind == V_CONST) {
          bool r;
          if (ty == T_SCALAR) r = ins.op == Op::Eq ? v[0]->f == v[1]->f : ins.op == Op::Lt ? v[0]->f < v[1]->f : v[0]->f <= v[1]->f;
          else r = ins.op == Op::Eq ? v[0]->i == v[1]->i : ins.op == Op::Lt ? v[0]->i < v[1]->i : v[0]->i <= v[1]->i;
          res = cbool(r);
          break;
        }
        uint8_t code;
        if (ty == T_SCALAR) code = ins.op == Op::Eq ? IP_FEQ : ins.op == Op::Lt ? IP_FLT : IP_FLE;
        else if (ty == T_BOOL) code = IP_BEQ;
        else code = ins.op == Op::Eq ? IP_IEQ : ins.op == Op::Lt ? IP_ILT : IP_ILE;
        res = emit(code, T_BOOL, {v[0], v[1]});
        break;
      }
      case Op::And: case Op::Or: {
        if (v[0]->kind == V_CONST && v[1]->kind == V_CONST) {
          res = cbool(ins.op == Op::And ? (v[0]->i && v[1]->i) : (v[0]->i || v[1]->i));
          break;
        }
        res = emit(ins.op == Op::And ? IP_AND : IP_OR, T_BOOL, {v[0], v[1]});
        break;
      }
      case Op::Select: {
        if (v[0]->kind == V_CONST) { res = v[0]->i ? *v[1] : *v[2]; break; }
        res = emit(IP_SELECT, v[1]->ty, {v[0], v[1], v[2]});
        break;
      }
      case Op::Wrap: {
        if (v[0]->kind != V_DEV && v[1]->kind != V_DEV) {
          if (v[1]->i == 0) fail(EGB_ERR_GENERATOR, "wrap by zero");
          int64_t r = ((v[0]->i % v[1]->i) + v[1]->i) % v[1]->i;
          res = host(r);
          break;
        }
        res = emit(IP_IWRAP, T_INDEX, {v[0], v[1]});
        break;
      }
      case Op::Negate: {
        if (v[0]->kind == V_CONST) { res = v[0]->ty == T_SCALAR ? cscalar(0.0 - v[0]->f) : cindex(-v[0]->i); break; }
        if (v[0]->kind == V_HOST) { res = host(-v[0]->i); break; }
        res = emit(v[0]->ty == T_SCALAR ? IP_FNEG : IP_INEG, v[0]->ty, {v[0]});
        break;
      }
      case Op::Sin: res = emit(IP_SIN, T_SCALAR, {v[0]}); break;
      case Op::Cos: res = emit(IP_COS, T_SCALAR, {v[0]}); break;
      case Op::Exp: res = emit(IP_EXP, T_SCALAR, {v[0]}); break;
      case Op::Ln: res = emit(IP_LN, T_SCALAR, {v[0]}); break;
      case Op::Sqrt: res = emit(IP_SQRT, T_SCALAR, {v[0]}); break;
      case Op::Pow: res = emit(IP_POW, T_SCALAR, {v[0], v[1]}); break;
      // The reference has no CPU lowering for these three (llvmgen.nim:501-502); its OpenCL path has
      // log10/log2 (clgen.nim:45-47). The device supports all of them.
      case Op::Log10: res = emit(IP_LOG10, T_SCALAR, {v[0]}); break;
      case Op::Log2: res = emit(IP_LOG2, T_SCALAR, {v[0]}); break;
      case Op::Log: res = emit(IP_LOGB, T_SCALAR, {v[0], v[1]}); break;
      case Op::ToScalar:
        if (v[0]->kind != V_DEV) {
          // host-known index (shape / len / literal): (float)i is what the reference computes at run
          // time with sitofp; materialise it as a scalar literal (not foldable further, like there)
          const float f = (float)v[0]->i;
          uint32_t b;
          memcpy(&b, &f, 4);
          res.kind = V_DEV;
          res.ty = T_SCALAR;
          res.slot = literal(T_SCALAR, b);
          res.deps = 0;
          break;
        }
        res = emit(IP_TOSCALAR, T_SCALAR, {v[0]});
        break;
      case Op::ToIndex: res = emit(IP_TOINDEX, T_INDEX, {v[0]}); break;
      case Op::Shape: {
        const auto& shape = shape_of(ins.tensor);
        const int64_t rank = (int64_t)shape.size();
        const int64_t d = ins.dim < 0 ? rank + ins.dim : ins.dim;
        if (d < 0 || d >= rank) fail(EGB_ERR_SHAPE, "tensor%d has no dimension %d", ins.tensor - 1, ins.dim);
        res = host(shape[d]);
        break;
      }
      case Op::Len: {
        int64_t n = 1;
        for (auto s : shape_of(ins.tensor)) n *= s;
        res = host(n);
        break;
      }
      case Op::ShapeLen: res = host((int64_t)shape_of(ins.tensor).size()); break;
      case Op::Epoch: uses_epoch = true; res = host(epoch); break;
      case Op::Array: {
        res.kind = V_CONST;
        res.ty = T_ARRAY;
        res.items = a;
        break;
      }
      case Op::ArrayLen: res = cindex((int64_t)v[0]->items.size()); break;
      case Op::ArrayRead: {
        const Val& arr = *v[0];
        if (arr.ty != T_ARRAY) fail(EGB_ERR_GENERATOR, "ArrayRead of a non-array value");
        const bool nested = !arr.items.empty() && arr.row_len == 0 && get(arr.items[0]).ty == T_ARRAY;
        if (nested && v[1]->kind == V_DEV) {
          // dynamic row of an array of arrays: remember the row index, flatten the rows
          res.kind = V_CONST;
          res.ty = T_ARRAY;
          res.row_slot = v[1]->slot;
          res.deps = v[1]->deps;
          for (int r : arr.items) {
            const Val& row = get(r);
            if (row.ty != T_ARRAY || row.row_len != 0) fail(EGB_ERR_GENERATOR, "unsupported nested array literal");
            if (res.row_len == 0) res.row_len = (int)row.items.size();
            if ((int)row.items.size() != res.row_len) fail(EGB_ERR_GENERATOR, "ragged nested array literal");
            res.items.insert(res.items.end(), row.items.begin(), row.items.end());
          }
          break;
        }
        if (arr.row_len > 0) {
          // element of a dynamically selected row: flat index = row * row_len + column
          Val row;
          row.kind = V_DEV; row.ty = T_INDEX; row.slot = arr.row_slot; row.deps = arr.deps;
          const Val len = cindex(arr.row_len);
          const Val scaled = emit(IP_IMUL, T_INDEX, {&row, &len});
          const Val flat = emit(IP_IADD, T_INDEX, {&scaled, v[1]});
          const int base = array_table_len;
          if (base + (int)arr.items.size() > 64) fail(EGB_ERR_GENERATOR, "array literals too large");
          uint32_t deps = flat.deps;
          for (int r : arr.items) {
            ip.array_table[array_table_len++] = (uint8_t)slot_of(get(r));
            deps |= get(r).deps;
          }
          res = emit(IP_ARRAY_READ, T_SCALAR, {&flat}, base);
          res.deps = deps;
          break;
        }
        if (v[1]->kind != V_DEV) {
          if (v[1]->i < 0 || v[1]->i >= (int64_t)arr.items.size()) fail(EGB_ERR_GENERATOR, "array index out of range");
          res = get(arr.items[v[1]->i]);
          break;
        }
        const int base = array_table_len;
        if (base + (int)arr.items.size() > 64) fail(EGB_ERR_GENERATOR, "array literals too large");
        uint32_t deps = v[1]->deps;
        for (int r : arr.items) {
          ip.array_table[array_table_len++] = (uint8_t)slot_of(get(r));
          deps |= get(r).deps;
        }
        res = emit(IP_ARRAY_READ, T_SCALAR, {v[1]}, base);
        res.deps = deps;
        break;
      }
      default: fail(EGB_ERR_GENERATOR, "Unable to generate device code for Instr%s", op_name(ins.op));
    }
    env[ins.res] = res;
  }

  void lower_setup(const LinearIndex& li) {
    for (auto& ins : li.setup)
      if (!env.count(ins.res)) lower_instr(ins);
  }

  int64_t host_value(const LinearIndex& li, const char* what) {
    lower_setup(li);
    int64_t v = li.constant;
    for (auto& kv : li.factors) {
      const Val& x = get(kv.first);
      if (x.kind == V_DEV) fail(EGB_ERR_GENERATOR, "%s depends on a loop iterator", what);
      v += kv.second * x.i;
    }
    return v;
  }

  // Flat element index of a tensor access as offset + sum(coef * slot).
  void flatten(const TensorOp& op, IpTensorOp& out, uint32_t* deps_out) {
    std::map<int, int64_t> terms;  // slot -> coefficient
    int64_t offset = 0;
    uint32_t deps = 0;
    const auto& shape = shape_of(op.tensor);
    if (!op.is_raw && op.dims.size() != shape.size())
      fail(EGB_ERR_SHAPE, "tensor%d is accessed with %zu indices but has rank %zu", op.tensor - 1, op.dims.size(),
           shape.size());
    int64_t stride = 1;
    for (size_t dd = op.dims.size(); dd-- > 0;) {
      const LinearIndex& li = op.dims[dd];
      lower_setup(li);
      offset += li.constant * stride;
      for (auto& kv : li.factors) {
        const Val& x = get(kv.first);
        if (x.kind == V_DEV) {
          terms[x.slot] += kv.second * stride;
          deps |= x.deps;
        } else {
          offset += kv.second * x.i * stride;
        }
      }
      if (!op.is_raw) stride *= shape[dd];
    }
    out.offset = offset;
    out.nterms = 0;
    for (auto& kv : terms) {
      if (kv.second == 0) continue;
      if (out.nterms >= IP_MAX_TERMS) fail(EGB_ERR_GENERATOR, "tensor index has too many terms");
      out.slot[out.nterms] = (uint8_t)kv.first;
      out.coef[out.nterms] = kv.second;
      out.nterms++;
    }
    if (deps_out) *deps_out = deps;
  }
};

bool full_range(Lowerer& lw, const Loop& loop, int64_t extent) {
  return lw.host_value(loop.start, "loop bound") == 0 && loop.step == 1 &&
         lw.host_value(loop.stop, "loop bound") == extent;
}

}  // namespace

bool covers_whole_tensor(const Kernel& k, const ShapeTable& shapes) {
  // every element of the written tensor is produced exactly once: the write dims are distinct
  // iterators, each running over the full extent of its dimension.
  auto it = shapes.find(k.write.tensor);
  if (it == shapes.end()) return false;
  const auto& shape = it->second;
  Lowerer lw(k, shapes, 0);
  std::vector<IpInstr> scratch;
  lw.sink = &scratch;
  for (size_t i = 0; i < k.loops.size(); ++i) {
    Val v;
    v.kind = V_DEV;
    v.ty = T_INDEX;
    v.slot = lw.new_slot();
    lw.env[k.loops[i].iter] = v;
  }
  std::set<int> seen;
  int64_t len = 1;
  for (auto s : shape) len *= s;
  if (!k.write.is_raw && k.write.dims.size() != shape.size()) return false;
  for (size_t d = 0; d < k.write.dims.size(); ++d) {
    const int reg = k.write.dims[d].only_register();
    if (!reg || seen.count(reg)) return false;
    seen.insert(reg);
    const Loop* loop = nullptr;
    for (auto& l : k.loops)
      if (l.iter == reg) loop = &l;
    if (!loop || !loop->has_bounds) return false;
    const int64_t extent = k.write.is_raw ? len : shape[d];
    try {
      if (!full_range(lw, *loop, extent)) return false;
    } catch (const Error&) {
      return false;
    }
  }
  if (k.write.dims.empty()) return false;
  return true;
}

bool kernel_loops_full(const Kernel& k, const ShapeTable& shapes) {
  Lowerer lw(k, shapes, 0);
  std::vector<IpInstr> scratch;
  lw.sink = &scratch;
  for (auto& l : k.loops) {
    Val v;
    v.kind = V_DEV;
    v.ty = T_INDEX;
    v.slot = lw.new_slot();
    lw.env[l.iter] = v;
  }
  try {
    for (auto& l : k.loops) {
      if (!l.has_bounds) return false;
      int64_t extent = -1;
      auto visit = [&](const TensorOp& op) {
        if (extent >= 0) return;
        auto sh = shapes.find(op.tensor);
        if (sh == shapes.end()) return;
        for (size_t d = 0; d < op.dims.size(); ++d) {
          if (op.dims[d].only_register() != l.iter) continue;
          if (op.is_raw) {
            extent = 1;
            for (auto s : sh->second) extent *= s;
          } else if (d < sh->second.size()) {
            extent = sh->second[d];
          }
          return;
        }
      };
      visit(k.write);
      for (auto& r : k.reads) visit(r);
      if (extent < 0 || !full_range(lw, l, extent)) return false;
    }
  } catch (const Error&) {
    return false;
  }
  return true;
}

// Loop-nest simplification for the large generic kernels (nothing but the index arithmetic sees the
// iterators): every loop becomes 0 .. count with unit step (start and step folded into the accesses),
// loops of one iteration disappear, and adjacent loops of the same kind whose accesses are contiguous
// across them (coefficient of the outer == coefficient of the inner x count of the inner, for every
// access) merge into one. A [N,H,W,F] elementwise kernel becomes one flat loop; sum over (n, y, x) of
// t[n,y,x,f] becomes one reduction loop of stride F.
static void simplify_loops(IpProgram& ip) {
  auto arity = [](uint8_t op) {
    switch (op) {
      case IP_FNEG: case IP_SIN: case IP_COS: case IP_EXP: case IP_LN: case IP_SQRT: case IP_LOG10: case IP_LOG2:
      case IP_INEG: case IP_TOSCALAR: case IP_TOINDEX: case IP_ARRAY_READ: return 1;
      case IP_SELECT: return 3;
      default: return 2;
    }
  };
  bool as_value[256] = {false};
  for (int i = 0; i < ip.ninstrs; ++i) {
    const IpInstr& in = ip.instrs[i];
    const int n = arity(in.op);
    as_value[in.a] = true;
    if (n >= 2) as_value[in.b] = true;
    if (n >= 3) as_value[in.c] = true;
  }
  for (int i = 0; i < ip.ninstrs; ++i)
    if (ip.instrs[i].op == IP_ARRAY_READ) return;   // array elements are referenced through a slot table
  std::vector<IpTensorOp*> ops;
  for (int r = 0; r < ip.nreads; ++r) ops.push_back(&ip.reads[r]);
  ops.push_back(&ip.write);
  auto loop_of_slot = [&](int slot) {
    for (int l = 0; l < ip.nloops; ++l)
      if (ip.loops[l].slot == slot) return l;
    return -1;
  };
  for (auto* op : ops)
    for (int t = 0; t < op->nterms; ++t)
      if (loop_of_slot(op->slot[t]) < 0) return;   // an index term that is not a loop iterator
  for (int l = 0; l < ip.nloops; ++l)
    if (as_value[ip.loops[l].slot]) return;         // an iterator used as data
  // canonical terms: one per loop, zero coefficients dropped, start/step folded
  auto coef_of = [&](const IpTensorOp& op, int slot) {
    int64_t c = 0;
    for (int t = 0; t < op.nterms; ++t)
      if (op.slot[t] == slot) c += op.coef[t];
    return c;
  };
  for (auto* op : ops) {
    int64_t coef[IP_MAX_LOOPS];
    for (int l = 0; l < ip.nloops; ++l) {
      const int64_t c = coef_of(*op, ip.loops[l].slot);
      op->offset += c * ip.loops[l].start;
      coef[l] = c * ip.loops[l].step;
    }
    op->nterms = 0;
    for (int l = 0; l < ip.nloops; ++l)
      if (coef[l] != 0 && ip.loops[l].count > 1) {
        op->coef[op->nterms] = coef[l];
        op->slot[op->nterms] = ip.loops[l].slot;
        ++op->nterms;
      }
  }
  for (int l = 0; l < ip.nloops; ++l) {
    ip.loops[l].start = 0;
    ip.loops[l].step = 1;
  }
  auto erase_loop = [&](int l) {
    for (int j = l; j + 1 < ip.nloops; ++j) ip.loops[j] = ip.loops[j + 1];
    if (l < ip.npar) --ip.npar;
    --ip.nloops;
  };
  for (int l = ip.nloops - 1; l >= 0; --l)
    if (ip.loops[l].count == 1 && ip.nloops > 1) erase_loop(l);
  bool changed = true;
  while (changed) {
    changed = false;
    for (int l = 0; l + 1 < ip.nloops && !changed; ++l) {
      const bool same_kind = (l + 1 < ip.npar) || (l >= ip.npar);
      if (!same_kind) continue;
      const int so = ip.loops[l].slot, si = ip.loops[l + 1].slot;
      bool ok = true;
      for (auto* op : ops) ok = ok && coef_of(*op, so) == coef_of(*op, si) * ip.loops[l + 1].count;
      if (!ok) continue;
      for (auto* op : ops) {
        int w = 0;
        for (int t = 0; t < op->nterms; ++t)
          if (op->slot[t] != so) {
            op->coef[w] = op->coef[t];
            op->slot[w] = op->slot[t];
            ++w;
          }
        op->nterms = (uint8_t)w;
      }
      ip.loops[l + 1].count *= ip.loops[l].count;
      erase_loop(l);
      changed = true;
    }
  }
}

Lowered lower_kernel(const Kernel& k, const ShapeTable& shapes, const std::map<int, void*>& ptrs, int64_t epoch,
                     bool strict, bool overwrite, int sm_count) {
  Lowerer lw(k, shapes, epoch);
  IpProgram& ip = lw.ip;
  const int nloops = (int)k.loops.size();
  if (nloops > IP_MAX_LOOPS) fail(EGB_ERR_GENERATOR, "kernel has more than %d loops", IP_MAX_LOOPS);
  if ((int)k.reads.size() > IP_MAX_OPS) fail(EGB_ERR_GENERATOR, "kernel has more than %d tensor reads", IP_MAX_OPS);

  // loop iterators live in slots; bounds are host values
  lw.sink = &lw.index_instrs;
  std::vector<int> iter_slot(nloops);
  for (int i = 0; i < nloops; ++i) {
    Val v;
    v.kind = V_DEV;
    v.ty = T_INDEX;
    v.slot = lw.new_slot();
    v.deps = 1u << i;
    iter_slot[i] = v.slot;
    lw.env[k.loops[i].iter] = v;
  }
  std::vector<int64_t> start(nloops), count(nloops);
  for (int i = 0; i < nloops; ++i) {
    const Loop& l = k.loops[i];
    if (!l.has_bounds) fail(EGB_ERR_GENERATOR, "loop without bounds");
    if (l.step <= 0) fail(EGB_ERR_GENERATOR, "loop step must be positive");
    start[i] = lw.host_value(l.start, "loop bound");
    const int64_t stop = lw.host_value(l.stop, "loop bound");
    count[i] = stop > start[i] ? (stop - start[i] + l.step - 1) / l.step : 0;
  }

  // index arithmetic of the tensor accesses (may emit per-point index instructions)
  uint32_t write_deps = 0;
  for (size_t r = 0; r < k.reads.size(); ++r) {
    IpTensorOp& op = ip.reads[r];
    lw.flatten(k.reads[r], op, nullptr);
    auto p = ptrs.find(k.reads[r].tensor);
    if (p == ptrs.end() || !p->second) fail(EGB_ERR_RUNTIME, "tensor%d has no storage", k.reads[r].tensor - 1);
    op.base = (uint64_t)p->second;
    Val v;
    v.kind = V_DEV;
    v.ty = T_SCALAR;
    v.slot = lw.new_slot();
    v.deps = 0xffffffffu;
    op.dst = (uint8_t)v.slot;
    lw.env[k.reads[r].data] = v;
  }
  ip.nreads = (uint8_t)k.reads.size();
  lw.flatten(k.write, ip.write, &write_deps);
  {
    auto p = ptrs.find(k.write.tensor);
    if (p == ptrs.end() || !p->second) fail(EGB_ERR_RUNTIME, "tensor%d has no storage", k.write.tensor - 1);
    ip.write.base = (uint64_t)p->second;
  }

  // expression
  lw.sink = &lw.body_instrs;
  for (auto& ins : k.instrs) lw.lower_instr(ins);
  const Val& value = lw.get(k.write.data);
  if (value.ty != T_SCALAR) fail(EGB_ERR_GENERATOR, "kernel writes a non-scalar value");
  if (value.kind == V_DEV && lw.index_instrs.size() > 0) {
    // (index instructions emitted while lowering the body stay in the body; nothing to do)
  }
  ip.write.dst = (uint8_t)lw.slot_of(value);

  if (lw.index_instrs.size() > 32) fail(EGB_ERR_GENERATOR, "index expressions too complex for the device program");
  if (lw.body_instrs.size() > (size_t)IP_MAX_INSTRS) fail(EGB_ERR_GENERATOR, "expression too large for the device program");
  ip.nindex_instrs = (uint8_t)lw.index_instrs.size();
  for (size_t i = 0; i < lw.index_instrs.size(); ++i) ip.index_instrs[i] = lw.index_instrs[i];
  ip.ninstrs = (uint8_t)lw.body_instrs.size();
  for (size_t i = 0; i < lw.body_instrs.size(); ++i) ip.instrs[i] = lw.body_instrs[i];

  // partition loops: independent ones are thread-mapped, the rest are walked in the given order
  std::vector<int> par, red;
  for (int i = 0; i < nloops; ++i) (k.loops[i].mode >= 1 ? par : red).push_back(i);
  auto write_coef = [&](int loop) {
    int64_t c = 0;
    for (int t = 0; t < ip.write.nterms; ++t)
      if (ip.write.slot[t] == iter_slot[loop]) c = ip.write.coef[t] < 0 ? -ip.write.coef[t] : ip.write.coef[t];
    return c;
  };
  std::stable_sort(par.begin(), par.end(), [&](int x, int y) { return write_coef(x) > write_coef(y); });
  uint32_t red_mask = 0;
  for (int i : red) red_mask |= 1u << i;
  ip.scatter = (write_deps & red_mask) != 0;
  ip.npar = (uint8_t)par.size();
  ip.nloops = (uint8_t)nloops;
  ip.npoints = 1;
  ip.nred = 1;
  int pos = 0;
  for (int i : par) {
    ip.loops[pos].start = start[i];
    ip.loops[pos].step = k.loops[i].step;
    ip.loops[pos].count = count[i];
    ip.loops[pos].slot = (uint8_t)iter_slot[i];
    ip.npoints *= count[i];
    ++pos;
  }
  for (int i : red) {
    ip.loops[pos].start = start[i];
    ip.loops[pos].step = k.loops[i].step;
    ip.loops[pos].count = count[i];
    ip.loops[pos].slot = (uint8_t)iter_slot[i];
    ip.nred *= count[i];
    ++pos;
  }
  ip.accumulate = (overwrite && !ip.scatter) ? 0 : 1;
  // small kernels keep their loop structure (row chains look for the batch loop)
  if (!strict && ip.nindex_instrs == 0 && ip.npoints * ip.nred > (1 << 16)) simplify_loops(ip);
  bool iter_slots_small = true;   // the 4-wide kernels index iterator values by slot
  for (int l = 0; l < ip.nloops; ++l) iter_slots_small = iter_slots_small && ip.loops[l].slot < IP_MAX_LOOPS;

  // 4-wide fast path: no reduction, the innermost independent loop has unit stride, every access either
  // streams with it (coefficient 1) or does not depend on it (broadcast along the row), and the
  // expression only uses fp32 / boolean operations. Outer iterators only enter the address arithmetic.
  {
    bool ok = ip.npar >= 1 && ip.npar == ip.nloops && !ip.scatter && ip.nindex_instrs == 0 && lw.nslots <= 64 &&
              iter_slots_small && ip.loops[ip.npar - 1].step == 1 && ip.npoints > 0;
    const int inner_slot = ok ? ip.loops[ip.npar - 1].slot : -1;
    const int64_t inner_start = ok ? ip.loops[ip.npar - 1].start : 0;
    auto classify = [&](IpTensorOp& op, bool must_stream) {
      int64_t inner_coef = 0;
      bool outer_mult4 = true;
      for (int t = 0; t < op.nterms; ++t) {
        if (op.slot[t] == inner_slot) {
          inner_coef = op.coef[t];
        } else {
          bool is_loop = false;
          for (int l = 0; l < ip.npar; ++l) is_loop = is_loop || ip.loops[l].slot == op.slot[t];
          if (!is_loop) ok = false;
          outer_mult4 = outer_mult4 && (op.coef[t] % 4 == 0);
        }
      }
      if (inner_coef == 1) op.streaming = 1;
      else if (inner_coef == 0 && !must_stream) op.streaming = 0;
      else ok = false;
      const int64_t first = op.offset + (op.streaming ? inner_start : 0);
      op.aligned16 = outer_mult4 && (((op.base >> 2) + (uint64_t)first) & 3) == 0 && (op.base & 3) == 0;
    };
    if (ok) {
      classify(ip.write, true);
      for (int r = 0; r < ip.nreads; ++r) classify(ip.reads[r], false);
      for (int i = 0; i < ip.ninstrs && ok; ++i) {
        const uint8_t op = ip.instrs[i].op;
        const bool fp = (op >= IP_FADD && op <= IP_LOGB) || op == IP_FEQ || op == IP_FLT || op == IP_FLE ||
                        op == IP_BEQ || op == IP_AND || op == IP_OR || op == IP_SELECT;
        ok = ok && fp;
      }
      // index-typed literals cannot be splatted into fp32 lanes; they are only present if unused
      for (int i = 0; i < ip.nlits && ok; ++i) ok = ok && (ip.lits[i] >> 32) == 0;
    }
    ip.vec4 = ok ? 1 : 0;
  }
  // streaming reduction (vec4 == 2): exactly one reduction loop with unit stride and a long range, every
  // read streams with it or ignores it, fp32 / boolean expression; outer iterators only address
  if (!ip.vec4) {
    bool ok = ip.nloops == ip.npar + 1 && !ip.scatter && ip.nindex_instrs == 0 && lw.nslots <= 64 &&
              iter_slots_small && ip.loops[ip.npar].step == 1 && ip.loops[ip.npar].count >= 4096 && ip.npoints > 0;
    const int red_slot = ok ? ip.loops[ip.npar].slot : -1;
    const int64_t red_start = ok ? ip.loops[ip.npar].start : 0;
    auto classify = [&](IpTensorOp& op, bool is_write) {
      int64_t c_red = 0;
      bool outer_mult4 = true;
      for (int t = 0; t < op.nterms; ++t) {
        if (op.slot[t] == red_slot) {
          c_red = op.coef[t];
        } else {
          bool is_loop = false;
          for (int l = 0; l < ip.npar; ++l) is_loop = is_loop || ip.loops[l].slot == op.slot[t];
          if (!is_loop) ok = false;
          outer_mult4 = outer_mult4 && (op.coef[t] % 4 == 0);
        }
      }
      if (is_write) {
        if (c_red != 0) ok = false;
        return;
      }
      if (c_red == 1) op.streaming = 1;
      else if (c_red == 0) op.streaming = 0;
      else ok = false;
      const int64_t first = op.offset + (op.streaming ? red_start : 0);
      op.aligned16 = outer_mult4 && (((op.base >> 2) + (uint64_t)first) & 3) == 0 && (op.base & 3) == 0;
    };
    if (ok) {
      classify(ip.write, true);
      bool any_stream = false;
      for (int r = 0; r < ip.nreads; ++r) {
        classify(ip.reads[r], false);
        any_stream = any_stream || ip.reads[r].streaming;
      }
      ok = ok && any_stream;
      for (int i = 0; i < ip.ninstrs && ok; ++i) {
        const uint8_t op = ip.instrs[i].op;
        const bool fp = (op >= IP_FADD && op <= IP_LOGB) || op == IP_FEQ || op == IP_FLT || op == IP_FLE ||
                        op == IP_BEQ || op == IP_AND || op == IP_OR || op == IP_SELECT;
        ok = ok && fp;
      }
      for (int i = 0; i < ip.nlits && ok; ++i) ok = ok && (ip.lits[i] >> 32) == 0;
    }
    if (ok) ip.vec4 = 2;
  }
  // reduction with streaming output points (vec4 == 3): one reduction loop with arbitrary strides, the
  // innermost independent loop has unit stride and the write and every read either stream with it or
  // ignore it
  if (!ip.vec4) {
    bool ok = ip.npar >= 1 && ip.nloops == ip.npar + 1 && !ip.scatter && ip.nindex_instrs == 0 && lw.nslots <= 64 &&
              iter_slots_small && ip.loops[ip.npar - 1].step == 1 && ip.loops[ip.npar - 1].count >= 4 &&
              ip.nred >= 64 && ip.npoints * ip.nred > (1 << 16);
    const int inner_slot = ok ? ip.loops[ip.npar - 1].slot : -1;
    const int64_t inner_start = ok ? ip.loops[ip.npar - 1].start : 0;
    auto classify = [&](IpTensorOp& op, bool is_write) {
      int64_t inner_coef = 0;
      bool other_mult4 = true;
      for (int t = 0; t < op.nterms; ++t) {
        if (op.slot[t] == inner_slot) {
          inner_coef = op.coef[t];
        } else {
          bool is_loop = false;
          for (int l = 0; l < ip.nloops; ++l) is_loop = is_loop || ip.loops[l].slot == op.slot[t];
          if (!is_loop) ok = false;
          int64_t step = 1, start = 0;
          for (int l = 0; l < ip.nloops; ++l)
            if (ip.loops[l].slot == op.slot[t]) { step = ip.loops[l].step; start = ip.loops[l].start; }
          other_mult4 = other_mult4 && ((op.coef[t] * step) % 4 == 0) && ((op.coef[t] * start) % 4 == 0);
        }
      }
      if (inner_coef == 1) op.streaming = 1;
      else if (inner_coef == 0 && !is_write) op.streaming = 0;
      else ok = false;
      const int64_t first = op.offset + (op.streaming ? inner_start : 0);
      op.aligned16 = other_mult4 && (((op.base >> 2) + (uint64_t)first) & 3) == 0 && (op.base & 3) == 0;
    };
    if (ok) {
      classify(ip.write, true);
      for (int r = 0; r < ip.nreads; ++r) classify(ip.reads[r], false);
      for (int i = 0; i < ip.ninstrs && ok; ++i) {
        const uint8_t op = ip.instrs[i].op;
        const bool fp = (op >= IP_FADD && op <= IP_LOGB) || op == IP_FEQ || op == IP_FLT || op == IP_FLE ||
                        op == IP_BEQ || op == IP_AND || op == IP_OR || op == IP_SELECT;
        ok = ok && fp;
      }
      for (int i = 0; i < ip.nlits && ok; ++i) ok = ok && (ip.lits[i] >> 32) == 0;
    }
    if (ok) ip.vec4 = 3;
  }
  ip.nslots = (uint8_t)std::min(lw.nslots, 255);

  Lowered out;
  out.uses_epoch = lw.uses_epoch;
  out.nslots = lw.nslots;
  // thread layout
  int pb = 256, rb = 1, points_fast = 1;
  if (!strict && !ip.scatter && ip.nred > 1) {
    const int64_t want_threads = (int64_t)sm_count * 256 * 2;
    int64_t r = 1;
    while (r < 256 && ip.npoints * r < want_threads && r * 2 <= ip.nred) r *= 2;
    rb = (int)r;
    pb = 256 / rb;
    // which dimension is contiguous in memory? compare the smallest read stride of the fastest
    // point loop with that of the fastest reduction loop
    auto min_read_coef = [&](int slot) {
      int64_t best = INT64_MAX;
      for (int q = 0; q < ip.nreads; ++q)
        for (int t = 0; t < ip.reads[q].nterms; ++t)
          if (ip.reads[q].slot[t] == slot) {
            int64_t c = ip.reads[q].coef[t] < 0 ? -ip.reads[q].coef[t] : ip.reads[q].coef[t];
            best = std::min(best, c);
          }
      return best;
    };
    if (ip.npar == 0) {
      points_fast = 0;
    } else {
      const int64_t cp = min_read_coef(ip.loops[ip.npar - 1].slot);
      const int64_t cr = ip.nloops == ip.npar ? INT64_MAX : min_read_coef(ip.loops[ip.nloops - 1].slot);
      points_fast = cp <= cr ? 1 : 0;
    }
    // when neighbouring lanes walk the output points (contiguous reads along the points), keep at
    // least a full warp of points per block: wider parallelism then comes from splitting the reduction
    // over blocks (interp_reduction_splits), not from shrinking the coalesced dimension
    if (points_fast && pb < 32 && ip.npoints >= 32) {
      pb = 32;
      rb = 8;
    }
  }
  out.pb = pb;
  out.rb = rb;
  out.points_fast = points_fast;
  out.ip = ip;
  return out;
}

// ------------------------------------------------------------------ contraction pattern

bool match_gemm(const Kernel& k, const ShapeTable& shapes, GemmPattern& g) {
  if (k.loops.size() != 3 || k.reads.size() != 2 || k.instrs.size() != 1) return false;
  const Instr& mul = k.instrs[0];
  if (mul.op != Op::Mul || mul.args.size() != 2 || k.write.data != mul.res) return false;
  const TensorOp& r0 = k.reads[0];
  const TensorOp& r1 = k.reads[1];
  if (!((mul.args[0] == r0.data && mul.args[1] == r1.data) || (mul.args[0] == r1.data && mul.args[1] == r0.data)))
    return false;
  if (k.write.is_raw || r0.is_raw || r1.is_raw) return false;
  if (k.write.dims.size() != 2 || r0.dims.size() != 2 || r1.dims.size() != 2) return false;
  const int m = k.write.dims[0].only_register(), n = k.write.dims[1].only_register();
  if (!m || !n || m == n) return false;
  int kk = 0;
  for (auto& l : k.loops)
    if (l.iter != m && l.iter != n) kk = l.iter;
  if (!kk) return false;
  auto regs = [](const TensorOp& op, int& a, int& b) {
    a = op.dims[0].only_register();
    b = op.dims[1].only_register();
    return a && b;
  };
  int a0, a1, b0, b1;
  if (!regs(r0, a0, a1) || !regs(r1, b0, b1)) return false;
  const TensorOp *A = nullptr, *B = nullptr;
  auto has = [](int x, int y, int r) { return x == r || y == r; };
  if (has(a0, a1, m) && has(a0, a1, kk) && has(b0, b1, n) && has(b0, b1, kk)) {
    A = &r0;
    B = &r1;
  } else if (has(b0, b1, m) && has(b0, b1, kk) && has(a0, a1, n) && has(a0, a1, kk)) {
    A = &r1;
    B = &r0;
  } else {
    return false;
  }
  if (A->tensor == k.write.tensor || B->tensor == k.write.tensor) return false;
  // full-range loops over the operand extents
  auto sa = shapes.find(A->tensor), sb = shapes.find(B->tensor), sc = shapes.find(k.write.tensor);
  if (sa == shapes.end() || sb == shapes.end() || sc == shapes.end()) return false;
  if (sa->second.size() != 2 || sb->second.size() != 2 || sc->second.size() != 2) return false;
  g.trans_a = A->dims[0].only_register() == kk;  // stored [K, M]
  g.trans_b = B->dims[0].only_register() == n;   // stored [N, K]
  g.M = sc->second[0];
  g.N = sc->second[1];
  g.K = g.trans_a ? sa->second[0] : sa->second[1];
  const int64_t am = g.trans_a ? sa->second[1] : sa->second[0];
  const int64_t bk = g.trans_b ? sb->second[1] : sb->second[0];
  const int64_t bn = g.trans_b ? sb->second[0] : sb->second[1];
  if (am != g.M || bk != g.K || bn != g.N) return false;
  Lowerer lw(k, shapes, 0);
  std::vector<IpInstr> scratch;
  lw.sink = &scratch;
  for (auto& l : k.loops) {
    Val v;
    v.kind = V_DEV;
    v.ty = T_INDEX;
    v.slot = lw.new_slot();
    lw.env[l.iter] = v;
  }
  try {
    for (auto& l : k.loops) {
      if (!l.has_bounds) return false;
      const int64_t extent = l.iter == m ? g.M : l.iter == n ? g.N : g.K;
      if (!full_range(lw, l, extent)) return false;
    }
  } catch (const Error&) {
    return false;
  }
  g.a_tensor = A->tensor;
  g.b_tensor = B->tensor;
  g.c_tensor = k.write.tensor;
  g.lda = sa->second[1];
  g.ldb = sb->second[1];
  g.ldc = sc->second[1];
  return true;
}

// ------------------------------------------------------------------ conv2 pattern

bool match_conv2(const Kernel& k, const ShapeTable& shapes, ConvPattern& c) {
  if (k.loops.size() != 7 || k.reads.size() != 2 || k.instrs.size() != 1) return false;
  const Instr& mul = k.instrs[0];
  if (mul.op != Op::Mul || k.write.data != mul.res) return false;
  if (!((mul.args[0] == k.reads[0].data && mul.args[1] == k.reads[1].data) ||
        (mul.args[0] == k.reads[1].data && mul.args[1] == k.reads[0].data)))
    return false;
  const TensorOp* ops[3] = {&k.write, &k.reads[0], &k.reads[1]};
  const TensorOp* img = nullptr;
  std::vector<const TensorOp*> plain;
  for (auto* op : ops) {
    if (op->is_raw || op->dims.size() != 4) return false;
    bool all_single = true;
    for (auto& d : op->dims) all_single = all_single && d.only_register() != 0;
    if (all_single) {
      plain.push_back(op);
    } else {
      if (img) return false;
      img = op;
    }
  }
  if (!img || plain.size() != 2) return false;
  // image role: [n, y+dy, x+dx, c]
  auto pair_of = [](const LinearIndex& li, int& a, int& b) {
    if (li.constant != 0 || !li.setup.empty() || li.factors.size() != 2) return false;
    auto it = li.factors.begin();
    if (it->second != 1) return false;
    a = it->first;
    ++it;
    if (it->second != 1) return false;
    b = it->first;
    return true;
  };
  const int n = img->dims[0].only_register(), ch = img->dims[3].only_register();
  int a1, b1, a2, b2;
  if (!n || !ch || !pair_of(img->dims[1], a1, b1) || !pair_of(img->dims[2], a2, b2)) return false;
  // filter role: [f, dy, dx, c]; output role: [n, y, x, f]
  const TensorOp *fil = nullptr, *out = nullptr;
  for (auto* op : plain) {
    if (op->dims[3].only_register() == ch && op->dims[0].only_register() != n) fil = op;
    else if (op->dims[0].only_register() == n) out = op;
  }
  if (!fil || !out) return false;
  const int f = fil->dims[0].only_register(), dy = fil->dims[1].only_register(), dx = fil->dims[2].only_register();
  if ((dy != a1 && dy != b1) || (dx != a2 && dx != b2)) return false;
  const int y = dy == a1 ? b1 : a1, x = dx == a2 ? b2 : a2;
  if (out->dims[1].only_register() != y || out->dims[2].only_register() != x || out->dims[3].only_register() != f)
    return false;
  std::set<int> regs = {n, ch, f, dy, dx, y, x};
  if (regs.size() != 7) return false;
  if (img->tensor == fil->tensor || img->tensor == out->tensor || fil->tensor == out->tensor) return false;
  auto si = shapes.find(img->tensor), sf = shapes.find(fil->tensor), so = shapes.find(out->tensor);
  if (si == shapes.end() || sf == shapes.end() || so == shapes.end()) return false;
  if (si->second.size() != 4 || sf->second.size() != 4 || so->second.size() != 4) return false;
  c.N = (int)si->second[0]; c.H = (int)si->second[1]; c.W = (int)si->second[2]; c.C = (int)si->second[3];
  c.F = (int)sf->second[0]; c.KH = (int)sf->second[1]; c.KW = (int)sf->second[2];
  if (sf->second[3] != c.C || so->second[0] != c.N || so->second[1] != c.H - c.KH + 1 ||
      so->second[2] != c.W - c.KW + 1 || so->second[3] != c.F)
    return false;
  if (!kernel_loops_full(k, shapes)) return false;
  c.img_tensor = img->tensor;
  c.fil_tensor = fil->tensor;
  c.out_tensor = out->tensor;
  c.kind = &k.write == out ? ConvPattern::FORWARD : &k.write == fil ? ConvPattern::D_FILTERS : ConvPattern::D_IMAGES;
  return true;
}

}  // namespace egb
