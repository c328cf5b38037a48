// Program IR utilities + parser of the serialized program text.
#include "program.hpp"

#include <stdlib.h>
#include <string.h>

#include <sstream>

#include "egb_internal.hpp"

namespace egb {

static const char* kOpNames[] = {"Index", "Scalar", "Boolean", "Add", "Sub", "Mul", "Div", "IndexDiv", "Mod",
                                 "Wrap", "Negate", "Sin", "Cos", "Exp", "Pow", "Sqrt", "Log", "Log10", "Log2",
                                 "Ln", "Eq", "Lt", "Le", "And", "Or", "Select", "ToScalar", "ToIndex", "Shape",
                                 "Len", "ShapeLen", "Array", "ArrayLen", "ArrayRead", "Epoch", "Invalid"};

const char* op_name(Op op) { return kOpNames[(int)op]; }

Op op_from_name(const std::string& s) {
  for (int i = 0; i < (int)Op::Invalid; ++i)
    if (s == kOpNames[i]) return (Op)i;
  fail(EGB_ERR_PARSER, "unknown opcode '%s'", s.c_str());
}

LinearIndex LinearIndex::scaled(int64_t b) const {
  LinearIndex r;
  if (b == 0) return r;
  r.setup = setup;
  r.constant = constant * b;
  for (auto& kv : factors) r.factors[kv.first] = kv.second * b;
  return r;
}

LinearIndex LinearIndex::plus(const LinearIndex& o) const {
  LinearIndex r = *this;
  r.constant += o.constant;
  r.setup.insert(r.setup.end(), o.setup.begin(), o.setup.end());
  for (auto& kv : o.factors) {
    auto it = r.factors.find(kv.first);
    if (it != r.factors.end()) {
      it->second += kv.second;
      if (it->second == 0) r.factors.erase(it);
    } else {
      r.factors[kv.first] = kv.second;
    }
  }
  return r;
}

int LinearIndex::only_register() const {
  if (constant == 0 && factors.size() == 1 && factors.begin()->second == 1) return factors.begin()->first;
  return 0;
}

static bool same_instr(const Instr& a, const Instr& b) {
  return a.op == b.op && a.args == b.args && a.res == b.res && a.tensor == b.tensor && a.scalar == b.scalar &&
         a.index == b.index && a.dim == b.dim;
}

bool LinearIndex::same_as(const LinearIndex& o) const {
  if (constant != o.constant || factors != o.factors || setup.size() != o.setup.size()) return false;
  for (size_t i = 0; i < setup.size(); ++i)
    if (!same_instr(setup[i], o.setup[i])) return false;
  return true;
}

std::shared_ptr<Kernel> Kernel::clone() const {
  auto k = std::make_shared<Kernel>(*this);
  if (custom_grad) {
    k->custom_grad = std::make_shared<CustomGrad>();
    k->custom_grad->tensors = custom_grad->tensors;
    k->custom_grad->subs = custom_grad->subs;
    for (auto& g : custom_grad->kernels) k->custom_grad->kernels.push_back(g->clone());
  }
  return k;
}

static void sub_instrs(std::vector<Instr>& v, const std::map<int, int>& subs) {
  for (auto& i : v)
    if (i.tensor) {
      auto it = subs.find(i.tensor);
      if (it != subs.end()) i.tensor = it->second;
    }
}

void Kernel::substitute_tensors(const std::map<int, int>& subs) {
  auto sub = [&](int& t) {
    auto it = subs.find(t);
    if (it != subs.end()) t = it->second;
  };
  for (auto& r : reads) {
    sub(r.tensor);
    for (auto& d : r.dims) sub_instrs(d.setup, subs);
  }
  sub(write.tensor);
  for (auto& d : write.dims) sub_instrs(d.setup, subs);
  sub_instrs(instrs, subs);
  for (auto& l : loops) {
    sub_instrs(l.start.setup, subs);
    sub_instrs(l.stop.setup, subs);
  }
}

Target* Program::find_target(const std::string& name) {
  for (auto& t : targets)
    if (t->name == name) return t.get();
  return nullptr;
}

// ------------------------------------------------------------------------------ parser

namespace {

struct Reader {
  std::vector<std::string> toks;
  size_t pos = 0;
  explicit Reader(const std::string& text) {
    std::istringstream ss(text);
    std::string t;
    while (ss >> t) toks.push_back(t);
  }
  const std::string& next() {
    if (pos >= toks.size()) fail(EGB_ERR_PARSER, "unexpected end of program text");
    return toks[pos++];
  }
  void expect(const char* s) {
    const std::string& t = next();
    if (t != s) fail(EGB_ERR_PARSER, "program text: expected '%s', got '%s' (token %zu)", s, t.c_str(), pos - 1);
  }
  int64_t i64() {
    const std::string& t = next();
    char* end = nullptr;
    long long v = strtoll(t.c_str(), &end, 10);
    if (*end) fail(EGB_ERR_PARSER, "program text: expected integer, got '%s'", t.c_str());
    return v;
  }
  int i32() { return (int)i64(); }
  double f64() {
    const std::string& t = next();
    char* end = nullptr;
    double v = strtod(t.c_str(), &end);  // accepts decimal, hex floats, inf, nan
    if (*end) fail(EGB_ERR_PARSER, "program text: expected number, got '%s'", t.c_str());
    return v;
  }
  std::string str() {  // percent-decoded, "-" = empty
    const std::string& t = next();
    if (t == "-") return "";
    std::string out;
    for (size_t i = 0; i < t.size(); ++i) {
      if (t[i] == '%' && i + 2 <= t.size() - 1) {
        out.push_back((char)strtol(t.substr(i + 1, 2).c_str(), nullptr, 16));
        i += 2;
      } else {
        out.push_back(t[i]);
      }
    }
    return out;
  }
};

Instr read_instr(Reader& r) {
  r.expect("I");
  Instr i;
  i.op = op_from_name(r.next());
  i.res = r.i32();
  i.tensor = r.i32();
  i.dim = r.i32();
  int n = r.i32();
  for (int k = 0; k < n; ++k) i.args.push_back(r.i32());
  i.scalar = r.f64();
  i.index = r.i64();
  return i;
}

LinearIndex read_li(Reader& r) {
  r.expect("LI");
  LinearIndex li;
  int ns = r.i32(), nf = r.i32();
  li.constant = r.i64();
  for (int k = 0; k < ns; ++k) li.setup.push_back(read_instr(r));
  for (int k = 0; k < nf; ++k) {
    int reg = r.i32();
    li.factors[reg] = r.i64();
  }
  return li;
}

TensorOp read_op(Reader& r, const char* tag) {
  r.expect(tag);
  TensorOp op;
  op.tensor = r.i32();
  op.is_raw = r.i32() != 0;
  op.data = r.i32();
  int n = r.i32();
  for (int k = 0; k < n; ++k) op.dims.push_back(read_li(r));
  return op;
}

std::shared_ptr<Kernel> read_kernel(Reader& r) {
  r.expect("K");
  auto k = std::make_shared<Kernel>();
  k->gen = (GenKind)r.i32();
  k->gen_tensor = r.i32();
  int nresh = r.i32();
  for (int i = 0; i < nresh; ++i) k->reshape.push_back(r.i64());
  k->nregs = r.i32();
  int nloops = r.i32(), nreads = r.i32(), ninstrs = r.i32();
  k->res = r.i32();
  int has_custom = r.i32();
  for (int i = 0; i < nloops; ++i) {
    r.expect("L");
    Loop l;
    l.iter = r.i32();
    l.has_bounds = r.i32() != 0;
    l.step = r.i64();
    l.mode = r.i32();
    l.start = read_li(r);
    l.stop = read_li(r);
    k->loops.push_back(l);
  }
  for (int i = 0; i < nreads; ++i) k->reads.push_back(read_op(r, "R"));
  for (int i = 0; i < ninstrs; ++i) k->instrs.push_back(read_instr(r));
  k->write = read_op(r, "W");
  if (has_custom) {
    r.expect("C");
    k->custom_grad = std::make_shared<CustomGrad>();
    int nt = r.i32();
    for (int i = 0; i < nt; ++i) {
      int t = r.i32();
      k->custom_grad->tensors[t] = r.i32();
    }
    int nsub = r.i32();
    for (int i = 0; i < nsub; ++i) {
      int a = r.i32();
      k->custom_grad->subs[a] = r.i32();
    }
    int nk = r.i32();
    for (int i = 0; i < nk; ++i) k->custom_grad->kernels.push_back(read_kernel(r));
  }
  return k;
}

}  // namespace

std::shared_ptr<Program> parse_program(const std::string& text) {
  Reader r(text);
  r.expect("egbprog");
  if (r.i32() != 1) fail(EGB_ERR_PARSER, "unsupported program text version");
  auto prog = std::make_shared<Program>();
  std::string st = r.next();
  if (st == "f64") prog->f64 = true;
  else if (st != "f32") fail(EGB_ERR_PARSER, "unknown scalar type '%s'", st.c_str());
  r.expect("tensors");
  int nt = r.i32();
  for (int i = 0; i < nt; ++i) {
    r.expect("T");
    TensorDef t;
    t.kind = (TensorKind)r.i32();
    int rank = r.i32();
    for (int d = 0; d < rank; ++d) t.shape.push_back(r.i64());
    t.range_lo = r.f64();
    t.range_hi = r.f64();
    t.cache = r.i32();
    t.name = r.str();
    prog->tensors.push_back(t);
  }
  r.expect("targets");
  int ntg = r.i32();
  for (int i = 0; i < ntg; ++i) {
    r.expect("target");
    auto t = std::make_shared<Target>();
    t->name = r.str();
    t->output = r.i32();
    t->compile_target = r.i32();
    int ns = r.i32(), nk = r.i32();
    for (int s = 0; s < ns; ++s) {
      r.expect("S");
      ShapeConstraint sc;
      std::string kind = r.next();
      sc.dest = r.i32();
      sc.priority = r.i32();
      if (kind == "copy") {
        sc.kind = ShapeKind::Copy;
        sc.src = r.i32();
      } else if (kind == "dims") {
        sc.kind = ShapeKind::Dims;
        int n = r.i32();
        for (int d = 0; d < n; ++d) sc.dims.push_back(read_li(r));
      } else {
        fail(EGB_ERR_PARSER, "unknown shape constraint kind '%s'", kind.c_str());
      }
      t->shapes.push_back(sc);
    }
    for (int k = 0; k < nk; ++k) t->kernels.push_back(read_kernel(r));
    prog->targets.push_back(t);
  }
  r.expect("end");
  return prog;
}

std::string describe_kernel(const Kernel& k) {
  std::ostringstream ss;
  ss << "t" << k.write.tensor << (k.write.is_raw ? "{" : "[");
  for (size_t i = 0; i < k.write.dims.size(); ++i) ss << (i ? "," : "") << "d";
  ss << (k.write.is_raw ? "}" : "]") << " ++= f(";
  for (size_t i = 0; i < k.reads.size(); ++i) ss << (i ? "," : "") << "t" << k.reads[i].tensor;
  ss << ") loops=" << k.loops.size() << " instrs=" << k.instrs.size();
  return ss.str();
}

}  // namespace egb
