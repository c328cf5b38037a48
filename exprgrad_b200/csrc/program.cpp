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
  int i32() {
    const int64_t v = i64();
    if (v < INT32_MIN || v > INT32_MAX) fail(EGB_ERR_PARSER, "program text: %lld does not fit a 32-bit field", (long long)v);
    return (int)v;
  }
  // an element count: every element takes at least one token, so a count beyond the rest of the text is malformed
  int count() {
    const int v = i32();
    if (v < 0 || (size_t)v > toks.size() - pos) fail(EGB_ERR_PARSER, "program text: bad element count %d", v);
    return v;
  }
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
      if (t[i] == '%' && i + 2 < t.size() + 0 && i + 2 <= t.size() - 1) {
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
  int n = r.count();
  for (int k = 0; k < n; ++k) i.args.push_back(r.i32());
  i.scalar = r.f64();
  i.index = r.i64();
  return i;
}

LinearIndex read_li(Reader& r) {
  r.expect("LI");
  LinearIndex li;
  int ns = r.count(), nf = r.count();
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
  int n = r.count();
  for (int k = 0; k < n; ++k) op.dims.push_back(read_li(r));
  return op;
}

std::shared_ptr<Kernel> read_kernel(Reader& r) {
  r.expect("K");
  auto k = std::make_shared<Kernel>();
  k->gen = (GenKind)r.i32();
  k->gen_tensor = r.i32();
  int nresh = r.count();
  for (int i = 0; i < nresh; ++i) k->reshape.push_back(r.i64());
  k->nregs = r.i32();
  int nloops = r.count(), nreads = r.count(), ninstrs = r.count();
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
    int nt = r.count();
    for (int i = 0; i < nt; ++i) {
      int t = r.i32();
      k->custom_grad->tensors[t] = r.i32();
    }
    int nsub = r.count();
    for (int i = 0; i < nsub; ++i) {
      int a = r.i32();
      k->custom_grad->subs[a] = r.i32();
    }
    int nk = r.count();
    for (int i = 0; i < nk; ++i) k->custom_grad->kernels.push_back(read_kernel(r));
  }
  return k;
}

// Referential integrity of a parsed program: every tensor id, register and enumeration value the passes and the
// planner will index with is inside its table. The text comes from another process (the Nim front-end, a checkpoint
// file): a damaged one must be an error (ParserError), never an out-of-bounds access.
struct Validator {
  const Program& prog;
  int nt;
  void tensor(int id, bool allow_none, bool allow_placeholder, const char* what) const {
    if (id == 0 && allow_none) return;
    if (id < 0 && allow_placeholder) return;   // gradient placeholders inside customGrad kernels (CustomGrad::tensors)
    if (id < 1 || id > nt) fail(EGB_ERR_PARSER, "program text: %s refers to tensor %d of %d", what, id, nt);
  }
  static void reg(int r, int nregs, bool allow_none, const char* what) {
    if (r == 0 && allow_none) return;
    if (r < 1 || r > nregs) fail(EGB_ERR_PARSER, "program text: %s uses register %d of %d", what, r, nregs);
  }
  void instr(const Instr& i, int nregs, bool placeholder) const {
    if (i.op == Op::Invalid) fail(EGB_ERR_PARSER, "program text: invalid instruction");
    reg(i.res, nregs, false, "an instruction result");
    for (int a : i.args) reg(a, nregs, false, "an instruction argument");
    tensor(i.tensor, true, placeholder, "an instruction");
    if (i.dim < -EGB_MAX_RANK || i.dim > EGB_MAX_RANK) fail(EGB_ERR_PARSER, "program text: dimension %d out of range", i.dim);
  }
  void index(const LinearIndex& li, int nregs, bool placeholder) const {
    for (auto& s : li.setup) instr(s, nregs, placeholder);
    for (auto& kv : li.factors) reg(kv.first, nregs, false, "an index factor");
  }
  void access(const TensorOp& op, int nregs, bool placeholder, bool needs_tensor) const {
    tensor(op.tensor, !needs_tensor, placeholder, "a tensor access");
    reg(op.data, nregs, true, "a tensor access");
    if (op.dims.size() > (size_t)EGB_MAX_RANK) fail(EGB_ERR_PARSER, "program text: access of rank %zu", op.dims.size());
    for (auto& d : op.dims) index(d, nregs, placeholder);
  }
  void kernel(const Kernel& k, bool placeholder, int depth) const {
    if (depth > 4) fail(EGB_ERR_PARSER, "program text: customGrad kernels nested too deeply");
    if ((int)k.gen < 0 || (int)k.gen > (int)GenKind::Reshape) fail(EGB_ERR_PARSER, "program text: unknown generator kind %d", (int)k.gen);
    if (k.nregs < 0 || k.nregs > (1 << 20)) fail(EGB_ERR_PARSER, "program text: %d registers", k.nregs);
    tensor(k.gen_tensor, !k.is_generator(), placeholder, "a generator");
    for (auto& l : k.loops) {
      reg(l.iter, k.nregs, false, "a loop iterator");
      if (l.step == 0) fail(EGB_ERR_PARSER, "program text: loop with step 0");
      if (l.mode < 0 || l.mode > 1) fail(EGB_ERR_PARSER, "program text: unknown loop mode %d", l.mode);
      index(l.start, k.nregs, placeholder);
      index(l.stop, k.nregs, placeholder);
    }
    for (auto& r : k.reads) access(r, k.nregs, placeholder, true);
    for (auto& i : k.instrs) instr(i, k.nregs, placeholder);
    access(k.write, k.nregs, placeholder, !k.is_generator() || k.gen == GenKind::Gradient || k.gen == GenKind::Reshape);
    reg(k.res, k.nregs, true, "a kernel result");
    if (k.custom_grad) {
      for (auto& kv : k.custom_grad->tensors) tensor(kv.first, false, true, "a customGrad table");
      for (auto& kv : k.custom_grad->subs) {
        tensor(kv.first, false, true, "a customGrad substitution");
        tensor(kv.second, false, true, "a customGrad substitution");
      }
      for (auto& c : k.custom_grad->kernels) kernel(*c, true, depth + 1);
    }
  }
};

void validate_program(const Program& prog) {
  Validator v{prog, (int)prog.tensors.size()};
  for (auto& t : prog.tensors) {
    if ((int)t.kind < 0 || (int)t.kind > (int)TensorKind::Random) fail(EGB_ERR_PARSER, "program text: unknown tensor kind %d", (int)t.kind);
    if (t.shape.size() > (size_t)EGB_MAX_RANK) fail(EGB_ERR_PARSER, "program text: tensor of rank %zu", t.shape.size());
    v.tensor(t.cache, true, false, "a cache");
  }
  for (auto& t : prog.targets) {
    v.tensor(t->output, true, false, "a target output");
    if (t->compile_target < 0 || t->compile_target > 2) fail(EGB_ERR_PARSER, "program text: unknown compile target %d", t->compile_target);
    for (int id : t->tensors) v.tensor(id, false, false, "a target");
    for (auto& sc : t->shapes) {
      v.tensor(sc.dest, false, false, "a shape constraint");
      if (sc.kind == ShapeKind::Copy) v.tensor(sc.src, false, false, "a shape constraint");
      if (sc.priority < PRIO_CONDITION || sc.priority > PRIO_USER) fail(EGB_ERR_PARSER, "program text: unknown constraint priority %d", sc.priority);
      if (sc.rank < 0 || sc.rank > EGB_MAX_RANK) fail(EGB_ERR_PARSER, "program text: rank constraint %d", sc.rank);
      // constraint indices are evaluated on their own small register files: only the tensors they name are checked
      for (auto& d : sc.dims)
        for (auto& s : d.setup) v.tensor(s.tensor, true, false, "a shape constraint");
      for (auto& rd : sc.reads) {
        v.tensor(rd.first, false, false, "a shape constraint");
        for (auto& dim : rd.second)
          for (auto& li : dim)
            for (auto& s : li.setup) v.tensor(s.tensor, true, false, "a shape constraint");
      }
      for (auto& d : sc.write)
        for (auto& s : d.setup) v.tensor(s.tensor, true, false, "a shape constraint");
    }
    for (auto& k : t->kernels) v.kernel(*k, false, 0);
  }
  for (auto& kv : prog.grad_tensors)
    for (auto& pr : kv.second) {
      v.tensor(pr.first, false, false, "the gradient table");
      v.tensor(pr.second, false, false, "the gradient table");
    }
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
  prog->compiled = r.i32() != 0;
  r.expect("tensors");
  int nt = r.count();
  for (int i = 0; i < nt; ++i) {
    r.expect("T");
    TensorDef t;
    t.kind = (TensorKind)r.i32();
    int rank = r.count();
    for (int d = 0; d < rank; ++d) t.shape.push_back(r.i64());
    t.range_lo = r.f64();
    t.range_hi = r.f64();
    t.cache = r.i32();
    t.name = r.str();
    prog->tensors.push_back(t);
  }
  r.expect("targets");
  int ntg = r.count();
  for (int i = 0; i < ntg; ++i) {
    r.expect("target");
    auto t = std::make_shared<Target>();
    t->name = r.str();
    t->output = r.i32();
    t->compile_target = r.i32();
    int ns = r.count(), nk = r.count(), ntens = r.count();
    for (int k = 0; k < ntens; ++k) t->tensors.push_back(r.i32());
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
        int n = r.count();
        for (int d = 0; d < n; ++d) sc.dims.push_back(read_li(r));
      } else if (kind == "rank") {
        sc.kind = ShapeKind::Rank;
        sc.rank = r.i32();
      } else if (kind == "linear") {
        sc.kind = ShapeKind::Linear;
        int nr = r.count();
        for (int a = 0; a < nr; ++a) {
          int tensor = r.i32();
          int nd = r.count();
          std::vector<std::vector<LinearIndex>> dims(nd);
          for (int d = 0; d < nd; ++d) {
            int ni = r.count();
            for (int e = 0; e < ni; ++e) dims[d].push_back(read_li(r));
          }
          sc.reads.emplace_back(tensor, dims);
        }
        int nw = r.count();
        for (int d = 0; d < nw; ++d) sc.write.push_back(read_li(r));
      } else {
        fail(EGB_ERR_PARSER, "unknown shape constraint kind '%s'", kind.c_str());
      }
      t->shapes.push_back(sc);
    }
    for (int k = 0; k < nk; ++k) t->kernels.push_back(read_kernel(r));
    prog->targets.push_back(t);
  }
  // optional: the loss-gradient table recorded by `generate` (target -> (tensor, gradient tensor) pairs). A
  // compiled program that was serialised and parsed again (checkpoints, Model.save / load_model) keeps it, so the
  // data-parallel runtime still finds the parameter-gradient bucket. Programs compiled by passes.nim do not
  // carry it; the runtime then derives the bucket from the optimizer kernels (runtime.cpp).
  std::string tail = r.next();
  if (tail == "grads") {
    const int ng = r.count();
    for (int i = 0; i < ng; ++i) {
      r.expect("G");
      const std::string name = r.str();
      const int np = r.count();
      auto& table = prog->grad_tensors[name];
      for (int q = 0; q < np; ++q) {
        const int t = r.i32();
        table[t] = r.i32();
      }
    }
    tail = r.next();
  }
  if (tail != "end") fail(EGB_ERR_PARSER, "program text: expected 'end', got '%s'", tail.c_str());
  validate_program(*prog);
  if (prog->compiled) {
    for (size_t i = 0; i < prog->tensors.size(); ++i) {
      const TensorDef& t = prog->tensors[i];
      if (t.kind == TensorKind::Param) prog->params.push_back((int)i + 1);
      else if (t.kind == TensorKind::Cache) prog->caches.push_back((int)i + 1);
      else if (t.kind == TensorKind::Input) prog->inputs[t.name] = (int)i + 1;
    }
  }
  return prog;
}

// ------------------------------------------------------------------------------ serializer

namespace {

struct Writer {
  std::ostringstream ss;
  void tok(const std::string& s) { ss << s << ' '; }
  void i(int64_t v) { ss << v << ' '; }
  void f(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%a", v);
    ss << buf << ' ';
  }
  void str(const std::string& s) {
    if (s.empty()) {
      ss << "- ";
      return;
    }
    static const char* hex = "0123456789abcdef";
    for (unsigned char c : s) {
      if (c <= ' ' || c == '%' || c == '-' || c >= 127) ss << '%' << hex[c >> 4] << hex[c & 15];
      else ss << c;
    }
    ss << ' ';
  }
  void nl() { ss << '\n'; }
};

void write_instr(Writer& w, const Instr& i) {
  w.tok("I");
  w.tok(op_name(i.op));
  w.i(i.res);
  w.i(i.tensor);
  w.i(i.dim);
  w.i((int64_t)i.args.size());
  for (int a : i.args) w.i(a);
  w.f(i.scalar);
  w.i(i.index);
}

void write_li(Writer& w, const LinearIndex& li) {
  w.tok("LI");
  w.i((int64_t)li.setup.size());
  w.i((int64_t)li.factors.size());
  w.i(li.constant);
  for (auto& s : li.setup) write_instr(w, s);
  for (auto& kv : li.factors) {
    w.i(kv.first);
    w.i(kv.second);
  }
}

void write_op(Writer& w, const TensorOp& op, const char* tag) {
  w.tok(tag);
  w.i(op.tensor);
  w.i(op.is_raw ? 1 : 0);
  w.i(op.data);
  w.i((int64_t)op.dims.size());
  for (auto& d : op.dims) write_li(w, d);
}

void write_kernel(Writer& w, const Kernel& k) {
  w.tok("K");
  w.i((int)k.gen);
  w.i(k.gen_tensor);
  w.i((int64_t)k.reshape.size());
  for (auto v : k.reshape) w.i(v);
  w.i(k.nregs);
  w.i((int64_t)k.loops.size());
  w.i((int64_t)k.reads.size());
  w.i((int64_t)k.instrs.size());
  w.i(k.res);
  w.i(k.custom_grad ? 1 : 0);
  w.nl();
  for (auto& l : k.loops) {
    w.tok("L");
    w.i(l.iter);
    w.i(l.has_bounds ? 1 : 0);
    w.i(l.step);
    w.i(l.mode);
    write_li(w, l.start);
    write_li(w, l.stop);
    w.nl();
  }
  for (auto& r : k.reads) {
    write_op(w, r, "R");
    w.nl();
  }
  for (auto& i : k.instrs) write_instr(w, i);
  w.nl();
  write_op(w, k.write, "W");
  w.nl();
  if (k.custom_grad) {
    w.tok("C");
    w.i((int64_t)k.custom_grad->tensors.size());
    for (auto& kv : k.custom_grad->tensors) {
      w.i(kv.first);
      w.i(kv.second);
    }
    w.i((int64_t)k.custom_grad->subs.size());
    for (auto& kv : k.custom_grad->subs) {
      w.i(kv.first);
      w.i(kv.second);
    }
    w.i((int64_t)k.custom_grad->kernels.size());
    w.nl();
    for (auto& g : k.custom_grad->kernels) write_kernel(w, *g);
  }
}

}  // namespace

std::string serialize_program(const Program& prog) {
  Writer w;
  w.tok("egbprog");
  w.i(1);
  w.tok(prog.f64 ? "f64" : "f32");
  w.i(prog.compiled ? 1 : 0);
  w.nl();
  w.tok("tensors");
  w.i((int64_t)prog.tensors.size());
  w.nl();
  for (auto& t : prog.tensors) {
    w.tok("T");
    w.i((int)t.kind);
    w.i((int64_t)t.shape.size());
    for (auto d : t.shape) w.i(d);
    w.f(t.range_lo);
    w.f(t.range_hi);
    w.i(t.cache);
    w.str(t.name);
    w.nl();
  }
  w.tok("targets");
  w.i((int64_t)prog.targets.size());
  w.nl();
  for (auto& t : prog.targets) {
    w.tok("target");
    w.str(t->name);
    w.i(t->output);
    w.i(t->compile_target);
    w.i((int64_t)t->shapes.size());
    w.i((int64_t)t->kernels.size());
    w.i((int64_t)t->tensors.size());
    for (int id : t->tensors) w.i(id);
    w.nl();
    for (auto& sc : t->shapes) {
      w.tok("S");
      switch (sc.kind) {
        case ShapeKind::Copy:
          w.tok("copy"); w.i(sc.dest); w.i(sc.priority); w.i(sc.src);
          break;
        case ShapeKind::Dims:
          w.tok("dims"); w.i(sc.dest); w.i(sc.priority); w.i((int64_t)sc.dims.size());
          for (auto& d : sc.dims) write_li(w, d);
          break;
        case ShapeKind::Rank:
          w.tok("rank"); w.i(sc.dest); w.i(sc.priority); w.i(sc.rank);
          break;
        case ShapeKind::Linear:
          w.tok("linear"); w.i(sc.dest); w.i(sc.priority); w.i((int64_t)sc.reads.size());
          for (auto& rd : sc.reads) {
            w.i(rd.first);
            w.i((int64_t)rd.second.size());
            for (auto& dim : rd.second) {
              w.i((int64_t)dim.size());
              for (auto& li : dim) write_li(w, li);
            }
          }
          w.i((int64_t)sc.write.size());
          for (auto& d : sc.write) write_li(w, d);
          break;
      }
      w.nl();
    }
    for (auto& k : t->kernels) write_kernel(w, *k);
  }
  if (!prog.grad_tensors.empty()) {
    w.tok("grads");
    w.i((int64_t)prog.grad_tensors.size());
    w.nl();
    for (auto& kv : prog.grad_tensors) {
      w.tok("G");
      w.str(kv.first);
      w.i((int64_t)kv.second.size());
      for (auto& pr : kv.second) {
        w.i(pr.first);
        w.i(pr.second);
      }
      w.nl();
    }
  }
  w.tok("end");
  w.nl();
  return w.ss.str();
}

// Canonical text of a kernel: iterator-space, tensor accesses and the value expression as a tree.
// Reads are named R0, R1, ... in read order, loop iterators I0, I1, ... in loop order; independent
// loops are marked with '!'. The planner matches these strings against the shapes the reference's
// layer library produces (exprgrad/layers/base.nim, dnn.nim) to pick fused device kernels.
static std::string index_text(const LinearIndex& li, const std::map<int, int>& loop_pos) {
  std::ostringstream ss;
  bool first = true;
  for (auto& kv : li.factors) {
    if (!first) ss << "+";
    first = false;
    if (kv.second != 1) ss << kv.second << "*";
    auto it = loop_pos.find(kv.first);
    if (it != loop_pos.end()) ss << "I" << it->second;
    else ss << "r" << kv.first;
  }
  if (li.constant != 0 || first) ss << (first ? "" : "+") << li.constant;
  return ss.str();
}

static std::string access_text(const TensorOp& op, const std::map<int, int>& loop_pos) {
  std::ostringstream ss;
  ss << (op.is_raw ? "{" : "[");
  for (size_t i = 0; i < op.dims.size(); ++i) ss << (i ? "," : "") << index_text(op.dims[i], loop_pos);
  ss << (op.is_raw ? "}" : "]");
  return ss.str();
}

std::string expr_text(const Kernel& k, int reg, int depth) {
  if (depth > 64) return "?";
  for (size_t i = 0; i < k.reads.size(); ++i)
    if (k.reads[i].data == reg) return "R" + std::to_string(i);
  for (size_t i = 0; i < k.loops.size(); ++i)
    if (k.loops[i].iter == reg) return "I" + std::to_string(i);
  for (auto& ins : k.instrs) {
    if (ins.res != reg) continue;
    char buf[64];
    switch (ins.op) {
      case Op::Scalar: snprintf(buf, sizeof(buf), "%.17g", ins.scalar); return buf;
      case Op::Index: return "i" + std::to_string(ins.index);
      case Op::Boolean: return ins.index ? "true" : "false";
      case Op::Shape: return "shape(T" + std::to_string(ins.tensor) + "," + std::to_string(ins.dim) + ")";
      case Op::Len: return "len(T" + std::to_string(ins.tensor) + ")";
      case Op::ShapeLen: return "rank(T" + std::to_string(ins.tensor) + ")";
      default: break;
    }
    std::string s = op_name(ins.op);
    for (auto& c : s) c = (char)tolower(c);
    s += "(";
    for (size_t i = 0; i < ins.args.size(); ++i) s += (i ? "," : "") + expr_text(k, ins.args[i], depth + 1);
    return s + ")";
  }
  return "r" + std::to_string(reg);
}

std::string loop_modes_text(const Kernel& k) {
  std::string s;
  for (auto& l : k.loops) s += l.mode >= 1 ? "!" : ".";
  return s;
}

std::string access_text_of(const Kernel& k, int read_index) {
  std::map<int, int> loop_pos;
  for (size_t i = 0; i < k.loops.size(); ++i) loop_pos[k.loops[i].iter] = (int)i;
  return access_text(read_index < 0 ? k.write : k.reads[(size_t)read_index], loop_pos);
}

std::string describe_kernel(const Kernel& k) {
  std::map<int, int> loop_pos;
  std::ostringstream ss;
  ss << "loops=";
  for (size_t i = 0; i < k.loops.size(); ++i) {
    loop_pos[k.loops[i].iter] = (int)i;
    ss << (k.loops[i].mode >= 1 ? "!" : ".");
  }
  ss << " W" << access_text(k.write, loop_pos);
  for (size_t i = 0; i < k.reads.size(); ++i) ss << " R" << i << access_text(k.reads[i], loop_pos);
  ss << " : " << expr_text(k, k.write.data, 0);
  return ss.str();
}

}  // namespace egb
