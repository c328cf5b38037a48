// Structural expression / access matching (see pattern.hpp).
#include "pattern.hpp"

#include <ctype.h>
#include <stdlib.h>

#include "egb_internal.hpp"
#include "lower.hpp"

namespace egb {

namespace {

struct PatParser {
  const std::string& s;
  size_t pos = 0;
  explicit PatParser(const std::string& text) : s(text) {}
  void skip() {
    while (pos < s.size() && isspace((unsigned char)s[pos])) ++pos;
  }
  PatNode node() {
    skip();
    PatNode n;
    if (pos >= s.size()) fail(EGB_ERR_RUNTIME, "internal: empty pattern");
    const char c = s[pos];
    if (c == '$' || c == '#') {
      ++pos;
      size_t start = pos;
      while (pos < s.size() && isdigit((unsigned char)s[pos])) ++pos;
      n.kind = c == '$' ? PatNode::READ : PatNode::CAPTURE;
      n.index = atoi(s.substr(start, pos - start).c_str());
      return n;
    }
    if (isdigit((unsigned char)c) || c == '-' || c == '.') {
      char* end = nullptr;
      n.kind = PatNode::CONST;
      n.value = strtod(s.c_str() + pos, &end);
      pos = (size_t)(end - s.c_str());
      return n;
    }
    size_t start = pos;
    while (pos < s.size() && (isalnum((unsigned char)s[pos]) || s[pos] == '_')) ++pos;
    std::string name = s.substr(start, pos - start);
    if (name == "shape") {   // shape($n,d)
      skip();
      if (pos + 1 >= s.size() || s[pos] != '(' || s[pos + 1] != '$') fail(EGB_ERR_RUNTIME, "internal: pattern: shape($n,d)");
      pos += 2;
      size_t a0 = pos;
      while (pos < s.size() && isdigit((unsigned char)s[pos])) ++pos;
      n.kind = PatNode::SHAPE;
      n.index = atoi(s.substr(a0, pos - a0).c_str());
      if (pos >= s.size() || s[pos] != ',') fail(EGB_ERR_RUNTIME, "internal: pattern: shape($n,d)");
      ++pos;
      char* end = nullptr;
      n.value = (double)strtol(s.c_str() + pos, &end, 10);
      pos = (size_t)(end - s.c_str());
      if (pos >= s.size() || s[pos] != ')') fail(EGB_ERR_RUNTIME, "internal: pattern: shape($n,d)");
      ++pos;
      return n;
    }
    // opcode names are written in lower case (describe_kernel notation); op_from_name wants "Add", "ToScalar" ...
    static const char* names[] = {"Add", "Sub", "Mul", "Div", "IndexDiv", "Mod", "Wrap", "Negate", "Sin", "Cos", "Exp", "Pow",
                                  "Sqrt", "Log", "Log10", "Log2", "Ln", "Eq", "Lt", "Le", "And", "Or", "Select", "ToScalar",
                                  "ToIndex", "Epoch"};
    n.kind = PatNode::OP;
    n.op = Op::Invalid;
    for (const char* cand : names) {
      std::string low = cand;
      for (auto& ch : low) ch = (char)tolower(ch);
      if (low == name) n.op = op_from_name(cand);
    }
    if (n.op == Op::Invalid) fail(EGB_ERR_RUNTIME, "internal: unknown opcode '%s' in pattern", name.c_str());
    skip();
    if (pos >= s.size() || s[pos] != '(') fail(EGB_ERR_RUNTIME, "internal: pattern: expected '(' after %s", name.c_str());
    ++pos;
    skip();
    if (pos < s.size() && s[pos] == ')') {
      ++pos;
      return n;
    }
    for (;;) {
      n.args.push_back(node());
      skip();
      if (pos < s.size() && s[pos] == ',') {
        ++pos;
        continue;
      }
      if (pos < s.size() && s[pos] == ')') {
        ++pos;
        break;
      }
      fail(EGB_ERR_RUNTIME, "internal: malformed pattern '%s'", s.c_str());
    }
    return n;
  }
};

bool commutative(Op op) { return op == Op::Add || op == Op::Mul || op == Op::Eq || op == Op::And || op == Op::Or; }

const Instr* def_of(const Kernel& k, int reg) {
  for (auto& ins : k.instrs)
    if (ins.res == reg) return &ins;
  return nullptr;
}

}  // namespace

PatNode parse_pattern(const std::string& text) {
  PatParser p(text);
  PatNode n = p.node();
  p.skip();
  if (p.pos != text.size()) fail(EGB_ERR_RUNTIME, "internal: trailing text in pattern '%s'", text.c_str());
  return n;
}

bool unify(const Kernel& k, int reg, const PatNode& pat, PatMatch& m) {
  if (pat.kind == PatNode::READ) {
    for (size_t i = 0; i < k.reads.size(); ++i) {
      if (k.reads[i].data != reg) continue;
      auto it = m.read_of.find(pat.index);
      if (it != m.read_of.end()) return it->second == (int)i;
      auto tt = m.tensor_of.find(pat.index);
      if (tt != m.tensor_of.end() && tt->second != k.reads[i].tensor) return false;
      m.read_of[pat.index] = (int)i;
      m.tensor_of[pat.index] = k.reads[i].tensor;
      return true;
    }
    return false;
  }
  const Instr* ins = def_of(k, reg);
  if (!ins) return false;
  if (pat.kind == PatNode::CAPTURE) {
    if (ins->op != Op::Scalar) return false;
    auto it = m.literal.find(pat.index);
    if (it != m.literal.end()) return it->second == ins->scalar;
    m.literal[pat.index] = ins->scalar;
    return true;
  }
  if (pat.kind == PatNode::CONST) return ins->op == Op::Scalar && ins->scalar == pat.value;
  if (pat.kind == PatNode::SHAPE) {
    if (ins->op != Op::Shape || ins->dim != (int)pat.value) return false;
    auto tt = m.tensor_of.find(pat.index);
    if (tt != m.tensor_of.end()) return tt->second == ins->tensor;
    m.tensor_of[pat.index] = ins->tensor;
    return true;
  }
  if (ins->op != pat.op || ins->args.size() != pat.args.size()) return false;
  {
    PatMatch trial = m;
    bool ok = true;
    for (size_t i = 0; i < pat.args.size() && ok; ++i) ok = unify(k, ins->args[i], pat.args[i], trial);
    if (ok) {
      m = trial;
      return true;
    }
  }
  if (commutative(pat.op) && pat.args.size() == 2) {
    PatMatch trial = m;
    if (unify(k, ins->args[1], pat.args[0], trial) && unify(k, ins->args[0], pat.args[1], trial)) {
      m = trial;
      return true;
    }
  }
  return false;
}

bool match_form(const Kernel& k, const KernelForm& form, PatMatch& m) {
  if (loop_modes_text(k) != form.loops || access_text_of(k, -1) != form.write) return false;
  if (k.reads.size() != form.reads.size()) return false;
  PatMatch trial;
  if (!unify(k, k.write.data, parse_pattern(form.expr), trial)) return false;
  if (trial.read_of.size() != form.reads.size()) return false;
  for (size_t q = 0; q < form.reads.size(); ++q) {
    auto it = trial.read_of.find((int)q);
    if (it == trial.read_of.end() || access_text_of(k, it->second) != form.reads[q]) return false;
  }
  m = trial;
  return true;
}

bool match_map_shape(const Kernel& k, const ShapeTable& shapes, MapShape& out) {
  if (k.loops.empty() || k.write.dims.empty()) return false;
  for (auto& l : k.loops)
    if (l.mode < 1 || !l.has_bounds) return false;
  auto wsh = shapes.find(k.write.tensor);
  if (wsh == shapes.end()) return false;
  if (!covers_whole_tensor(k, shapes)) return false;
  // every loop is one write dimension (no reduction hidden behind an independent flag)
  if (k.write.dims.size() != k.loops.size()) return false;
  int64_t n = 1;
  for (auto d : wsh->second) n *= d;
  out.n = n;
  out.row = k.write.is_raw || wsh->second.empty() ? n : wsh->second.back();
  out.reads.clear();
  out.scalar_offset.clear();
  for (auto& r : k.reads) {
    auto rsh = shapes.find(r.tensor);
    if (rsh == shapes.end()) return false;
    MapAccess acc = MapAccess::NONE;
    int64_t scalar_off = 0;
    if (r.is_raw == k.write.is_raw && r.dims.size() == k.write.dims.size()) {
      bool same = true;
      for (size_t d = 0; d < r.dims.size() && same; ++d) same = r.dims[d].same_as(k.write.dims[d]);
      int64_t rn = 1;
      for (auto d : rsh->second) rn *= d;
      if (same && (r.is_raw ? rn == n : rsh->second == wsh->second)) acc = MapAccess::SAME;
    }
    if (acc == MapAccess::NONE && !k.write.is_raw && !r.is_raw && k.write.dims.size() >= 2 && r.dims.size() == 1 &&
        r.dims[0].same_as(k.write.dims.back()) && rsh->second.size() == 1 && rsh->second[0] == wsh->second.back())
      acc = MapAccess::ROW;
    if (acc == MapAccess::NONE && !r.dims.empty() && (r.is_raw ? r.dims.size() == 1 : r.dims.size() == rsh->second.size())) {
      // every index a constant inside the tensor: one fixed element, row-major flat offset
      bool fixed = true;
      int64_t rn = 1;
      for (auto d : rsh->second) rn *= d;
      for (size_t d = 0; d < r.dims.size() && fixed; ++d) {
        const LinearIndex& li = r.dims[d];
        const int64_t extent = r.is_raw ? rn : rsh->second[d];
        fixed = li.setup.empty() && li.factors.empty() && li.constant >= 0 && li.constant < extent;
        if (fixed) scalar_off = scalar_off * extent + li.constant;
      }
      if (fixed) acc = MapAccess::SCALAR;
      else scalar_off = 0;
    }
    if (acc == MapAccess::NONE) return false;
    out.reads.push_back(acc);
    out.scalar_offset.push_back(scalar_off);
  }
  return true;
}

namespace {

struct EltPattern {
  int kind;
  int nreads;
  const char* text;
  bool epoch;
  int scalar_operand = -1;   // the operand that must be a fixed element (MapAccess::SCALAR); every other one must not be
};

// The forms of exprgrad/layers/base.nim and dnn.nim and of their adjoints as `derive` emits them
// (passes.nim:383-549: select / div / exp / negate / mul rules applied to the forward expression).
const EltPattern kPatterns[] = {
    {ELT_COPY, 1, "$0", false},
    {ELT_RELU, 1, "select(le(0,$0),$0,0)", false},                                       // dnn.nim:26-27
    {ELT_LEAKY, 1, "mul(select(le(0,$0),1,#0),$0)", false},                              // dnn.nim:29-30
    {ELT_SIGMOID, 1, "div(1,add(1,exp(negate($0))))", false},                            // dnn.nim:32-33
    {ELT_TANH, 1, "div(sub(exp($0),exp(negate($0))),add(exp($0),exp(negate($0))))", false},   // dnn.nim:35-40
    {ELT_SCALE_NEG, 1, "mul(negate($0),#0)", false},                                     // gradientDescent, base.nim:37-38
    {ELT_SCALE, 1, "mul($0,#0)", false},                                                 // base.nim:23
    {ELT_DIV_CONST, 1, "div($0,#0)", false},                                             // base.nim:25
    {ELT_ADD, 2, "add($0,$1)", false},                                                   // base.nim:19
    {ELT_SUB, 2, "sub($0,$1)", false},                                                   // base.nim:20
    {ELT_MUL, 2, "mul($0,$1)", false},
    {ELT_RELU_ADJ, 2, "select(le(0,$0),$1,0)", false},
    {ELT_LEAKY_ADJ, 2, "mul($1,select(le(0,$0),1,#0))", false},
    {ELT_SIGMOID_ADJ, 2,
     "negate(mul(mul(negate(1),div($1,mul(add(1,exp(negate($0))),add(1,exp(negate($0)))))),exp(negate($0))))", false},
    {ELT_TANH_ADJ, 2,
     "add(negate(mul(add(mul(negate(sub(exp($0),exp(negate($0)))),div($1,mul(add(exp($0),exp(negate($0))),add(exp($0),"
     "exp(negate($0)))))),negate(div($1,add(exp($0),exp(negate($0)))))),exp(negate($0)))),mul(add(mul(negate(sub(exp($0),"
     "exp(negate($0)))),div($1,mul(add(exp($0),exp(negate($0))),add(exp($0),exp(negate($0)))))),div($1,add(exp($0),"
     "exp(negate($0))))),exp($0)))",
     false},
    // adam, base.nim:40-53: m += m*(b1-1) + (1-b1)*g ; v += v*(b2-1) + (1-b2)*g*g ; p += -eta*mhat/(sqrt(vhat)+eps)
    {ELT_ADAM_M, 2, "add(mul($0,sub(#0,1)),mul(sub(1,#0),$1))", false},
    {ELT_ADAM_V, 2, "add(mul($0,sub(#0,1)),mul(sub(1,#0),mul($1,$1)))", false},
    {ELT_ADAM_STEP, 2,
     "div(mul(negate(#0),div($0,sub(1,pow(#1,toscalar(epoch()))))),add(sqrt(div($1,sub(1,pow(#2,toscalar(epoch()))))),#3))",
     true},
    // adjoint of sq(x) = x * x under a scalar loss (derive of mul, passes.nim:399-403): d[i] = g[0] * x[i] + g[0] * x[i]
    // (two spellings: a commutative node is tried in the written order first, so each one pins $1 to the seed for one
    // operand order of the products)
    {ELT_SQ_ADJ, 2, "add(mul($1,$0),mul($1,$0))", false, 1},
    {ELT_SQ_ADJ, 2, "add(mul($0,$1),mul($0,$1))", false, 1},
};

}  // namespace

const char* elt_kind_name(int kind) {
  static const char* names[] = {"none", "copy", "relu", "leakyRelu", "sigmoid", "tanh", "scale", "sgd-axpy", "div-const", "add", "sub",
                                "mul", "relu-adjoint", "leakyRelu-adjoint", "sigmoid-adjoint", "tanh-adjoint", "adam-m", "adam-v",
                                "adam-step", "bias-row-add", "square-adjoint"};
  return kind >= 0 && kind < ELT_KIND_COUNT ? names[kind] : "?";
}

bool match_eltwise(const Kernel& k, const ShapeTable& shapes, EltSpec& out) {
  if (k.reads.empty() || k.reads.size() > 3) return false;
  MapShape ms;
  if (!match_map_shape(k, shapes, ms)) return false;
  static std::vector<PatNode> parsed;
  if (parsed.empty())
    for (auto& p : kPatterns) parsed.push_back(parse_pattern(p.text));
  for (size_t pi = 0; pi < parsed.size(); ++pi) {
    const EltPattern& p = kPatterns[pi];
    PatMatch m;
    if (!unify(k, k.write.data, parsed[pi], m)) continue;
    // every read of the kernel must be accounted for by the pattern (a read the expression ignores is fine to drop,
    // but then it would not have survived dead-code elimination - treat it as "no match")
    if ((int)m.read_of.size() != p.nreads) continue;
    bool any_row = false, all_row = true, scalars_ok = true;
    EltSpec s;
    s.kind = p.kind;
    s.nreads = p.nreads;
    for (int q = 0; q < p.nreads; ++q) {
      const int ri = m.read_of.at(q);
      s.read_tensor[q] = k.reads[ri].tensor;
      s.row_read[q] = ms.reads[ri] == MapAccess::ROW;
      s.scalar_read[q] = ms.reads[ri] == MapAccess::SCALAR;
      s.scalar_offset[q] = ms.scalar_offset[ri];
      scalars_ok = scalars_ok && s.scalar_read[q] == (q == p.scalar_operand);
      any_row = any_row || s.row_read[q];
      all_row = all_row && s.row_read[q];
    }
    // a fixed-element operand only where the form's kernel loads one (the binding of a commutative pattern is not
    // unique: mul($1,$0) may have bound the operands the other way round - try the next pattern rather than give up)
    if (!scalars_ok) continue;
    if (any_row) {
      // the only broadcast form with a dedicated kernel: out[y, x] (+)= b[x]  (bias add, dnn.nim:22-24)
      if (!(p.kind == ELT_COPY && all_row)) return false;
      s.kind = ELT_BIAS_ROW;
    }
    for (auto& kv : m.literal)
      if (kv.first < 4) s.lit[kv.first] = kv.second;
    s.uses_epoch = p.epoch;
    s.n = ms.n;
    s.row = ms.row;
    out = s;
    return true;
  }
  return false;
}

}  // namespace egb
