// Host-side program representation of the B200 backend: the structured kernels exprgrad hands over
// at the point where its own pipeline would start CPU/OpenCL scheduling (exprgrad/model.nim:52-61).
// Data model follows exprgrad/ir.nim:41-270 (Instr, LinearIndex, Loop, TensorOp, Kernel,
// ShapeConstraint, TensorDef, Target, Program); ids are 1-based, 0 = none (ir.nim:289-317).
#pragma once
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace egb {

enum class Op : uint8_t {
  // literals
  Index, Scalar, Boolean,
  // math
  Add, Sub, Mul, Div, IndexDiv, Mod, Wrap, Negate, Sin, Cos, Exp, Pow, Sqrt, Log, Log10, Log2, Ln,
  // conditional
  Eq, Lt, Le, And, Or, Select,
  // conversions
  ToScalar, ToIndex,
  // tensor
  Shape, Len, ShapeLen,
  // array
  Array, ArrayLen, ArrayRead,
  // misc
  Epoch,
  Invalid
};

const char* op_name(Op op);
Op op_from_name(const std::string& s);

struct Instr {
  Op op = Op::Invalid;
  std::vector<int> args;
  int res = 0;
  int tensor = 0;
  double scalar = 0.0;  // scalarLit
  int64_t index = 0;    // indexLit / booleanLit
  int dim = 0;
};

struct LinearIndex {
  std::vector<Instr> setup;
  std::map<int, int64_t> factors;  // reg -> factor (ordered: deterministic iteration)
  int64_t constant = 0;

  static LinearIndex from_const(int64_t c) { LinearIndex l; l.constant = c; return l; }
  static LinearIndex from_reg(int r) { LinearIndex l; l.factors[r] = 1; return l; }
  LinearIndex scaled(int64_t b) const;
  LinearIndex plus(const LinearIndex& o) const;
  LinearIndex minus(const LinearIndex& o) const { return plus(o.scaled(-1)); }
  int only_register() const;  // passes.nim:995-999
  bool same_as(const LinearIndex& o) const;
};

struct Loop {
  int iter = 0;
  bool has_bounds = false;
  LinearIndex start, stop;
  int64_t step = 1;
  int mode = 0;  // 0 none, 1 independent
};

struct TensorOp {
  int tensor = 0;
  bool is_raw = false;
  std::vector<LinearIndex> dims;
  int data = 0;
};

enum class GenKind : uint8_t { None, Backwards, Gradient, Reshape };

struct Kernel;
struct CustomGrad {
  std::map<int, int> tensors;  // tensor -> placeholder grad id (negative)
  std::vector<std::shared_ptr<Kernel>> kernels;
  std::map<int, int> subs;
};

struct Kernel {
  GenKind gen = GenKind::None;
  int gen_tensor = 0;
  std::vector<int64_t> reshape;
  std::shared_ptr<CustomGrad> custom_grad;
  int nregs = 0;
  std::vector<Loop> loops;
  std::vector<TensorOp> reads;
  std::vector<Instr> instrs;
  int res = 0;
  TensorOp write;
  int alloc_reg() { return ++nregs; }
  bool is_generator() const { return gen != GenKind::None; }
  std::shared_ptr<Kernel> clone() const;
  void substitute_tensors(const std::map<int, int>& subs);
};

enum class ShapeKind : uint8_t { Dims, Linear, Copy, Rank };
enum { PRIO_CONDITION = 0, PRIO_INFERRED = 1, PRIO_USER = 2 };

struct ShapeConstraint {
  ShapeKind kind = ShapeKind::Copy;
  int dest = 0;
  int priority = PRIO_INFERRED;
  int rank = 0;
  std::vector<LinearIndex> dims;
  // reads: tensor -> per dim list of indices, in first-appearance order
  std::vector<std::pair<int, std::vector<std::vector<LinearIndex>>>> reads;
  std::vector<LinearIndex> write;
  int src = 0;
};

enum class TensorKind : uint8_t { Result, Input, Param, Cache, Random };

struct TensorDef {
  TensorKind kind = TensorKind::Result;
  std::vector<int64_t> shape;
  std::string name;
  double range_lo = -0.1, range_hi = 0.1;  // initRange / randomRange
  int cache = 0;
};

struct Target {
  std::string name;
  int output = 0;
  int compile_target = 2;  // 0 cpu, 1 threads, 2 gpu (ir.nim:205)
  std::vector<int> tensors;  // insertion-ordered set
  std::vector<ShapeConstraint> shapes;
  std::vector<std::shared_ptr<Kernel>> kernels;
};

struct Program {
  std::vector<TensorDef> tensors;  // id = index + 1
  std::map<std::string, int> inputs;
  std::vector<int> params, caches;
  std::vector<std::shared_ptr<Target>> targets;  // declaration order
  bool f64 = false;
  bool compiled = false;
  // loss-gradient bookkeeping recorded by `generate`: tensor -> its gradient tensor, per target name.
  // The data-parallel runtime uses it to find the parameter-gradient bucket.
  std::map<std::string, std::map<int, int>> grad_tensors;
  TensorDef& tdef(int id) { return tensors[id - 1]; }
  const TensorDef& tdef(int id) const { return tensors[id - 1]; }
  int alloc_tensor(const TensorDef& t) { tensors.push_back(t); return (int)tensors.size(); }
  Target* find_target(const std::string& name);
};

// Text form produced by the front-end (see exprgrad_b200/frontend.py `serialize`, and the Nim
// serializer sketched in INTEGRATION.md). Accepts a source program (stage 0: what `toProgram`
// produces, parser.nim:404-417) or an already compiled one (stage 1: after the semantics-defining
// passes, i.e. what exprgrad's own passes.nim hands to a code generator).
std::shared_ptr<Program> parse_program(const std::string& text);
std::string serialize_program(const Program& prog);

// The semantics-defining prefix of exprgrad/model.nim:46-77 (see passes.cpp).
void compile_program(Program& prog);

typedef std::map<int, std::vector<int64_t>> ShapeTable;
// exprgrad/passes.nim:1386-1436 - run-time shape inference, integer exact.
ShapeTable infer_shapes(const Program& prog, const Target& target, const ShapeTable& inputs);

std::string describe_kernel(const Kernel& k);
// canonical pieces of it: "!.!" (independent / reduction loops in loop order) and the index tuple of the write
// (read_index < 0) or of one read, e.g. "[I0,I1]", "{I0}", "[I0,2*I1+1,I3]"
std::string loop_modes_text(const Kernel& k);
std::string access_text_of(const Kernel& k, int read_index);
std::string expr_text(const Kernel& k, int reg, int depth = 0);

}  // namespace egb
