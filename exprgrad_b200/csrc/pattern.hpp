// Structural matching of kernel expressions and index patterns against the forms exprgrad's layer library
// produces (exprgrad/layers/base.nim:19-67, dnn.nim:19-100) and `derive` turns them into
// (exprgrad/passes.nim:383-549).
//
// A pattern is written as the canonical expression text (the same notation `describe_kernel` prints), e.g.
//     select(le(0,$0),$0,0)                      relu          dnn.nim:26-27
//     mul($1,select(le(0,$0),1,#0))              adjoint of leakyRelu
// and is PARSED into a tree; `unify` then walks the kernel's instruction DAG (ir.nim Instr: op + argument
// registers) against that tree:
//     $n   binds to read operand n of the pattern (any read of the kernel; the same $n must bind the same read)
//     #n   binds to a scalar literal and captures its value
//     1, 0.5 ...  a scalar literal with exactly this value
//     shape($n,d) the Shape instruction of dimension d of the tensor operand $n reads (ir.nim InstrShape)
//     name(args)  an instruction with this opcode; Add / Mul / Eq / And / Or also match with swapped operands
// so the operand order of commutative operations, the numbering of registers and the order of the kernel's
// reads do not matter - only the computation does.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "program.hpp"

namespace egb {

struct PatNode {
  enum Kind { OP, READ, CAPTURE, CONST, SHAPE } kind = OP;
  Op op = Op::Invalid;
  int index = 0;      // READ / SHAPE: operand number, CAPTURE: capture number
  double value = 0;   // CONST; SHAPE: dimension
  std::vector<PatNode> args;
};

PatNode parse_pattern(const std::string& text);

struct PatMatch {
  std::map<int, int> read_of;     // pattern operand -> index into kernel.reads
  std::map<int, int> tensor_of;   // pattern operand -> tensor id (set by a read or by shape($n, d))
  std::map<int, double> literal;  // capture -> value
};

// does the value of register `reg` of kernel `k` compute the pattern?
bool unify(const Kernel& k, int reg, const PatNode& pat, PatMatch& m);

// A whole kernel form: loop modes ("!." = independent, reduction), the index tuple of the write and of every
// pattern operand (in the notation of access_text_of: iterators I0, I1, ... by loop position) and the expression.
// The expression is unified structurally; the access tuples are compared per operand AFTER unification, so the
// order of the kernel's reads is irrelevant.
struct KernelForm {
  const char* loops;
  const char* write;
  const char* expr;
  std::vector<const char*> reads;   // access of $0, $1, ...
};
bool match_form(const Kernel& k, const KernelForm& form, PatMatch& m);

// Access structure of a map kernel: every loop independent, the write covers its whole tensor, and each read
// either uses the same index tuple as the write on a tensor of the same shape (SAME: identical flat offsets),
// is a row broadcast (ROW: the tensor is indexed by the write's last dimension only, dnn.nim:22-24), or is one
// fixed element (SCALAR: every index is a constant, e.g. the seed `dL[0]` that `derive` multiplies into the adjoint
// of a scalar loss, passes.nim:383-403).
enum class MapAccess { NONE, SAME, ROW, SCALAR };
struct MapShape {
  int64_t n = 0;          // elements written
  int64_t row = 0;        // extent of the last write dimension
  std::vector<MapAccess> reads;
  std::vector<int64_t> scalar_offset;   // per read: flat element offset of a SCALAR access (0 otherwise)
};
bool match_map_shape(const Kernel& k, const ShapeTable& shapes, MapShape& out);

// Specialised streaming kernels (eltwise_stream.cu): which one computes kernel `k`, if any.
enum EltKind {
  ELT_NONE = 0,
  ELT_COPY, ELT_RELU, ELT_LEAKY, ELT_SIGMOID, ELT_TANH, ELT_SCALE, ELT_SCALE_NEG, ELT_DIV_CONST,
  ELT_ADD, ELT_SUB, ELT_MUL, ELT_RELU_ADJ, ELT_LEAKY_ADJ, ELT_SIGMOID_ADJ, ELT_TANH_ADJ,
  ELT_ADAM_M, ELT_ADAM_V, ELT_ADAM_STEP, ELT_BIAS_ROW, ELT_SQ_ADJ,
  ELT_KIND_COUNT
};
struct EltSpec {
  int kind = ELT_NONE;
  int nreads = 0;
  int read_tensor[3] = {0, 0, 0};   // tensors bound to $0, $1, $2
  bool row_read[3] = {false, false, false};
  bool scalar_read[3] = {false, false, false};   // operand is one fixed element (MapAccess::SCALAR) ...
  int64_t scalar_offset[3] = {0, 0, 0};          // ... at this flat offset of its tensor
  double lit[4] = {0, 0, 0, 0};     // captured literals (f64, rounded by the launcher like llvmgen.nim:213-218)
  bool uses_epoch = false;
  int64_t n = 0, row = 0;
};
bool match_eltwise(const Kernel& k, const ShapeTable& shapes, EltSpec& out);
const char* elt_kind_name(int kind);

}  // namespace egb
