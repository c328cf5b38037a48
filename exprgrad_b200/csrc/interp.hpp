// Device program of the generic loop-nest kernel (interp.cu): a lowered exprgrad `Kernel`
// (exprgrad/ir.nim:211-220) with every shape-dependent quantity already evaluated on the host.
//
// One output POINT = one assignment of the independent loops (LoopIndependent, passes.nim:1774-1781:
// iterators that appear alone in a write dimension). A group of `group` threads (1, 32 or the whole
// block) owns a point, walks the remaining (reduction) loops, and accumulates into the output.
// With group == 1 the additions happen in the reference's order (ascending nested loops, separate
// fmul/fadd because the file is compiled with -fmad=false), which makes results bit-identical to the
// reference's CPU lowering (llvmgen.nim:219-221, 277-297) up to libm differences in exp/log/pow.
#pragma once
#include <stdint.h>

namespace egb {

constexpr int IP_MAX_LOOPS = 8;
constexpr int IP_MAX_OPS = 12;     // reads
constexpr int IP_MAX_TERMS = 8;    // (slot, coefficient) terms of one flattened tensor index
constexpr int IP_MAX_INSTRS = 160;
constexpr int IP_MAX_SLOTS = 192;
constexpr int IP_SMEM_SLOTS = 22;  // 22 slots x 256 threads x 8 B = 44 KB of shared memory

enum IpOp : uint8_t {
  IP_NOP = 0,
  // scalar (fp32)
  IP_FADD, IP_FSUB, IP_FMUL, IP_FDIV, IP_FNEG, IP_SIN, IP_COS, IP_EXP, IP_LN, IP_SQRT, IP_POW, IP_LOG10, IP_LOG2,
  IP_LOGB,
  // index (int64)
  IP_IADD, IP_ISUB, IP_IMUL, IP_IDIV, IP_IMOD, IP_IWRAP, IP_INEG,
  // comparisons -> boolean
  IP_FEQ, IP_FLT, IP_FLE, IP_IEQ, IP_ILT, IP_ILE, IP_BEQ,
  IP_AND, IP_OR,
  IP_SELECT,     // dst = slots[a].b ? slots[b] : slots[c]   (raw 64-bit move)
  IP_TOSCALAR,   // sitofp
  IP_TOINDEX,    // fptosi (truncate)
  IP_ARRAY_READ  // dst = slots[array_table[imm + slots[a].i]]
};

struct IpInstr {
  uint8_t op, dst, a, b;
  uint8_t c, pad;
  uint16_t imm;
};

struct IpTensorOp {
  uint64_t base;                     // device pointer (float*)
  int64_t offset;                    // constant part of the flat element index
  int64_t coef[IP_MAX_TERMS];
  uint8_t slot[IP_MAX_TERMS];
  uint8_t nterms;
  uint8_t dst;                       // reads: destination slot; write: value slot
  uint8_t streaming;                 // vec4 path: index = offset + loop iterator (else: loop-invariant)
  uint8_t aligned16;                 // vec4 path: element (offset + loop start) is 16-byte aligned
  uint8_t pad[4];
};

struct IpLoop {
  int64_t start, step, count;        // iter = start + step * i, i in [0, count)
  uint8_t slot;
  uint8_t pad[7];
};

struct alignas(16) IpProgram {
  // loops[0 .. npar) are the independent loops (thread-mapped, last one fastest);
  // loops[npar .. nloops) are reduction loops in the reference's nesting order (outermost first).
  IpLoop loops[IP_MAX_LOOPS];
  IpTensorOp reads[IP_MAX_OPS];
  IpTensorOp write;
  IpInstr index_instrs[32];          // per-point index arithmetic that is not affine (IndexDiv/Mod/Wrap ...)
  IpInstr instrs[IP_MAX_INSTRS];
  uint64_t lits[48];                 // literal pool (raw 64-bit: float in the low word, or int64)
  uint8_t lit_slot[48];              // slot that holds literal i (loaded once per thread)
  uint8_t array_table[64];
  int64_t npoints;                   // product of the independent loop counts
  int64_t nred;                      // product of the reduction loop counts
  uint8_t nloops, npar, nreads, ninstrs, nindex_instrs, nlits;
  uint8_t accumulate;                // 1: out += value (InstrWrite), 0: out = value (InstrOverwrite)
  uint8_t scatter;                   // write index depends on a reduction loop: read-modify-write per iteration
  uint8_t vec4;                      // pure streaming elementwise kernel: eligible for the 4-wide fast path
  uint8_t nslots;                    // registers used (<= IP_SMEM_SLOTS: the register file lives in shared memory)
};

static_assert(sizeof(IpProgram) % 16 == 0, "IpProgram is copied in 16-byte units");
static_assert(sizeof(IpProgram) <= 4000, "IpProgram travels as a kernel parameter");

}  // namespace egb
