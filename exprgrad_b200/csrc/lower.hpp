// Kernel lowering interface (lower.cpp).
#pragma once
#include <map>
#include <set>

#include "egb_internal.hpp"
#include "interp.hpp"
#include "program.hpp"

namespace egb {

struct Lowered {
  IpProgram ip;
  int pb = 256, rb = 1, points_fast = 1;
  bool uses_epoch = false;
  int nslots = 0;
};

// `ptrs` maps tensor id -> device pointer. `overwrite`: this kernel is the first writer of its
// (zero-initialised) output and covers it completely, so it may store instead of accumulate
// (the reference's InstrOverwrite, passes.nim:882-897).
Lowered lower_kernel(const Kernel& k, const ShapeTable& shapes, const std::map<int, void*>& ptrs, int64_t epoch,
                     bool strict, bool overwrite, int sm_count);

bool covers_whole_tensor(const Kernel& k, const ShapeTable& shapes);
// Every loop runs over the full extent [0, size) of a tensor dimension it indexes alone.
bool kernel_loops_full(const Kernel& k, const ShapeTable& shapes);

struct GemmPattern {
  int a_tensor = 0, b_tensor = 0, c_tensor = 0;
  bool trans_a = false, trans_b = false;
  int64_t M = 0, N = 0, K = 0, lda = 0, ldb = 0, ldc = 0;
};
// C[m,n] += sum_k A[..]*B[..] with plain iterator indices and full-range loops
// (exprgrad/layers/base.nim:27-28 and its adjoints, passes.nim:519-549).
bool match_gemm(const Kernel& k, const ShapeTable& shapes, GemmPattern& g);

struct ConvPattern {
  enum Kind { FORWARD, D_FILTERS, D_IMAGES } kind = FORWARD;
  int img_tensor = 0, fil_tensor = 0, out_tensor = 0;  // roles: [N,H,W,C], [F,KH,KW,C], [N,OH,OW,F]
  int N = 0, H = 0, W = 0, C = 0, F = 0, KH = 0, KW = 0;
};
// out[n,y,x,f] += img[n,y+dy,x+dx,c] * w[f,dy,dx,c] (exprgrad/layers/dnn.nim:45-49) or one of its two
// adjoints (the written tensor decides which).
bool match_conv2(const Kernel& k, const ShapeTable& shapes, ConvPattern& c);

}  // namespace egb
