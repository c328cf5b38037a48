// Launch-plan construction and execution (see runtime.hpp).
#include "runtime.hpp"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <set>

#include "lower.hpp"

namespace egb {

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int64_t shape_len(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

std::string shape_text(const std::vector<int64_t>& s) {
  std::string out = "[";
  for (size_t i = 0; i < s.size(); ++i) out += (i ? "," : "") + std::to_string(s[i]);
  return out + "]";
}

// splitmix64: parameter initialisation U(initRange) (model.nim:244-247). The reference draws from
// Nim's global RNG, which is not reproducible elsewhere; parity tests inject parameters explicitly.
struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace

Plan::~Plan() {
  if (window.mapped) comm_close_window(window);
  for (void* b : chain_bufs) cudaFree(b);
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (arena) cudaFree(arena);
}

Model::~Model() {
  if (ctx) cudaStreamSynchronize(ctx->stream);
  plans.clear();
  for (auto& kv : state)
    if (kv.second.owned && kv.second.ptr) cudaFree(kv.second.ptr);
}

// ------------------------------------------------------------------ plan construction

namespace {

void* tensor_ptr(Model& m, Plan& plan, int id) {
  auto s = m.state.find(id);
  if (s != m.state.end()) return s->second.ptr;
  auto b = plan.bound.find(id);
  if (b != plan.bound.end() && b->second) return const_cast<void*>(b->second);
  auto t = plan.tensors.find(id);
  if (t != plan.tensors.end()) return t->second.ptr;
  return nullptr;
}

// levels: a node's level is one more than the highest level among the earlier nodes it conflicts
// with (read-after-write, write-after-read, write-after-write); nodes of one level are independent
void compute_levels(Plan& plan) {
  auto hits = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
    for (auto x : a)
      for (auto y : b)
        if (x == y) return true;
    return false;
  };
  for (size_t i = 0; i < plan.nodes.size(); ++i) {
    Node& ni = plan.nodes[i];
    int level = 0;
    for (size_t j = 0; j < i; ++j) {
      const Node& nj = plan.nodes[j];
      if (hits(nj.writes, ni.reads) || hits(nj.writes, ni.writes) || hits(nj.reads, ni.writes))
        level = std::max(level, nj.level + 1);
    }
    ni.level = level;
  }
}

// Rewrite a lowered program so that loops[0] is a row loop of `rows` iterations and check that every
// access to a tensor written inside the group stays within the row (see interp_rowchain_kernel).
bool to_row_form(IpProgram& p, int64_t rows, const std::map<uint64_t, int64_t>& row_stride) {
  if (p.scatter || p.npar == 0 || rows <= 0) return false;
  auto out_stride = row_stride.find(p.write.base);
  if (out_stride == row_stride.end()) return false;
  auto coef_of = [](const IpTensorOp& op, int slot) {
    int64_t c = 0;
    for (int t = 0; t < op.nterms; ++t)
      if (op.slot[t] == slot) c = op.coef[t];
    return c;
  };
  int row = -1;
  for (int l = 0; l < p.npar; ++l)
    if (p.loops[l].count == rows && p.loops[l].start == 0 && p.loops[l].step == 1 &&
        coef_of(p.write, p.loops[l].slot) == out_stride->second && row < 0)
      row = l;
  if (row < 0) {
    // a single raw loop over rows * width elements: split it into (row, column)
    if (p.nloops != 1 || p.loops[0].start != 0 || p.loops[0].step != 1 || p.loops[0].count % rows != 0) return false;
    const int64_t width = p.loops[0].count / rows;
    if (width != out_stride->second) return false;
    const int old = p.loops[0].slot;
    auto arity = [](uint8_t op) {
      switch (op) {
        case IP_FNEG: case IP_SIN: case IP_COS: case IP_EXP: case IP_LN: case IP_SQRT: case IP_LOG10: case IP_LOG2:
        case IP_INEG: case IP_TOSCALAR: case IP_TOINDEX: case IP_ARRAY_READ: return 1;
        case IP_SELECT: return 3;
        default: return 2;
      }
    };
    for (int i = 0; i < p.ninstrs; ++i) {  // the iterator must not be used as a value
      const IpInstr& in = p.instrs[i];
      const int n = arity(in.op);
      if (in.a == old || (n >= 2 && in.b == old) || (n >= 3 && in.c == old)) return false;
    }
    if (p.nindex_instrs) return false;
    // find an unused slot for the row iterator
    bool used[IP_MAX_SLOTS] = {false};
    used[old] = true;
    for (int i = 0; i < p.ninstrs; ++i) used[p.instrs[i].dst] = used[p.instrs[i].a] = used[p.instrs[i].b] = used[p.instrs[i].c] = true;
    for (int i = 0; i < p.nlits; ++i) used[p.lit_slot[i]] = true;
    for (int q = 0; q < p.nreads; ++q) used[p.reads[q].dst] = true;
    used[p.write.dst] = true;
    int rs = -1;
    for (int i = 0; i < IP_MAX_SLOTS && rs < 0; ++i)
      if (!used[i]) rs = i;
    if (rs < 0) return false;
    auto split = [&](IpTensorOp& op) {
      const int64_t c = coef_of(op, old);
      if (c == 0) return true;
      if (op.nterms >= IP_MAX_TERMS) return false;
      op.slot[op.nterms] = (uint8_t)rs;
      op.coef[op.nterms] = c * width;
      op.nterms++;
      return true;
    };
    for (int q = 0; q < p.nreads; ++q)
      if (!split(p.reads[q])) return false;
    if (!split(p.write)) return false;
    p.loops[1] = p.loops[0];
    p.loops[1].count = width;
    p.loops[0].slot = (uint8_t)rs;
    p.loops[0].start = 0;
    p.loops[0].step = 1;
    p.loops[0].count = rows;
    p.nloops = 2;
    p.npar = 2;
  } else if (row != 0) {
    const IpLoop r = p.loops[row];
    for (int l = row; l > 0; --l) p.loops[l] = p.loops[l - 1];
    p.loops[0] = r;
  }
  // row locality of every access to a tensor the group writes
  const int rslot = p.loops[0].slot;
  auto local = [&](const IpTensorOp& op) {
    auto st = row_stride.find(op.base);
    if (st == row_stride.end()) return true;  // external tensor: read-only for the whole group
    if (coef_of(op, rslot) != st->second) return false;
    int64_t lo = op.offset, hi = op.offset;
    for (int t = 0; t < op.nterms; ++t) {
      if (op.slot[t] == rslot) continue;
      int l = -1;
      for (int q = 0; q < p.nloops; ++q)
        if (p.loops[q].slot == op.slot[t]) l = q;
      if (l < 0) return false;  // index computed by per-point instructions: range unknown
      const int64_t first = p.loops[l].start, last = p.loops[l].start + p.loops[l].step * (p.loops[l].count - 1);
      const int64_t a = op.coef[t] * first, b = op.coef[t] * last;
      lo += std::min(a, b);
      hi += std::max(a, b);
    }
    return lo >= 0 && hi < st->second;
  };
  for (int q = 0; q < p.nreads; ++q)
    if (!local(p.reads[q])) return false;
  return local(p.write);
}

// db[x] = sum_y dh[y, x]: the adjoint of the row-broadcast bias add (dnn.nim:22-24 through derive)
bool is_column_sum(const Kernel& k) {
  static const KernelForm form{".!", "[I1]", "$0", {"[I0,I1]"}};
  PatMatch m;
  return match_form(k, form, m);
}

// adam as one pass over the parameter (eltwise_stream.cu): three ELTWISE nodes - first moment, second moment, step -
// that the reference emits per parameter (layers/base.nim:40-53) and that read each other's outputs. Matched on the
// lowered nodes: same length, the same gradient in both moment updates, the step reading exactly those two caches,
// nothing in between that touches any of the four tensors.
bool fuse_adam(Plan& plan) {
  auto hits = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
    for (auto x : a)
      for (auto y : b)
        if (x == y) return true;
    return false;
  };
  bool any = false;
  for (int i = 0; i < (int)plan.nodes.size(); ++i) {
    const Node& a = plan.nodes[i];
    if (a.kind != Node::ELTWISE || a.elt.kind != ELT_ADAM_M || !a.elt.accumulate || a.elt.in[0] != a.elt.out) continue;
    int jb = -1, jc = -1;
    for (int j = i + 1; j < (int)plan.nodes.size() && j <= i + 8; ++j) {
      const Node& n = plan.nodes[j];
      if (n.kind != Node::ELTWISE) continue;
      if (jb < 0 && n.elt.kind == ELT_ADAM_V && n.elt.accumulate && n.elt.in[0] == n.elt.out && n.elt.in[1] == a.elt.in[1] &&
          n.elt.n == a.elt.n)
        jb = j;
      else if (jb >= 0 && jc < 0 && n.elt.kind == ELT_ADAM_STEP && n.elt.accumulate && n.elt.in[0] == a.elt.out &&
               n.elt.in[1] == plan.nodes[jb].elt.out && n.elt.n == a.elt.n)
        jc = j;
    }
    if (jb < 0 || jc < 0) continue;
    bool clean = true;
    for (int j = i + 1; j < jc && clean; ++j) {
      if (j == jb) continue;
      const Node& n = plan.nodes[j];
      for (int q : {i, jb, jc})
        clean = clean && !hits(n.writes, plan.nodes[q].reads) && !hits(n.writes, plan.nodes[q].writes) && !hits(n.reads, plan.nodes[q].writes);
    }
    if (!clean) continue;
    const Node &b = plan.nodes[jb], &c = plan.nodes[jc];
    Node f;
    f.kind = Node::ELTWISE;
    f.elt.kind = ELT_ADAM_FUSED;
    f.elt.out = c.elt.out;
    f.elt.in[0] = a.elt.in[1];
    f.elt.nreads = 1;
    f.elt.adam_m = a.elt.out;
    f.elt.adam_v = b.elt.out;
    f.elt.n = a.elt.n;
    f.elt.accumulate = true;
    f.elt.p[0] = a.elt.p[0]; f.elt.p[1] = a.elt.p[1];
    f.elt.p2[0] = b.elt.p[0]; f.elt.p2[1] = b.elt.p[1];
    for (int q = 0; q < 4; ++q) f.elt.p3[q] = c.elt.p[q];
    if (!eltwise_stream_supported(f.elt)) continue;
    f.uses_epoch = true;
    f.kernel_index = c.kernel_index;
    f.label = "eltwise adam (kernels " + std::to_string(a.kernel_index) + " " + std::to_string(b.kernel_index) + " " +
              std::to_string(c.kernel_index) + " in one pass) n=" + std::to_string((long long)a.elt.n);
    for (int q : {i, jb, jc}) {
      for (auto r : plan.nodes[q].reads) f.reads.push_back(r);
      for (auto w : plan.nodes[q].writes) f.writes.push_back(w);
    }
    std::vector<Node> kept;
    for (int j = 0; j < (int)plan.nodes.size(); ++j) {
      if (j == i || j == jb) continue;
      kept.push_back(j == jc ? f : plan.nodes[j]);
    }
    plan.nodes = kept;
    any = true;
    i = -1;
  }
  return any;
}

// The classification head in one launch (head_rows.cu): a SOFTMAX_XENT node absorbs the contraction that produces its
// logits (<= 16 classes) and the contraction that consumes its gradient planes with the class dimension as k. Matched
// on the lowered nodes (operand pointers, shapes, fused stages) with hazard checks against every node in between;
// anything else stays as three launches, with a note in the plan.
bool fuse_head(Model& m, Plan& plan) {
  auto hits = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
    for (auto x : a)
      for (auto y : b)
        if (x == y) return true;
    return false;
  };
  bool any = false;
  for (int xi = 0; xi < (int)plan.nodes.size(); ++xi) {
    if (plan.nodes[xi].kind != Node::SOFTMAX_XENT) continue;
    const Node& X = plan.nodes[xi];
    if (!X.sx_out_hi) continue;
    int gi = -1, go = -1;
    for (int j = 0; j < xi; ++j) {
      const Node& n = plan.nodes[j];
      if (n.kind == Node::GEMM && n.gemm.C == X.sx_h) gi = j;
    }
    for (int j = xi + 1; j < (int)plan.nodes.size(); ++j) {
      const Node& n = plan.nodes[j];
      if (n.kind == Node::GEMM && n.gemm.a_hi == X.sx_out_hi && !n.gemm.a_mn && n.gemm.K == X.sx_cols) { go = j; break; }
    }
    if (gi < 0 || go < 0) continue;
    const GemmArgs& in = plan.nodes[gi].gemm;
    const GemmArgs& out = plan.nodes[go].gemm;
    std::string why;
    if (in.M != X.sx_rows || in.N != X.sx_cols || in.ldc != X.sx_cols || in.a_mn || in.epi != EPI_NONE || in.colsum ||
        (in.flags & ~GEMM_BIAS) != 0 || in.alpha != 1.0f)
      why = "the contraction in front of the rows has stages the head kernel does not take";
    else if (out.M != X.sx_rows || out.alpha != 1.0f || (out.flags & GEMM_ACCUMULATE) || out.a_mid != X.sx_out_mid ||
             out.lda != X.sx_ld_out)
      why = "the adjoint contraction behind the rows accumulates or reads other planes";
    // nothing between the absorbed nodes and the fused node's position may touch what moves past it
    for (int j = gi + 1; j < xi && why.empty(); ++j) {
      const Node& n = plan.nodes[j];
      if (hits(n.reads, plan.nodes[gi].writes) || hits(n.writes, plan.nodes[gi].writes) || hits(n.writes, plan.nodes[gi].reads))
        why = "a node between the logits contraction and the rows uses its tensors";
    }
    for (int j = xi + 1; j < go && why.empty(); ++j) {
      const Node& n = plan.nodes[j];
      if (hits(n.reads, plan.nodes[go].writes) || hits(n.writes, plan.nodes[go].writes) || hits(n.writes, plan.nodes[go].reads))
        why = "a node between the rows and the adjoint contraction uses its tensors";
    }
    HeadParams h;
    h.a_hi = in.a_hi; h.a_mid = in.a_mid; h.lda = in.lda; h.Kin = in.K;
    h.win_hi = in.b_hi; h.win_mid = in.b_mid; h.win_ld = in.ldb; h.win_mn = in.b_mn ? 1 : 0;
    h.bias_in = (in.flags & GEMM_BIAS) ? in.bias : nullptr;
    h.Z = in.C;
    h.Y = X.sx_y; h.DL = X.sx_dl; h.S = X.sx_s; h.P = X.sx_p; h.DP = X.sx_dp; h.DH = X.sx_dh; h.DS = X.sx_ds;
    h.colsum_dz = X.sx_colsum; h.dh_hi = X.sx_out_hi; h.dh_mid = X.sx_out_mid; h.dh_ld = X.sx_ld_out;
    h.rows = X.sx_rows; h.cols = X.sx_cols;
    h.wout_hi = out.b_hi; h.wout_mid = out.b_mid; h.wout_ld = out.ldb; h.wout_mn = out.b_mn ? 1 : 0; h.Nout = out.N;
    h.bias_out = (out.flags & GEMM_BIAS) ? out.bias : nullptr;
    h.epi = out.epi; h.epi_param = out.epi_param; h.Hm = out.H; h.C = out.C; h.D = out.D; h.ldc = out.ldc;
    h.flags = out.flags; h.colsum_out = out.colsum; h.out_hi = out.out_hi; h.out_mid = out.out_mid; h.ld_out = out.ld_out;
    if (why.empty() && !head_rows_supported(h)) why = "shapes or alignment outside the head kernel's range";
    if (!why.empty()) {
      plan.notes.push_back("classification head stays three launches: " + why);
      continue;
    }
    // fp32 tables of both weight operands: built once per run by a small node that depends only on the producers of
    // the weight planes (normally the plan's root split), so it runs beside the first contractions
    void* tab = nullptr;
    EGB_CUDA(cudaMalloc(&tab, head_table_floats(h) * sizeof(float)));
    plan.chain_bufs.push_back(tab);
    h.tables = (float*)tab;
    const int64_t table_id = -((int64_t)1 << 40) - xi;
    Node prep;
    prep.kind = Node::HEADPREP;
    prep.head = h;
    prep.label = "weight tables of the head kernel";
    prep.writes.push_back(table_id);
    int prep_pos = 0;
    for (int j = 0; j < gi; ++j) {
      const Node& n = plan.nodes[j];
      bool producer = false;
      if (n.kind == Node::SPLIT) {
        producer = n.split_hi == in.b_hi || n.split_hi == out.b_hi;
        for (auto& job : n.split_jobs) producer = producer || job.hi == in.b_hi || job.hi == out.b_hi;
      } else if (n.kind == Node::GEMM) {
        producer = (n.gemm.flags & GEMM_SPLIT_OUT) && (n.gemm.out_hi == in.b_hi || n.gemm.out_hi == out.b_hi);
      }
      if (!producer) continue;
      for (auto w : n.writes)
        if (std::find(plan.nodes[gi].reads.begin(), plan.nodes[gi].reads.end(), w) != plan.nodes[gi].reads.end() ||
            std::find(plan.nodes[go].reads.begin(), plan.nodes[go].reads.end(), w) != plan.nodes[go].reads.end())
          prep.reads.push_back(w);
      prep_pos = j + 1;
    }
    Node f;
    f.kind = Node::HEAD;
    f.head = h;
    f.label = "head: " + plan.nodes[gi].label + " | " + X.label + " | " + plan.nodes[go].label;
    for (int j : {gi, xi, go}) {
      for (auto r : plan.nodes[j].reads) f.reads.push_back(r);
      for (auto w : plan.nodes[j].writes) f.writes.push_back(w);
    }
    f.reads.push_back(table_id);
    std::vector<Node> kept;
    for (int j = 0; j < (int)plan.nodes.size(); ++j) {
      if (j == prep_pos) kept.push_back(prep);
      if (j == gi || j == go) continue;
      kept.push_back(j == xi ? f : plan.nodes[j]);
    }
    plan.nodes = kept;
    any = true;
    xi = -1;  // restart: indices moved
  }
  return any;
}

// Replace runs of consecutive levels that consist only of small row-local generic kernels by one
// ROWCHAIN node each. Returns true if anything was fused.
bool fuse_row_chains(Model& m, Plan& plan) {
  int max_level = 0;
  for (auto& n : plan.nodes) max_level = std::max(max_level, n.level);
  std::vector<std::vector<int>> by_level(max_level + 1);
  for (size_t i = 0; i < plan.nodes.size(); ++i) by_level[plan.nodes[i].level].push_back((int)i);
  auto tensor_len = [&](int64_t id) {
    int64_t len = 1;
    auto sh = plan.shapes.find((int)id);
    if (sh == plan.shapes.end()) return (int64_t)-1;
    for (auto d : sh->second) len *= d;
    return len;
  };
  std::set<int> remove;
  std::vector<std::pair<int, Node>> insert;  // (position of the first fused node, chain node)
  int l = 0;
  while (l <= max_level) {
    auto level_ok = [&](int lv) {
      if (by_level[lv].empty()) return false;
      for (int i : by_level[lv]) {
        const Node& n = plan.nodes[i];
        // chains are for latency-bound small kernels; larger ones run faster as separate (4-wide) launches
        if (n.kind != Node::INTERP || n.ip.scatter || n.uses_epoch || n.ip.npoints * n.ip.nred > (1 << 16)) return false;
      }
      return true;
    };
    if (!level_ok(l)) {
      ++l;
      continue;
    }
    // candidate rows: taken from the first node
    std::vector<int> members;
    int64_t rows = -1;
    int end = l;
    for (; end <= max_level && level_ok(end); ++end) {
      std::vector<int> trial = members;
      for (int i : by_level[end]) trial.push_back(i);
      // the chain's programs and its register file live in shared memory (<= 96 KB)
      int trial_slots = 1;
      for (int i : trial) trial_slots = std::max(trial_slots, (int)plan.nodes[i].ip.nslots + 1);  // +1: row iterator
      if (trial.size() * sizeof(IpProgram) + (size_t)trial_slots * 256 * 8 > 92 * 1024) break;
      // rows = leading dimension shared by all written tensors of the trial group
      int64_t r = -1;
      bool ok = true;
      std::map<uint64_t, int64_t> stride;
      for (int i : trial) {
        const Node& n = plan.nodes[i];
        const int wt = (int)n.writes[0];
        auto sh = plan.shapes.find(wt);
        if (sh == plan.shapes.end() || sh->second.empty()) { ok = false; break; }
        if (r < 0) r = sh->second[0];
        if (sh->second[0] != r || r <= 1) { ok = false; break; }
        stride[n.ip.write.base] = tensor_len(wt) / r;
      }
      if (ok) {
        for (int i : trial) {
          IpProgram p = plan.nodes[i].ip;
          if (!to_row_form(p, r, stride)) { ok = false; break; }
        }
      }
      if (!ok) break;
      members = trial;
      rows = r;
    }
    if (members.size() >= 6) {
      // the reference's softmax + crossEntropy head and its adjoints, in exactly this wiring?
      const Target& target = *plan.target;
      auto K = [&](int m) -> const Kernel& { return *target.kernels[plan.nodes[members[m]].kernel_index]; };
      auto acc = [&](int m) { return (int)plan.nodes[members[m]].ip.accumulate; };
      bool ok = members.size() == 6;
      for (int m2 = 0; ok && m2 < 6; ++m2) ok = plan.nodes[members[m2]].kernel_index >= 0;
      // The six kernels of softmax (dnn.nim:90-94: no max-subtraction) + crossEntropy (base.nim:66-67) and their
      // derive()d adjoints, matched structurally: expression trees are unified (operand order of commutative
      // operations and read order do not matter), the index tuple of every operand is checked after binding.
      static const KernelForm kHead[6] = {
          {"!.", "[I0]", "exp($0)", {"[I0,I1]"}},                                               // s[y] = sum_x exp(h)
          {"!!", "[I0,I1]", "div(exp($0),$1)", {"[I0,I1]", "[I0]"}},                            // p = exp(h) / s[y]
          {"!", "{I0}", "div(mul(negate(div($2,toscalar(shape($1,0)))),$0),$1)", {"{I0}", "{I0}", "[0]"}},  // dp = -(dL/N) y / p
          {"!!", "[I0,I1]", "mul(div($2,$1),exp($0))", {"[I0,I1]", "[I0]", "[I0,I1]"}},         // dh = dp / s * e
          {"!.", "[I0]", "mul(negate(exp($0)),div($2,mul($1,$1)))", {"[I0,I1]", "[I0]", "[I0,I1]"}},  // ds = -e dp / s^2
          {"!!", "[I0,I1]", "mul($1,exp($0))", {"[I0,I1]", "[I0]"}}};                           // dh += ds e
      PatMatch hm[6];
      int failed_at = -1;
      for (int m2 = 0; ok && m2 < 6; ++m2)
        if (!match_form(K(m2), kHead[m2], hm[m2])) {
          ok = false;
          failed_at = m2;
        }
      ok = ok && acc(0) == 0 && acc(1) == 0 && acc(2) == 0 && acc(3) == 0 && acc(4) == 0 && acc(5) == 1;
      if (!ok && members.size() == 6)
        plan.notes.push_back("row chain of 6 kernels is not the softmax + crossEntropy head" +
                             (failed_at >= 0 ? " (kernel " + std::to_string(plan.nodes[members[failed_at]].kernel_index) + ": " +
                                                   describe_kernel(K(failed_at)) + ")"
                                             : std::string(" (write modes differ)")) +
                             ": runs as a generic row chain");
      if (ok) {
        const int H = hm[0].tensor_of.at(0), S = K(0).write.tensor, P = K(1).write.tensor;
        const int Y = hm[2].tensor_of.at(0), DL = hm[2].tensor_of.at(2), DP = K(2).write.tensor;
        const int DH = K(3).write.tensor, DS = K(4).write.tensor;
        ok = hm[1].tensor_of.at(0) == H && hm[1].tensor_of.at(1) == S && hm[2].tensor_of.at(1) == P &&
             hm[3].tensor_of.at(0) == H && hm[3].tensor_of.at(1) == S && hm[3].tensor_of.at(2) == DP &&
             hm[4].tensor_of.at(0) == H && hm[4].tensor_of.at(1) == S && hm[4].tensor_of.at(2) == DP &&
             hm[5].tensor_of.at(0) == H && hm[5].tensor_of.at(1) == DS && K(5).write.tensor == DH;
        if (!ok) plan.notes.push_back("softmax + crossEntropy kernels found but wired differently: runs as a generic row chain");
        std::set<int> distinct = {H, S, P, Y, DL, DP, DH, DS};
        const auto& hs = plan.shapes.at(H);
        ok = ok && distinct.size() == 8 && hs.size() == 2 && softmax_xent_supported(hs[1]) && plan.shapes.at(Y) == hs &&
             plan.shapes.at(P) == hs && plan.shapes.at(DP) == hs && plan.shapes.at(DH) == hs && tensor_len(DL) == 1 &&
             tensor_len(S) == hs[0] && tensor_len(DS) == hs[0];
        if (ok) {
          Node fx;
          fx.kind = Node::SOFTMAX_XENT;
          fx.label = "softmax + crossEntropy forward/adjoint rows (kernels";
          for (int i : members) {
            fx.label += " " + std::to_string(plan.nodes[i].kernel_index);
            for (auto r : plan.nodes[i].reads) fx.reads.push_back(r);
            for (auto w : plan.nodes[i].writes) fx.writes.push_back(w);
            remove.insert(i);
          }
          fx.label += ")";
          auto ptr = [&](int tensor) {
            for (int i : members) {
              const IpProgram& ip = plan.nodes[i].ip;
              if (plan.nodes[i].writes[0] == tensor) return (float*)ip.write.base;
              const Kernel& kk = *target.kernels[plan.nodes[i].kernel_index];
              for (size_t q = 0; q < kk.reads.size(); ++q)
                if (kk.reads[q].tensor == tensor) return (float*)ip.reads[q].base;
            }
            return (float*)nullptr;
          };
          fx.sx_h = ptr(H); fx.sx_y = ptr(Y); fx.sx_dl = ptr(DL);
          fx.sx_s = ptr(S); fx.sx_p = ptr(P); fx.sx_dp = ptr(DP); fx.sx_dh = ptr(DH); fx.sx_ds = ptr(DS);
          fx.sx_rows = (int)hs[0];
          fx.sx_cols = (int)hs[1];
          // The two nodes that follow the head in a dense net consume DH only: the bias-gradient column sum
          // (db[x] = sum_y dh[y,x]) and the bf16 operand split for the adjoint contractions. Both run inside
          // the row kernel (one level less on the critical path); the column sums meet through atomics, so
          // their tensor is zeroed by a memset node first.
          const int first_member = *std::min_element(members.begin(), members.end());
          for (int j = 0; j < (int)plan.nodes.size(); ++j) {
            if (remove.count(j) || std::find(members.begin(), members.end(), j) != members.end()) continue;
            Node& nj = plan.nodes[j];
            if (j < first_member) continue;
            // nothing between the head and this node may touch what it writes
            bool clean = true;
            for (int k2 = first_member; k2 < j && clean; ++k2) {
              if (remove.count(k2) || std::find(members.begin(), members.end(), k2) != members.end()) continue;
              for (auto w : nj.writes)
                for (auto r : plan.nodes[k2].reads) clean = clean && r != w;
              for (auto w : nj.writes)
                for (auto w2 : plan.nodes[k2].writes) clean = clean && w2 != w;
            }
            if (!clean) continue;
            auto drop = [&](int idx) {   // a fused node must not join a later row chain
              auto& lv = by_level[plan.nodes[idx].level];
              lv.erase(std::remove(lv.begin(), lv.end(), idx), lv.end());
              remove.insert(idx);
            };
            if (nj.kind == Node::SPLIT && nj.split_src == fx.sx_dh && !nj.split_transpose && nj.split_act == 0 && !fx.sx_out_hi &&
                nj.split_rows == fx.sx_rows && nj.split_cols == fx.sx_cols && nj.split_ld == fx.sx_cols) {
              fx.sx_out_hi = nj.split_hi;
              fx.sx_out_mid = nj.split_mid;
              fx.sx_ld_out = nj.split_dst_ld;
              for (auto w : nj.writes) fx.writes.push_back(w);
              fx.label += " + operand planes";
              drop(j);
            } else if (nj.kind == Node::INTERP && nj.kernel_index >= 0 && !fx.sx_colsum && !nj.ip.accumulate &&
                       is_column_sum(*target.kernels[nj.kernel_index]) &&
                       target.kernels[nj.kernel_index]->reads[0].tensor == DH && tensor_len((int)nj.writes[0]) == hs[1]) {
              fx.sx_colsum = (float*)nj.ip.write.base;
              const bool prezeroed = (char*)fx.sx_colsum >= plan.arena && (char*)fx.sx_colsum < plan.arena + plan.zero_bytes;
              if (!prezeroed) {   // (normally it lies in the region the plan's first node zeroes)
                Node z;
                z.kind = Node::MEMSET;
                z.label = "zero column sums of tensor" + std::to_string((int)nj.writes[0] - 1);
                z.ptr = fx.sx_colsum;
                z.bytes = (size_t)hs[1] * 4;
                z.writes.push_back(nj.writes[0]);
                insert.emplace_back(first_member, z);
              }
              for (auto w : nj.writes) fx.writes.push_back(w);
              fx.reads.push_back(nj.writes[0]);
              fx.label += " + column sums (kernel " + std::to_string(nj.kernel_index) + ")";
              drop(j);
            }
          }
          insert.emplace_back(first_member, fx);
          l = end;
          continue;
        }
      }
    }
    if (members.size() >= 2) {
      std::map<uint64_t, int64_t> stride;
      for (int i : members) stride[plan.nodes[i].ip.write.base] = tensor_len(plan.nodes[i].writes[0]) / rows;
      std::vector<IpProgram> progs;
      Node chain;
      chain.kind = Node::ROWCHAIN;
      chain.label = "row chain of " + std::to_string(members.size()) + " kernels:";
      for (int i : members) {
        IpProgram p = plan.nodes[i].ip;
        to_row_form(p, rows, stride);
        progs.push_back(p);
        chain.label += " " + std::to_string(plan.nodes[i].kernel_index);
        for (auto r : plan.nodes[i].reads) chain.reads.push_back(r);
        for (auto w : plan.nodes[i].writes) chain.writes.push_back(w);
        remove.insert(i);
      }
      void* dev = nullptr;
      EGB_CUDA(cudaMalloc(&dev, progs.size() * sizeof(IpProgram)));
      plan.chain_bufs.push_back(dev);
      EGB_CUDA(cudaMemcpyAsync(dev, progs.data(), progs.size() * sizeof(IpProgram), cudaMemcpyHostToDevice, m.ctx->stream));
      EGB_CUDA(cudaStreamSynchronize(m.ctx->stream));  // `progs` is a local
      chain.chain_progs = (IpProgram*)dev;
      chain.chain_n = (int)progs.size();
      for (auto& pr : progs) {
        int used = 0;   // highest slot index referenced + 1 (to_row_form may have added the row iterator)
        for (int q = 0; q < pr.nloops; ++q) used = std::max(used, pr.loops[q].slot + 1);
        chain.chain_slots = std::max(chain.chain_slots, std::max(used, (int)pr.nslots));
      }
      chain.chain_rows = rows;
      insert.emplace_back(*std::min_element(members.begin(), members.end()), chain);
      l = end;
    } else {
      ++l;
    }
  }
  if (insert.empty()) return false;
  std::vector<Node> out;
  for (size_t i = 0; i < plan.nodes.size(); ++i) {
    for (auto& ins : insert)
      if (ins.first == (int)i) out.push_back(ins.second);
    if (!remove.count((int)i)) out.push_back(plan.nodes[i]);
  }
  plan.nodes = out;
  return true;
}

// The fixed map forms of the layer library (pattern.cpp) run on template-specialised streaming kernels.
bool build_eltwise_node(Model& m, Plan& plan, const Kernel& k, int ki, bool overwrite, const std::map<int, void*>& ptrs) {
  if (m.strict || !m.eltwise) return false;
  EltSpec es;
  if (!match_eltwise(k, plan.shapes, es)) return false;
  EltLaunch e;
  e.kind = es.kind;
  e.nreads = es.nreads;
  e.n = es.n;
  e.row = es.row;
  e.accumulate = !overwrite;
  auto ptr = [&](int tensor) {
    auto it = ptrs.find(tensor);
    return it == ptrs.end() ? (float*)nullptr : (float*)it->second;
  };
  e.out = ptr(k.write.tensor);
  for (int q = 0; q < es.nreads && q < 2; ++q) {
    e.in[q] = ptr(es.read_tensor[q]);
    if (es.scalar_read[q] && e.in[q]) e.in[q] += es.scalar_offset[q];   // one fixed element of that tensor
  }
  // literals: constant sub-expressions are folded in float64 and then rounded to the scalar type
  // (propagateConstants passes.nim:1629-1650, llvmgen.nim:213-218); pow(b, epoch) is a run-time fp32 value
  switch (es.kind) {
    case ELT_ADAM_M: case ELT_ADAM_V:
      e.p[0] = (float)(es.lit[0] - 1.0);
      e.p[1] = (float)(1.0 - es.lit[0]);
      break;
    case ELT_ADAM_STEP:
      e.p[0] = (float)(0.0 - es.lit[0]);
      e.p[1] = 1.0f - powf((float)es.lit[1], (float)m.epoch);
      e.p[2] = 1.0f - powf((float)es.lit[2], (float)m.epoch);
      e.p[3] = (float)es.lit[3];
      break;
    default:
      for (int q = 0; q < 4; ++q) e.p[q] = (float)es.lit[q];
  }
  if (!eltwise_stream_supported(e)) return false;
  Node n;
  n.kind = Node::ELTWISE;
  n.label = std::string("eltwise ") + elt_kind_name(es.kind) + " kernel " + std::to_string(ki) + " -> tensor" +
            std::to_string(k.write.tensor - 1) + " n=" + std::to_string(es.n) + (e.accumulate ? " +=" : " =");
  n.elt = e;
  n.kernel_index = ki;
  n.uses_epoch = es.uses_epoch;
  for (auto& r : k.reads) n.reads.push_back(r.tensor);
  n.writes.push_back(k.write.tensor);
  if (e.accumulate) n.reads.push_back(k.write.tensor);
  plan.nodes.push_back(n);
  return true;
}

void build_nodes_impl(Model& m, Plan& plan) {
  const std::vector<KernelInfo>& info = plan.info;
  const Target& target = *plan.target;
  Context& ctx = *m.ctx;
  plan.nodes.clear();
  plan.notes.clear();
  plan.launches_per_run = 0;
  for (void* b : plan.chain_bufs) cudaFree(b);
  plan.chain_bufs.clear();
  std::map<int, void*> ptrs;
  for (int id : target.tensors) ptrs[id] = tensor_ptr(m, plan, id);
  for (auto& kv : plan.shapes)
    if (!ptrs.count(kv.first)) ptrs[kv.first] = tensor_ptr(m, plan, kv.first);

  auto plane_key = [](int tensor, int orientation) { return -(int64_t)(2 * tensor + orientation + 1); };
  if (plan.zero_bytes) {
    Node n;
    n.kind = Node::MEMSET;
    n.label = "zero results";
    n.ptr = plan.arena;
    n.bytes = plan.zero_bytes;
    for (auto& kv : plan.tensors)
      if ((char*)kv.second.ptr < plan.arena + plan.zero_bytes) n.writes.push_back(kv.first);
    plan.nodes.push_back(n);
  }
  for (int id : target.tensors) {
    const TensorDef& td = m.prog->tdef(id);
    if (td.kind == TensorKind::Random) {
      Node n;
      n.kind = Node::RANDOM;
      n.label = "random tensor" + std::to_string(id - 1);
      n.ptr = ptrs[id];
      n.bytes = (size_t)shape_len(plan.shapes.at(id)) * 4;
      n.lo = (float)td.range_lo;
      n.hi = (float)td.range_hi;
      n.tensor = id;
      n.writes.push_back(id);
      plan.nodes.push_back(n);
    }
  }

  // bf16 operand planes are cached per (tensor, orientation) until the tensor is written again
  std::map<std::pair<int, int>, std::pair<__nv_bfloat16*, __nv_bfloat16*>> planes;
  size_t plane_cursor = 0;
  char* plane_base = plan.arena + plan.plane_off;
  const size_t plane_cap = plan.plane_bytes;
  // Planes of a stored [rows, cols] fp32 matrix: row-major as stored (the GEMM descriptors take either
  // operand orientation, so one pair serves every contraction), or - only for large operands whose
  // natural orientation is MN-major, where the K-major tensor-core path is ~7% faster and the extra
  // streaming pass is cheap by comparison - an explicitly transposed copy.
  auto get_planes = [&](int tensor, bool transposed, int64_t rows, int64_t cols, int64_t ld, int64_t& out_ld) {
    const int64_t prow = transposed ? cols : rows, pcol = transposed ? rows : cols;
    out_ld = (pcol + 7) & ~int64_t(7);
    auto key = std::make_pair(tensor, transposed ? 1 : 0);
    auto it = planes.find(key);
    if (it != planes.end()) return it->second;
    const size_t bytes = align_up((size_t)prow * out_ld * 2, 256);
    if (plane_cursor + 2 * bytes > plane_cap) fail(EGB_ERR_RUNTIME, "internal: operand plane arena exhausted");
    auto* hi = (__nv_bfloat16*)(plane_base + plane_cursor);
    auto* mid = (__nv_bfloat16*)(plane_base + plane_cursor + bytes);
    plane_cursor += 2 * bytes;
    Node n;
    n.kind = Node::SPLIT;
    n.label = "split tensor" + std::to_string(tensor - 1) + (transposed ? " (transposed)" : "");
    n.split_src = (const float*)ptrs[tensor];
    n.split_rows = (int)rows;
    n.split_cols = (int)cols;
    n.split_ld = (int)ld;
    n.split_transpose = transposed;
    n.split_hi = hi;
    n.split_mid = mid;
    n.split_dst_ld = (int)out_ld;
    n.reads.push_back(tensor);
    n.writes.push_back(plane_key(tensor, transposed ? 1 : 0));
    plan.nodes.push_back(n);
    planes[key] = std::make_pair(hi, mid);
    return planes[key];
  };

  for (size_t ki = 0; ki < target.kernels.size(); ++ki) {
    const Kernel& k = *target.kernels[ki];
    const KernelInfo& inf = info[ki];
    if (inf.absorbed_by >= 0) continue;  // runs inside the epilogue of its contraction
    if ((int)ki == plan.bucket_before_kernel && plan.bucket_bytes && plan.window.mapped) {
      // Fused kernel(s): gradient exchange through peer memory + the gradientDescent updates it feeds. The bucket is
      // laid out by readiness; when its last tensor (the gradient the step's final contraction produces) is a
      // sizeable part of it, the bucket is exchanged in two launches: everything that is ready earlier travels on a
      // side stream while that contraction is still running (few CTAs: it only has the SMs the contraction leaves
      // free), and only the last gradient's exchange - which then depends on nothing but that contraction and keeps
      // its programmatic edge - is exposed at the end of the step.
      size_t last_off = 0;   // arena offset of the bucket tensor with the highest address
      for (auto& kv : plan.tensors) {
        const size_t off = (size_t)((const char*)kv.second.ptr - plan.arena);
        if (off >= plan.bucket_off && off < plan.bucket_off + plan.bucket_bytes) last_off = std::max(last_off, off);
      }
      const size_t tail_bytes = plan.bucket_off + plan.bucket_bytes - last_off;
      static const bool one_exchange = getenv("EGB_DP_ONE_EXCHANGE") != nullptr;
      const bool two = !one_exchange && last_off > plan.bucket_off && tail_bytes * 4 >= plan.bucket_bytes && plan.window.area_stride > 0;
      for (int part = 0; part < (two ? 2 : 1); ++part) {
        const size_t r_off = two && part == 1 ? last_off : plan.bucket_off;
        const size_t r_bytes = !two ? plan.bucket_bytes : part == 0 ? last_off - plan.bucket_off : tail_bytes;
        Node n;
        n.kind = Node::EXCHANGE;
        ExchangeParams& xp = n.exchange;
        memset((void*)&xp, 0, sizeof(xp));
        xp.rank = plan.window.rank;
        xp.world = plan.window.world;
        xp.n = (long long)(r_bytes / 4);
        static const int early_ctas = getenv("EGB_DP_EARLY_CTAS") ? atoi(getenv("EGB_DP_EARLY_CTAS")) : 128;
        xp.ctas = two && part == 0 ? early_ctas : 0;
        for (int r = 0; r < xp.world; ++r) {
          xp.bucket[r] = (float*)(plan.window.arena[r] + r_off);
          xp.flags[r] = (uint32_t*)((char*)plan.window.flags[r] + (size_t)part * plan.window.area_stride);
        }
        for (size_t q = 0; q < plan.exchange_segs.size() && xp.nseg < EX_MAX_SEG; ++q) {
          ExchangeSeg sg = plan.exchange_segs[q];
          const size_t seg_off = plan.bucket_off + (size_t)sg.off * 4;
          if (seg_off < r_off || seg_off >= r_off + r_bytes) continue;
          sg.off = (long long)((seg_off - r_off) / 4);
          sg.param = (float*)ptrs[plan.exchange_seg_param[q]];
          xp.seg[xp.nseg++] = sg;
          n.reads.push_back(plan.exchange_seg_param[q]);
          n.writes.push_back(plan.exchange_seg_param[q]);
        }
        for (auto& kv : plan.tensors) {
          const char* p0 = (const char*)kv.second.ptr;
          if (p0 >= plan.arena + r_off && p0 < plan.arena + r_off + r_bytes) {
            n.reads.push_back(kv.first);
            n.writes.push_back(kv.first);
          }
        }
        n.label = "peer exchange of the gradient bucket (" + std::to_string(r_bytes) + " bytes" + (two ? (part == 0 ? ", early part" : ", last gradient") : "") +
                  ", " + std::to_string(xp.world) + " ranks) + " + std::to_string(xp.nseg) + " fused gradientDescent updates";
        plan.nodes.push_back(n);
      }
    } else if ((int)ki == plan.bucket_before_kernel && plan.bucket_bytes) {
      for (auto& seg : plan.bucket_segments) {
        Node n;
        n.kind = Node::ALLREDUCE;
        n.label = "all-reduce(avg) gradient bucket segment, " + std::to_string(seg.second) + " bytes";
        n.ptr = plan.arena + seg.first;
        n.bytes = seg.second;
        for (auto& kv : plan.tensors) {
          const char* p0 = (const char*)kv.second.ptr;
          if (p0 >= plan.arena + seg.first && p0 < plan.arena + seg.first + seg.second) {
            n.reads.push_back(kv.first);
            n.writes.push_back(kv.first);
          }
        }
        plan.nodes.push_back(n);
      }
    }
    if (inf.in_exchange && plan.window.mapped) continue;  // its update ran inside the exchange kernel
    if (inf.is_gemm && !m.strict) {
      const GemmPattern& g = inf.gemm;
      Node n;
      n.kind = Node::GEMM;
      n.label = "gemm tensor" + std::to_string(g.c_tensor - 1);
      n.kernel_index = (int)ki;
      int64_t lda = 0, ldb = 0;
      // A stored [M,K] is K-major, stored [K,M] is MN-major; B stored [N,K] is K-major, [K,N] MN-major
      const bool a_copy = g.trans_a && prefer_transposed_copy(g.K, g.M);
      const bool b_copy = !g.trans_b && prefer_transposed_copy(g.K, g.N);
      auto pa = g.trans_a ? get_planes(g.a_tensor, a_copy, g.K, g.M, g.lda, lda)
                          : get_planes(g.a_tensor, false, g.M, g.K, g.lda, lda);
      auto pb = g.trans_b ? get_planes(g.b_tensor, false, g.N, g.K, g.ldb, ldb)
                          : get_planes(g.b_tensor, b_copy, g.K, g.N, g.ldb, ldb);
      n.gemm.a_hi = pa.first; n.gemm.a_mid = pa.second; n.gemm.lda = (int)lda; n.gemm.a_mn = g.trans_a && !a_copy;
      n.gemm.b_hi = pb.first; n.gemm.b_mid = pb.second; n.gemm.ldb = (int)ldb; n.gemm.b_mn = !g.trans_b && !b_copy;
      n.reads.push_back(plane_key(g.a_tensor, a_copy ? 1 : 0));
      n.reads.push_back(plane_key(g.b_tensor, b_copy ? 1 : 0));
      n.writes.push_back(g.c_tensor);
      if (!inf.overwrite) n.reads.push_back(g.c_tensor);
      if (inf.bias_tensor) n.reads.push_back(inf.bias_tensor);
      if (inf.h_tensor) n.reads.push_back(inf.h_tensor);
      if (inf.d_tensor) {
        n.writes.push_back(inf.d_tensor);
        if (inf.epi == EPI_SGD) n.reads.push_back(inf.d_tensor);
      }
      if (inf.colsum_tensor) {
        n.writes.push_back(inf.colsum_tensor);
        n.reads.push_back(inf.colsum_tensor);
      }
      if (inf.emit_planes) n.writes.push_back(plane_key(inf.final_tensor, 0));
      n.gemm.M = (int)g.M; n.gemm.N = (int)g.N; n.gemm.K = (int)g.K;
      n.gemm.C = (float*)ptrs[g.c_tensor];
      n.gemm.ldc = (int)g.ldc;
      n.gemm.flags = inf.overwrite ? 0 : GEMM_ACCUMULATE;
      n.gemm.alpha = 1.0f;
      if (!m.splitk) n.gemm.cluster_k = 1;  // option: no cluster split-K
      if (inf.bias_tensor) {
        n.gemm.flags |= GEMM_BIAS;
        n.gemm.bias = (const float*)ptrs[inf.bias_tensor];
      }
      n.gemm.epi = inf.epi;
      n.gemm.epi_param = inf.epi_param;
      if (inf.d_tensor) n.gemm.D = (float*)ptrs[inf.d_tensor];
      if (inf.h_tensor) n.gemm.H = (const float*)ptrs[inf.h_tensor];
      if (inf.colsum_tensor) n.gemm.colsum = (float*)ptrs[inf.colsum_tensor];
      if (!inf.absorbed.empty()) n.label += " +" + std::to_string(inf.absorbed.size()) + " fused";
      if (inf.emit_planes) {
        // the epilogue writes the operand planes of its final value: later contractions find them in the cache
        const int64_t cols = g.N, out_ld = (cols + 7) & ~int64_t(7);
        const size_t bytes = align_up((size_t)g.M * out_ld * 2, 256);
        if (plane_cursor + 2 * bytes > plane_cap) fail(EGB_ERR_RUNTIME, "internal: operand plane arena exhausted");
        n.gemm.out_hi = (__nv_bfloat16*)(plane_base + plane_cursor);
        n.gemm.out_mid = (__nv_bfloat16*)(plane_base + plane_cursor + bytes);
        n.gemm.ld_out = (int)out_ld;
        n.gemm.flags |= GEMM_SPLIT_OUT;
        plane_cursor += 2 * bytes;
      }
      if (gemm_2cta_eligible(n.gemm)) {
        // workspace of the tail-wave split (gemm_tcgen05_2cta.cu): zeroed with the arena, flags clean themselves
        const size_t wsb = align_up(gemm_2cta_workspace_bytes(ctx.sm_count), 256);
        if (plane_cursor + wsb <= plane_cap) {
          n.gemm.ws = plane_base + plane_cursor;
          n.gemm.ws_bytes = wsb;
          plane_cursor += wsb;
        }
      }
      plan.nodes.push_back(n);
      if (inf.emit_planes) planes[std::make_pair(inf.final_tensor, 0)] = std::make_pair(n.gemm.out_hi, n.gemm.out_mid);
    } else if (inf.is_conv && !m.strict) {
      const ConvPattern& cv = inf.conv;
      Node n;
      n.kind = Node::CONV;
      n.conv = cv;
      n.kernel_index = (int)ki;
      n.conv_accumulate = !inf.overwrite;
      if (cv.kind == ConvPattern::FORWARD) {
        n.label = "conv2 forward -> tensor" + std::to_string(cv.out_tensor - 1);
        n.conv_a = (const float*)ptrs[cv.img_tensor];
        n.conv_b = (const float*)ptrs[cv.fil_tensor];
        n.conv_out = (float*)ptrs[cv.out_tensor];
      } else if (cv.kind == ConvPattern::D_FILTERS) {
        n.label = "conv2 d_filters -> tensor" + std::to_string(cv.fil_tensor - 1);
        n.conv_a = (const float*)ptrs[cv.img_tensor];
        n.conv_b = (const float*)ptrs[cv.out_tensor];
        n.conv_out = (float*)ptrs[cv.fil_tensor];
      } else {
        n.label = "conv2 d_images -> tensor" + std::to_string(cv.img_tensor - 1);
        n.conv_a = (const float*)ptrs[cv.out_tensor];
        n.conv_b = (const float*)ptrs[cv.fil_tensor];
        n.conv_out = (float*)ptrs[cv.img_tensor];
      }
      for (auto& r : k.reads) n.reads.push_back(r.tensor);
      n.writes.push_back(k.write.tensor);
      n.reads.push_back(k.write.tensor);
      plan.nodes.push_back(n);
    } else if (build_eltwise_node(m, plan, k, (int)ki, inf.overwrite, ptrs)) {
      // one of the fixed elementwise / optimizer forms: specialised streaming kernel (no interpreter)
    } else {
      Lowered lw = lower_kernel(k, plan.shapes, ptrs, m.epoch, m.strict, inf.overwrite, ctx.sm_count);
      Node n;
      n.kind = Node::INTERP;
      n.label = "kernel " + std::to_string(ki) + " -> tensor" + std::to_string(k.write.tensor - 1);
      n.ip = lw.ip;
      n.pb = lw.pb;
      n.rb = lw.rb;
      n.points_fast = lw.points_fast;
      n.strict = m.strict;
      n.uses_epoch = lw.uses_epoch;
      n.kernel_index = (int)ki;
      for (auto& r : k.reads) n.reads.push_back(r.tensor);
      n.writes.push_back(k.write.tensor);
      n.rsplit = interp_reduction_splits(n.ip, n.pb, n.rb, m.strict, ctx.sm_count);
      if (n.rsplit > 1 && !n.ip.accumulate) {
        // the splits accumulate into the output, so an overwriting kernel first gets its (completely
        // covered) output tensor cleared
        Node z;
        z.kind = Node::MEMSET;
        z.label = "zero tensor" + std::to_string(k.write.tensor - 1) + " (reduction combined with atomics)";
        z.ptr = ptrs[k.write.tensor];
        z.bytes = (size_t)shape_len(plan.shapes.at(k.write.tensor)) * 4;
        z.writes.push_back(k.write.tensor);
        plan.nodes.push_back(z);
        n.ip.accumulate = 1;
      }
      if (n.ip.accumulate) n.reads.push_back(k.write.tensor);
      // A kernel that reads nothing and overwrites a result no other kernel of the target writes (the adjoint
      // seed `dL = 1`, passes.nim:614-636) produces the same values on every run: it is launched once, here,
      // and stays out of the plan - in the dense step it sat on the critical path in front of the head.
      bool constant = m.fuse && !m.strict && k.reads.empty() && !n.uses_epoch && !n.ip.accumulate && n.rsplit <= 1 &&
                      m.prog->tdef(k.write.tensor).kind == TensorKind::Result;
      if (constant) {
        int writers = 0;
        for (auto& other : target.kernels) writers += other->write.tensor == k.write.tensor;
        constant = writers == 1;
      }
      if (constant) {
        launch_interp(ctx, n.ip, n.pb, n.rb, n.points_fast, n.strict, ctx.stream, n.rsplit);
      } else {
        plan.nodes.push_back(n);
      }
    }
    // cached planes of every tensor this unit wrote are stale now (except the ones it just produced)
    std::vector<int> wrote = {k.write.tensor};
    for (int kj : inf.absorbed) wrote.push_back(target.kernels[kj]->write.tensor);
    for (int wtensor : wrote) {
      if (inf.emit_planes && wtensor == inf.final_tensor && inf.is_gemm && !m.strict) continue;
      for (auto it = planes.begin(); it != planes.end();)
        it = it->first.first == wtensor ? planes.erase(it) : std::next(it);
    }
  }
  if (m.fuse && m.eltwise && !m.strict) fuse_adam(plan);
  compute_levels(plan);
  // level order is a topological order: sorting by it makes every run of consecutive levels contiguous
  // (needed by the row-chain fusion) and keeps sequential (eager) execution valid
  std::stable_sort(plan.nodes.begin(), plan.nodes.end(), [](const Node& a, const Node& b) { return a.level < b.level; });
  // The operand splits that depend on nothing inside the plan (input batch, parameters) run as ONE launch:
  // they sit in front of the first contraction on the critical path.
  if (m.fuse) {
    std::vector<int> roots;
    for (int i = 0; i < (int)plan.nodes.size(); ++i) {
      const Node& n = plan.nodes[i];
      if (n.kind == Node::SPLIT && n.level == 0 && !n.split_transpose && (int)roots.size() < SplitBatch::MAX_JOBS) roots.push_back(i);
    }
    if (roots.size() >= 2) {
      Node& first = plan.nodes[roots[0]];
      std::string merged_label = "split";
      for (int i : roots) {
        const Node& n = plan.nodes[i];
        first.split_jobs.push_back(SplitJob{n.split_src, n.split_hi, n.split_mid, n.split_rows, n.split_cols, n.split_ld,
                                            n.split_dst_ld, n.split_act});
        const size_t tpos = n.label.find("tensor");
        merged_label += " " + (tpos == std::string::npos ? n.label : n.label.substr(tpos));
        if (i != roots[0]) {
          for (auto r : n.reads) first.reads.push_back(r);
          for (auto w : n.writes) first.writes.push_back(w);
        }
      }
      first.label = merged_label + " (one launch)";
      // the plan's leading memset (zero-initialised results) rides along: one node less in front of the first
      // contraction. Its region is 16-byte aligned by construction (arena base, 256-byte tensor alignment).
      int zero_node = -1;
      for (int i = 0; i < (int)plan.nodes.size(); ++i) {
        const Node& n = plan.nodes[i];
        if (n.kind == Node::MEMSET && n.level == 0 && n.ptr == plan.arena && n.bytes == plan.zero_bytes && n.bytes % 16 == 0) zero_node = i;
      }
      if (zero_node >= 0) {
        first.ptr = plan.nodes[zero_node].ptr;
        first.bytes = plan.nodes[zero_node].bytes;
        for (auto w : plan.nodes[zero_node].writes) first.writes.push_back(w);
        first.label += " + zero results";
      }
      std::vector<Node> kept;
      for (int i = 0; i < (int)plan.nodes.size(); ++i)
        if (i != zero_node && (i == roots[0] || std::find(roots.begin(), roots.end(), i) == roots.end())) kept.push_back(plan.nodes[i]);
      plan.nodes = kept;
      compute_levels(plan);
    }
  }
  if (m.rowchain && !m.strict) {
    if (fuse_row_chains(m, plan)) compute_levels(plan);
  }
  // Dead-store elimination: the fp32 form of a contraction's C / D tensor that only its own epilogue consumes
  // (fused stages, column sums, operand planes) and no later node reads is not written to memory - the same
  // reasoning as the reference's deadKernelElim (passes.nim:331-350), applied to stores. The epilogues of the
  // dense step are bound by L2 write bandwidth (C + D + two planes = 6 MB per contraction); this removes a third
  // to two thirds of it. Option keep_intermediates=1 materialises everything (egb_model_read_tensor on such a
  // tensor otherwise fails loudly).
  plan.unmaterialized.clear();
  if (m.fuse && !m.strict && !m.keep_intermediates) {
    for (size_t i = 0; i < plan.nodes.size(); ++i) {
      Node& n = plan.nodes[i];
      if (n.kind != Node::GEMM || n.kernel_index < 0 || (n.gemm.flags & GEMM_ACCUMULATE)) continue;
      const KernelInfo& inf = info[(size_t)n.kernel_index];
      auto dead = [&](int t) {
        if (t <= 0 || t == target.output || m.prog->tdef(t).kind != TensorKind::Result) return false;
        for (size_t j = i + 1; j < plan.nodes.size(); ++j)
          for (auto r : plan.nodes[j].reads)
            if (r == t) return false;
        return true;
      };
      const int c_t = inf.gemm.c_tensor, d_t = inf.d_tensor;
      // C is needed in memory unless some fused consumer exists at all (else the contraction itself would be dead)
      const bool has_consumer = inf.epi != EPI_NONE || inf.colsum_tensor || inf.emit_planes;
      if (has_consumer && dead(c_t) && (inf.epi != EPI_NONE || inf.final_tensor == c_t)) {
        n.gemm.flags |= GEMM_SKIP_C;
        plan.unmaterialized.insert(c_t);
      }
      if (d_t && inf.epi != EPI_NONE && inf.epi != EPI_SGD && (inf.colsum_tensor || inf.emit_planes) && dead(d_t)) {
        n.gemm.flags |= GEMM_SKIP_D;
        plan.unmaterialized.insert(d_t);
      }
      if (n.gemm.flags & (GEMM_SKIP_C | GEMM_SKIP_D))
        n.label += std::string(" [fp32") + ((n.gemm.flags & GEMM_SKIP_C) ? " C" : "") + ((n.gemm.flags & GEMM_SKIP_D) ? " D" : "") +
                   " not stored]";
    }
  }
  if (m.fuse && m.headfuse && !m.strict && fuse_head(m, plan)) compute_levels(plan);
  // Contractions of one level run concurrently (parallel graph branches). Each needs a whole SM per CTA
  // (shared memory), so grids that together exceed the machine run in two waves and the level takes twice
  // as long (timeline: 256 CTAs at level 7 of the dense step). Give every contraction of a level an SM
  // budget proportional to its work; its tile width / cluster split-K factor is then chosen inside it.
  if (m.concurrent) {
    std::map<int, double> level_work;
    std::map<int, int> level_gemms;
    for (auto& n : plan.nodes)
      if (n.kind == Node::GEMM) {
        level_work[n.level] += (double)n.gemm.M * n.gemm.N * n.gemm.K;
        level_gemms[n.level]++;
      }
    // The contraction chain that bounds the step (longest path by the capture's cost model) must never wait for
    // SMs: a side contraction of level L is still running when the chain's contraction of level L + 1 wants to
    // start (timeline of the dense step: the 64 CTAs of the layer-2 weight gradient kept the 112-CTA layer-1
    // weight gradient waiting for 3 us), so it only gets the SMs that one leaves free.
    std::vector<char> chain(plan.nodes.size(), 0);
    {
      const int nn = (int)plan.nodes.size();
      auto hit = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
        for (auto x : a)
          for (auto y : b)
            if (x == y) return true;
        return false;
      };
      std::vector<double> longest((size_t)nn, 0.0);
      std::vector<int> via((size_t)nn, -1);
      int end = -1;
      for (int i = 0; i < nn; ++i) {
        const Node& nd = plan.nodes[(size_t)i];
        const double cost = nd.kind == Node::GEMM ? 8.0 + 3e-9 * (double)nd.gemm.M * nd.gemm.N * nd.gemm.K : nd.kind == Node::EXCHANGE ? 12.0 : 4.0;
        longest[(size_t)i] = cost;
        for (int j = 0; j < i; ++j) {
          const Node& pj = plan.nodes[(size_t)j];
          if ((hit(pj.writes, nd.reads) || hit(pj.writes, nd.writes) || hit(pj.reads, nd.writes)) && longest[(size_t)j] + cost > longest[(size_t)i]) {
            longest[(size_t)i] = longest[(size_t)j] + cost;
            via[(size_t)i] = j;
          }
        }
        if (end < 0 || longest[(size_t)i] > longest[(size_t)end]) end = i;
      }
      for (int i = end; i >= 0; i = via[(size_t)i]) chain[(size_t)i] = 1;
    }
    std::map<int, int> chain_need;   // level -> CTAs of the chain's contraction at that level (planned for the whole machine)
    for (size_t i = 0; i < plan.nodes.size(); ++i) {
      const Node& n = plan.nodes[i];
      if (n.kind == Node::GEMM && chain[i] && level_gemms[n.level] == 1) chain_need[n.level] = gemm_planned_ctas(n.gemm, 0, m.ctx->sm_count);
    }
    for (size_t i = 0; i < plan.nodes.size(); ++i) {
      Node& n = plan.nodes[i];
      if (n.kind == Node::GEMM && level_gemms[n.level] > 1) {
        const double share = (double)n.gemm.M * n.gemm.N * n.gemm.K / level_work[n.level];
        int budget = (int)(m.ctx->sm_count * share) / 8 * 8;
        auto next = chain_need.find(n.level + 1);
        if (!chain[i] && next != chain_need.end()) budget = std::min(budget, (m.ctx->sm_count - next->second) / 8 * 8);
        if (budget < 16) budget = 16;
        n.gemm.sm_budget = budget;
        n.label += " sm<=" + std::to_string(budget);
      }
    }
  }
  for (auto& n : plan.nodes)
    if (n.kind != Node::MEMSET && n.kind != Node::ALLREDUCE) plan.launches_per_run++;
  plan.epoch_built = m.epoch;
  plan.graph_valid = false;
}

}  // namespace

void Model::build_nodes(Plan& plan) { build_nodes_impl(*this, plan); }

void Model::evict_plan(size_t index) {
  if (index >= plans.size()) return;
  if (ctx) cudaStreamSynchronize(ctx->stream);  // its graph / arena may still be in flight
  if (last_plan == plans[index].get()) last_plan = nullptr;
  plans.erase(plans.begin() + (long)index);
}

Plan& Model::get_plan(const std::string& target_name, const std::vector<int>& ids,
                      const std::vector<std::vector<int64_t>>& in_shapes) {
  Target* target = prog->find_target(target_name);
  if (!target) fail(EGB_ERR_RUNTIME, "%s is not a target of the model", target_name.c_str());
  std::vector<std::pair<int, std::vector<int64_t>>> sig;
  for (size_t i = 0; i < ids.size(); ++i) sig.emplace_back(ids[i], in_shapes[i]);
  std::sort(sig.begin(), sig.end());
  for (auto& p : plans)
    if (p->target_name == target_name && p->input_sig == sig) {
      p->last_used = ++use_clock;
      return *p;
    }
  // bound the cache: drop the least recently used plan(s) of this target
  for (;;) {
    int count = 0, lru = -1;
    for (size_t i = 0; i < plans.size(); ++i)
      if (plans[i]->target_name == target_name) {
        ++count;
        if (lru < 0 || plans[i]->last_used < plans[lru]->last_used) lru = (int)i;
      }
    if (count < std::max(1, max_plans_per_target) || lru < 0) break;
    evict_plan((size_t)lru);
  }

  // ---- run-time shape inference (passes.nim:1386-1436), once per input-shape signature
  ShapeTable inputs;
  for (auto& s : sig) inputs[s.first] = s.second;
  auto plan = std::make_unique<Plan>();
  plan->target_name = target_name;
  plan->target = target;
  plan->input_sig = sig;
  plan->shapes = infer_shapes(*prog, *target, inputs);
  for (auto& kv : state) plan->shapes[kv.first] = kv.second.shape;

  // ---- classify kernels, decide overwrite vs accumulate and which results need zero-filling
  const size_t nk = target->kernels.size();
  plan->info.resize(nk);
  std::set<int> written, read_first, needs_zero;
  size_t plane_bytes = 0;
  // (EGB_DP_FORCE: lay the plan out for data parallelism on a single rank - timing studies of the DP plan)
  const bool dp_force = getenv("EGB_DP_FORCE") != nullptr;
  const bool dp = comm && (comm_world(comm) > 1 || dp_force);
  for (size_t ki = 0; ki < nk; ++ki)
    if (target->kernels[ki]->is_generator())
      fail(EGB_ERR_GENERATOR, "program still contains generator kernels; compile it first");
  auto same_shape = [&](int a, int b) {
    auto x = plan->shapes.find(a), y = plan->shapes.find(b);
    return x != plan->shapes.end() && y != plan->shapes.end() && x->second == y->second;
  };
  auto is_fresh = [&](int t) {
    return prog->tdef(t).kind == TensorKind::Result && !written.count(t) && !read_first.count(t) && t != 0;
  };
  for (size_t ki = 0; ki < nk; ++ki) {
    const Kernel& k = *target->kernels[ki];
    KernelInfo& inf = plan->info[ki];
    if (inf.absorbed_by >= 0) continue;  // accounted for at the position of its contraction
    for (auto& r : k.reads) {
      if (!plan->shapes.count(r.tensor)) fail(EGB_ERR_SHAPE, "Missing shape for tensor%d", r.tensor - 1);
      if (!written.count(r.tensor)) read_first.insert(r.tensor);
    }
    inf.is_gemm = !strict && match_gemm(k, plan->shapes, inf.gemm);
    inf.is_conv = !strict && !inf.is_gemm && match_conv2(k, plan->shapes, inf.conv) &&
                  (inf.conv.kind != ConvPattern::D_IMAGES || conv2_dimg_supported(inf.conv.KW));
    const int wt = k.write.tensor;
    const bool fresh = is_fresh(wt);
    bool reads_self = false;
    for (auto& r : k.reads) reads_self = reads_self || r.tensor == wt;
    // a convolution (forward / d_images) produces every element of its output; d_filters accumulates
    // with atomics and therefore always needs a zeroed destination
    const bool conv_covers = inf.is_conv && inf.conv.kind != ConvPattern::D_FILTERS;
    inf.overwrite = fresh && !reads_self && (inf.is_gemm || conv_covers ||
                                             (!inf.is_conv && covers_whole_tensor(k, plan->shapes)));
    if (prog->tdef(wt).kind == TensorKind::Result && !written.count(wt) && !inf.overwrite) needs_zero.insert(wt);
    // a plain column sum (bias gradient) may be fused into the classification head, where the partial sums meet
    // through atomics: keep its (small) result in the zeroed region so that no extra memset node is needed
    if (fuse && !strict && prog->tdef(wt).kind == TensorKind::Result && !written.count(wt) && inf.overwrite && !inf.is_gemm &&
        !inf.is_conv && is_column_sum(k))
      needs_zero.insert(wt);
    written.insert(wt);
    if (inf.is_gemm) {
      const GemmPattern& g = inf.gemm;
      auto pad8 = [](int64_t v) { return (size_t)((v + 7) & ~int64_t(7)); };
      // upper bound over both possible orientations of each operand
      const size_t a_bytes = std::max((size_t)g.K * pad8(g.M), (size_t)g.M * pad8(g.K)) * 2;
      const size_t b_bytes = std::max((size_t)g.N * pad8(g.K), (size_t)g.K * pad8(g.N)) * 2;
      plane_bytes += 2 * align_up(a_bytes, 256) + 2 * align_up(b_bytes, 256);
      if (g.M >= 512 && g.N >= 256 && g.K >= 256) plane_bytes += align_up(gemm_2cta_workspace_bytes(ctx->sm_count), 256);  // tail-wave split
      const bool b_copy = !g.trans_b && prefer_transposed_copy(g.K, g.N);
      gemm_choose_config((int)g.M, (int)g.N, (int)g.K, !g.trans_b && !b_copy, ctx->sm_count, &inf.bn, &inf.tiles);
    }
    inf.final_tensor = wt;
    if (!(inf.is_gemm && inf.overwrite && fuse)) continue;

    // ---- epilogue fusion: look ahead for kernels that post-process this contraction's output
    const int C = wt;
    std::set<int> touched_r, touched_w;  // tensors used by kernels that stay between ki and the candidate
    bool has_bias = false, has_stage = false, has_colsum = false;
    for (size_t kj = ki + 1; kj < nk; ++kj) {
      const Kernel& c = *target->kernels[kj];
      KernelInfo& cinf = plan->info[kj];
      if (cinf.absorbed_by >= 0) continue;  // already runs earlier, inside another contraction
      bool conflict = touched_r.count(c.write.tensor) || touched_w.count(c.write.tensor);
      for (auto& r : c.reads) conflict = conflict || touched_w.count(r.tensor);
      bool take = false;
      const int final_t = inf.final_tensor;
      // Structural classification of the candidate (pattern.cpp): which map form it computes and which tensors
      // its operands are - independent of operand order, register numbering and read order.
      EltSpec es;
      const bool is_map = !conflict && match_eltwise(c, plan->shapes, es);
      if (is_map) {
        const int r0 = es.read_tensor[0], r1 = es.read_tensor[1];
        if (!has_bias && !has_stage && !has_colsum && es.kind == ELT_BIAS_ROW && c.write.tensor == C && r0 != C) {
          inf.bias_tensor = r0;                                             // h[y,x] += b[x]      (dnn.nim:22-24)
          has_bias = take = true;
        } else if (!has_stage && !has_colsum && r0 == C && is_fresh(c.write.tensor) && same_shape(c.write.tensor, C) &&
                   (es.kind == ELT_RELU || es.kind == ELT_LEAKY || es.kind == ELT_SIGMOID || es.kind == ELT_TANH)) {
          inf.epi = es.kind == ELT_RELU ? EPI_RELU : es.kind == ELT_LEAKY ? EPI_LEAKY : es.kind == ELT_SIGMOID ? EPI_SIGMOID : EPI_TANH;
          inf.epi_param = (float)es.lit[0];                                 // activations        (dnn.nim:26-40)
          inf.d_tensor = c.write.tensor;
          has_stage = take = true;
        } else if (!has_stage && !has_colsum && (es.kind == ELT_RELU_ADJ || es.kind == ELT_LEAKY_ADJ) && r1 == C && r0 != C &&
                   is_fresh(c.write.tensor) && same_shape(c.write.tensor, C) && same_shape(r0, C) && !touched_w.count(r0)) {
          inf.epi = es.kind == ELT_LEAKY_ADJ ? EPI_MASK_LEAKY : EPI_MASK_RELU;   // adjoint masks (derive of select)
          inf.epi_param = (float)es.lit[0];
          inf.d_tensor = c.write.tensor;
          inf.h_tensor = r0;
          has_stage = take = true;
        } else if (!dp && !has_stage && !has_colsum && es.kind == ELT_SCALE_NEG && r0 == C &&
                   prog->tdef(c.write.tensor).kind == TensorKind::Param && same_shape(c.write.tensor, C)) {
          inf.epi = EPI_SGD;                                                // P += (0 - g) * rate (base.nim:37-38)
          inf.epi_param = (float)es.lit[0];
          inf.d_tensor = c.write.tensor;
          has_stage = take = true;
        }
      } else if (!conflict && !has_colsum && is_column_sum(c) && c.reads[0].tensor == final_t && is_fresh(c.write.tensor) &&
                 kernel_loops_full(c, plan->shapes)) {
        inf.colsum_tensor = c.write.tensor;                                 // db[x] = sum_y d[y,x]
        has_colsum = take = true;
      }
      if (take) {
        cinf.absorbed_by = (int)ki;
        inf.absorbed.push_back((int)kj);
        if (inf.epi != EPI_SGD && inf.d_tensor) inf.final_tensor = inf.d_tensor;
        if (inf.colsum_tensor == c.write.tensor) needs_zero.insert(c.write.tensor);  // accumulated atomically
        written.insert(c.write.tensor);
        continue;
      }
      touched_w.insert(c.write.tensor);
      for (auto& r : c.reads) touched_r.insert(r.tensor);
    }
  }
  // does a later contraction consume a fused epilogue's final value as an operand? then emit its planes
  for (size_t ki = 0; ki < nk; ++ki) {
    KernelInfo& inf = plan->info[ki];
    if (!inf.is_gemm || inf.absorbed_by >= 0 || !inf.overwrite || inf.epi == EPI_SGD) continue;
    for (size_t kj = ki + 1; kj < nk && !inf.emit_planes; ++kj) {
      const KernelInfo& c = plan->info[kj];
      if (c.absorbed_by >= 0) continue;
      if (c.is_gemm) {
        const GemmPattern& g = c.gemm;
        const bool a_copy = g.trans_a && prefer_transposed_copy(g.K, g.M);
        const bool b_copy = !g.trans_b && prefer_transposed_copy(g.K, g.N);
        if ((g.a_tensor == inf.final_tensor && !a_copy) || (g.b_tensor == inf.final_tensor && !b_copy)) inf.emit_planes = true;
      }
      if (target->kernels[kj]->write.tensor == inf.final_tensor) break;  // rewritten before any use
    }
    if (inf.emit_planes) {
      const auto& sh = plan->shapes.at(inf.final_tensor);
      plane_bytes += 2 * align_up((size_t)sh[0] * (size_t)((sh[1] + 7) & ~int64_t(7)) * 2, 256);
    }
  }
  for (int id : target->tensors)
    if (prog->tdef(id).kind == TensorKind::Result && !written.count(id)) needs_zero.insert(id);  // read-only results

  // ---- data parallel: the gradients of the parameters form one contiguous bucket that is
  // all-reduced right before the first optimizer kernel (the first kernel that writes a param/cache)
  std::set<int> bucket;
  if (dp) {
    // first optimizer kernel = the first kernel that writes a parameter or a cache (parser.nim:757-766)
    for (size_t ki = 0; ki < nk; ++ki) {
      const TensorKind wk = prog->tdef(target->kernels[ki]->write.tensor).kind;
      if (wk == TensorKind::Param || wk == TensorKind::Cache) {
        plan->bucket_before_kernel = (int)ki;
        break;
      }
    }
    if (plan->bucket_before_kernel >= 0) {
      std::set<int> before;   // tensors written by the kernels in front of the optimizer block
      for (int ki = 0; ki < plan->bucket_before_kernel; ++ki) before.insert(target->kernels[(size_t)ki]->write.tensor);
      auto gt = prog->grad_tensors.find(target_name);
      if (gt != prog->grad_tensors.end()) {
        // the gradient table `generate` recorded (it travels with the serialised program)
        for (int pid : prog->params) {
          auto g = gt->second.find(pid);
          if (g != gt->second.end() && written.count(g->second)) bucket.insert(g->second);
        }
      } else {
        // a program compiled elsewhere (passes.nim) carries no table: the gradients are the result tensors of the
        // backward block that the optimizer kernels read
        for (size_t ki = (size_t)plan->bucket_before_kernel; ki < nk; ++ki) {
          const TensorKind wk = prog->tdef(target->kernels[ki]->write.tensor).kind;
          if (wk != TensorKind::Param && wk != TensorKind::Cache) continue;
          for (auto& r : target->kernels[ki]->reads)
            if (prog->tdef(r.tensor).kind == TensorKind::Result && before.count(r.tensor)) bucket.insert(r.tensor);
        }
      }
      if (bucket.empty())
        fail(EGB_ERR_RUNTIME,
             "data parallel: target %s updates parameters but no parameter-gradient tensors were found - the replicas "
             "would train on their own shards without averaging", target_name.c_str());
      // every gradient must be complete when the exchange runs (it sits in front of the first optimizer kernel)
      for (size_t ki = (size_t)plan->bucket_before_kernel; ki < nk; ++ki)
        if (bucket.count(target->kernels[ki]->write.tensor))
          fail(EGB_ERR_RUNTIME,
               "data parallel: gradient tensor%d is still written (kernel %zu) after the first optimizer kernel (%d) of "
               "target %s; split the target so that every backward pass precedes its optimizer",
               target->kernels[ki]->write.tensor - 1, ki, plan->bucket_before_kernel, target_name.c_str());
    }
  }
  const bool use_peer = dp && !bucket.empty() && dp_peer && comm_world(comm) <= EX_MAX_WORLD;
  // The bucket is laid out in the order in which the gradients become ready (position of the unit that
  // writes them last) and cut into segments of >= 256 KB: every segment is one all-reduce that can
  // start as soon as its gradients exist and overlaps the adjoint kernels of the layers below it.
  std::vector<int> bucket_order(bucket.begin(), bucket.end());
  {
    std::map<int, int> ready;
    for (size_t ki = 0; ki < nk; ++ki) {
      const int pos = plan->info[ki].absorbed_by >= 0 ? plan->info[ki].absorbed_by : (int)ki;
      const int wt = target->kernels[ki]->write.tensor;
      if (bucket.count(wt)) ready[wt] = std::max(ready.count(wt) ? ready[wt] : -1, pos);
    }
    std::stable_sort(bucket_order.begin(), bucket_order.end(), [&](int a, int b) { return ready[a] < ready[b]; });
    for (int id : bucket_order) needs_zero.insert(id);  // keeps the bucket contiguous inside the zeroed region
  }

  // ---- arena: [results that need zeroing .. | bucket (zeroed part first) | other results, inputs,
  //              random][bf16 operand planes]
  size_t cursor = 0;
  std::vector<int> order;
  auto in_arena = [&](int id) {
    const TensorKind kind = prog->tdef(id).kind;
    return kind == TensorKind::Result || kind == TensorKind::Input || kind == TensorKind::Random;
  };
  for (int id : target->tensors)
    if (needs_zero.count(id) && !bucket.count(id)) order.push_back(id);
  const size_t bucket_first = order.size();
  for (int id : bucket_order) order.push_back(id);
  const size_t n_zero = order.size();
  const size_t bucket_last = order.size();
  for (int id : target->tensors)
    if (!needs_zero.count(id) && !bucket.count(id) && in_arena(id)) order.push_back(id);
  std::map<int, size_t> offs;
  for (size_t i = 0; i < order.size(); ++i) {
    const int id = order[i];
    auto sh = plan->shapes.find(id);
    if (sh == plan->shapes.end()) fail(EGB_ERR_SHAPE, "Missing shape for tensor%d", id - 1);
    for (auto d : sh->second)
      if (d < 0) fail(EGB_ERR_SHAPE, "tensor%d has a negative dimension %s", id - 1, shape_text(sh->second).c_str());
    if (i == bucket_first) plan->bucket_off = cursor;
    offs[id] = cursor;
    cursor += align_up((size_t)shape_len(sh->second) * 4, 256);
    if (i + 1 == n_zero) plan->zero_bytes = cursor;
    if (i + 1 == bucket_last) plan->bucket_bytes = cursor - plan->bucket_off;
    if (i >= bucket_first && i < bucket_last) {
      // close a segment once it holds >= 256 KB (the last one takes the remainder)
      if (plan->bucket_segments.empty() || plan->bucket_segments.back().second >= (256u << 10))
        plan->bucket_segments.emplace_back(offs[id], 0);
      plan->bucket_segments.back().second = cursor - plan->bucket_segments.back().first;
    }
  }
  if (n_zero == 0) plan->zero_bytes = 0;
  if (bucket.empty()) plan->bucket_bytes = 0;
  plan->plane_off = cursor;
  plan->plane_bytes = plane_bytes;
  cursor += plane_bytes;
  plan->arena_bytes = std::max<size_t>(cursor, 256);
  cudaError_t e = cudaMalloc((void**)&plan->arena, plan->arena_bytes);
  if (e != cudaSuccess && !plans.empty()) {
    cudaGetLastError();  // out of memory: give back every cached plan and try once more
    while (!plans.empty()) evict_plan(plans.size() - 1);
    e = cudaMalloc((void**)&plan->arena, plan->arena_bytes);
  }
  if (e != cudaSuccess)
    fail(EGB_ERR_GPU, "cudaMalloc of %zu bytes for target %s failed: %s", plan->arena_bytes, target_name.c_str(),
         cudaGetErrorString(e));
  EGB_CUDA(cudaMemsetAsync(plan->arena, 0, plan->arena_bytes, ctx->stream));  // deterministic padding
  for (int id : order) {
    DevTensor t;
    t.ptr = plan->arena + offs[id];
    t.shape = plan->shapes[id];
    t.bytes = (size_t)shape_len(t.shape) * 4;
    plan->tensors[id] = t;
  }
  if (use_peer) {
    // gradientDescent updates (P += (0 - g) * rate, base.nim:37-38) whose gradient lives in the bucket run inside
    // the exchange kernel. Hoisting one to the exchange position is valid when nothing between that position and
    // the kernel touches the parameter or rewrites the gradient.
    for (size_t ki = (size_t)plan->bucket_before_kernel; ki < nk && (int)plan->exchange_segs.size() < EX_MAX_SEG; ++ki) {
      const Kernel& c = *target->kernels[ki];
      KernelInfo& cinf = plan->info[ki];
      if (cinf.absorbed_by >= 0 || cinf.overwrite) continue;
      EltSpec es;
      if (!match_eltwise(c, plan->shapes, es) || es.kind != ELT_SCALE_NEG) continue;
      const int G = es.read_tensor[0], P = c.write.tensor;
      if (!bucket.count(G) || prog->tdef(P).kind != TensorKind::Param) continue;
      if (shape_len(plan->shapes.at(G)) != shape_len(plan->shapes.at(P))) continue;
      bool clean = true;
      for (size_t kj = (size_t)plan->bucket_before_kernel; kj < ki && clean; ++kj) {
        if (plan->info[kj].in_exchange) continue;
        const Kernel& o = *target->kernels[kj];
        clean = o.write.tensor != P && o.write.tensor != G;
        for (auto& r : o.reads) clean = clean && r.tensor != P;
      }
      if (!clean) continue;
      ExchangeSeg sg;
      sg.off = (long long)((offs.at(G) - plan->bucket_off) / 4);
      sg.len = (long long)shape_len(plan->shapes.at(G));
      sg.rate = (float)es.lit[0];
      plan->exchange_segs.push_back(sg);
      plan->exchange_seg_param.push_back(P);
      cinf.in_exchange = true;
    }
    comm_open_window(comm, *ctx, plan->arena, plan->arena_bytes, plan->bucket_off, plan->bucket_bytes, plan->window);
  }
  Plan* raw = plan.get();
  raw->last_used = ++use_clock;
  build_nodes(*raw);
  plans.push_back(std::move(plan));
  return *raw;
}

// ------------------------------------------------------------------ execution

// May a HEAD node read its weight planes ahead of griddepcontrol.wait? `pred` is the node launched right before it on
// the same stream (null: none). Yes if that kernel triggers its dependents only AFTER its own wait (the single-CTA
// contraction kernels do, in their last epilogue): every earlier kernel of the stream is then complete when the head
// starts, producers on other streams are full dependencies anyway - only `pred` itself may still be running, so it
// must not be the one that writes those planes.
static bool head_may_read_weights_early(const Node* pred, const HeadParams& h) {
  if (!pred || pred->kind != Node::GEMM) return false;
  const GemmArgs& g = pred->gemm;
  if (g.bn == 0 && gemm_2cta_eligible(g)) return false;   // that kernel triggers at its start
  if ((g.flags & GEMM_SPLIT_OUT) && (g.out_hi == h.win_hi || g.out_hi == h.wout_hi)) return false;
  return true;
}

static void launch_node(Model& m, Node& n, cudaStream_t st) {
  Context& ctx = *m.ctx;
  switch (n.kind) {
    case Node::MEMSET: EGB_CUDA(cudaMemsetAsync(n.ptr, 0, n.bytes, st)); break;
    case Node::RANDOM:
      launch_fill_uniform(ctx, (float*)n.ptr, n.bytes / 4, n.lo, n.hi, m.seed, m.rng_counter, (uint64_t)n.tensor, st);
      break;
    case Node::SPLIT:
      if (!n.split_jobs.empty()) {
        launch_split_batch(ctx, n.split_jobs.data(), (int)n.split_jobs.size(), st, n.ptr, n.bytes);
        break;
      }
      launch_split_bf16(ctx, n.split_src, n.split_rows, n.split_cols, n.split_ld, n.split_transpose, n.split_hi,
                        n.split_mid, n.split_dst_ld, n.split_act, st);
      break;
    case Node::GEMM: launch_gemm_bf16x3(ctx, n.gemm, st); break;
    case Node::INTERP: launch_interp(ctx, n.ip, n.pb, n.rb, n.points_fast, n.strict, st, n.rsplit); break;
    case Node::SOFTMAX_XENT:
      launch_softmax_xent_rows(ctx, n.sx_h, n.sx_y, n.sx_dl, n.sx_s, n.sx_p, n.sx_dp, n.sx_dh, n.sx_ds, n.sx_rows, n.sx_cols,
                               n.sx_colsum, n.sx_out_hi, n.sx_out_mid, n.sx_ld_out, st);
      break;
    case Node::HEAD: launch_head_rows(ctx, n.head, st); break;
    case Node::HEADPREP: launch_head_tables(ctx, n.head, st); break;
    case Node::ELTWISE: launch_eltwise_stream(ctx, n.elt, st); break;
    case Node::EXCHANGE: launch_exchange(ctx, n.exchange, st); break;
    case Node::ROWCHAIN: launch_interp_rowchain(ctx, n.chain_progs, n.chain_n, n.chain_slots, n.chain_rows, st); break;
    case Node::CONV: {
      const ConvPattern& cv = n.conv;
      if (cv.kind == ConvPattern::FORWARD)
        launch_conv2_fwd(ctx, n.conv_a, n.conv_b, n.conv_out, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, n.conv_accumulate, st);
      else if (cv.kind == ConvPattern::D_FILTERS)
        launch_conv2_dw(ctx, n.conv_a, n.conv_b, n.conv_out, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, st);
      else
        launch_conv2_dimg(ctx, n.conv_a, n.conv_b, n.conv_out, cv.N, cv.H, cv.W, cv.C, cv.F, cv.KH, cv.KW, n.conv_accumulate, st);
      break;
    }
    case Node::ALLREDUCE:
      // the 256-byte alignment padding between bucket tensors travels with the payload (zeros)
      comm_all_reduce_avg(m.comm, (float*)n.ptr, n.bytes / 4, st);
      break;
    default: fail(EGB_ERR_RUNTIME, "internal: unknown plan node");
  }
}

// Issue the plan's nodes into the capturing stream level by level; within a level the nodes are
// spread over the main stream and two auxiliary streams (fork / join with events), which the capture
// turns into parallel branches of the CUDA graph.
static void capture_levels(Model& m, Plan& plan) {
  Context& c = *m.ctx;
  const int n = (int)plan.nodes.size();
  int max_level = 0;
  std::vector<int> width;
  for (auto& nd : plan.nodes) max_level = std::max(max_level, nd.level);
  width.assign(max_level + 1, 0);
  bool parallel = false;
  for (auto& nd : plan.nodes) parallel = parallel || ++width[nd.level] > 1;
  if (!m.concurrent || !parallel) {
    const Node* prev = nullptr;
    for (auto& nd : plan.nodes) {
      if (nd.kind == Node::HEAD) nd.head.w_early = head_may_read_weights_early(prev, nd.head) ? 1 : 0;
      launch_node(m, nd, c.stream);
      if (nd.kind != Node::MEMSET) prev = &nd;
    }
    return;
  }
  // Dependency-exact capture: every node waits (through events) only for the earlier nodes it really
  // conflicts with, so e.g. the weight-gradient contraction of layer l overlaps the input-gradient
  // chain of the layers below it. Nodes are issued in level order on eight streams; a node goes to a
  // stream whose last node is one of its predecessors (no false ordering), else to an idle one.
  constexpr int NS = 8;
  for (auto& st : c.aux_stream)
    if (!st) EGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  static cudaStream_t extra[NS - 3] = {};  // further capture streams, shared by all contexts of the process
  for (auto& st : extra)
    if (!st) EGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaStream_t streams[NS] = {c.stream, c.aux_stream[0], c.aux_stream[1], extra[0], extra[1], extra[2], extra[3], extra[4]};
  while ((int)c.fork_events.size() < n + 1) {
    cudaEvent_t e;
    EGB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c.fork_events.push_back(e);
  }
  auto hits = [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) {
    for (auto x : a)
      for (auto y : b)
        if (x == y) return true;
    return false;
  };
  std::vector<std::vector<int>> preds(n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const Node &a2 = plan.nodes[j], &b2 = plan.nodes[i];
      if (hits(a2.writes, b2.reads) || hits(a2.writes, b2.writes) || hits(a2.reads, b2.writes)) preds[i].push_back(j);
    }
  // transitive ancestors: a node may follow any of them on a stream without adding a dependency
  std::vector<std::vector<char>> anc(n, std::vector<char>(n, 0));
  for (int i = 0; i < n; ++i)
    for (int pj : preds[i]) {
      anc[i][pj] = 1;
      for (int a2 = 0; a2 < pj; ++a2)
        if (anc[pj][a2]) anc[i][a2] = 1;
    }
  // Critical path (longest chain by a rough cost): its nodes are issued back to back on the main stream,
  // where consecutive kernel launches become *programmatic* graph edges (the next kernel's launch latency
  // and prologue overlap the previous kernel; measured 4.5 us per full kernel -> kernel edge otherwise).
  // Everything else forks onto side streams.
  std::vector<double> cost(n), longest(n);
  std::vector<int> via(n, -1);
  for (int i = 0; i < n; ++i) {
    const Node& nd = plan.nodes[i];
    switch (nd.kind) {
      case Node::GEMM: cost[i] = 8.0 + 3e-9 * (double)nd.gemm.M * nd.gemm.N * nd.gemm.K; break;
      case Node::CONV: cost[i] = 50.0; break;
      case Node::MEMSET: cost[i] = 1.0; break;
      case Node::ALLREDUCE: cost[i] = 30.0; break;
      case Node::EXCHANGE: cost[i] = 12.0; break;
      default: cost[i] = 4.0;
    }
    longest[i] = cost[i];
    for (int pj : preds[i])
      if (longest[pj] + cost[i] > longest[i]) {
        longest[i] = longest[pj] + cost[i];
        via[i] = pj;
      }
  }
  std::vector<char> critical(n, 0);
  {
    int end = 0;
    for (int i = 1; i < n; ++i)
      if (longest[i] > longest[end]) end = i;
    for (int i = end; i >= 0; i = via[i]) critical[i] = 1;
    // A kernel with any cross-stream (full) dependency loses its programmatic edge as well (observed:
    // stream capture downgrades mixed dependencies), so the cheap ancestors of the critical path - operand
    // splits, seeds - are issued on the main stream too; a few serialised 2 us kernels at the start cost
    // less than a full-edge latency in front of every contraction of the chain.
    for (int i = 0; i < n; ++i) {
      if (critical[i] || cost[i] > 4.0) continue;
      for (int j = i + 1; j < n; ++j)
        if (critical[j] == 1 && anc[j][i]) {
          critical[i] = 2;
          break;
        }
    }
  }
  cudaEvent_t start = c.fork_events[n];
  EGB_CUDA(cudaEventRecord(start, c.stream));
  int tail[NS];
  bool forked[NS];
  for (int q = 0; q < NS; ++q) {
    tail[q] = -1;
    forked[q] = q == 0;
  }
  std::vector<int> stream_of(n, 0);
  for (int i = 0; i < n; ++i) {
    Node& nd = plan.nodes[i];
    int s = -1;
    if (nd.kind == Node::ALLREDUCE) {
      // Up to 4 ranks: all collectives on one dedicated stream, in plan order (every rank issues them in the
      // same order, and never two at a time), so that the all-reduce of an early bucket segment overlaps the
      // remaining adjoint contractions instead of sitting between them (2 GPUs: 131.8 -> 119.1 us per step).
      // At 8 ranks NCCL's kernels use many more CTAs and must be co-resident on every rank, while the
      // contractions hold one SM per CTA: the overlapped variant measured 298 us per step there against
      // ~210 us with the collectives in stream order - so they stay on the main stream.
      s = comm_world(m.comm) <= 4 ? NS - 1 : 0;
    } else if (nd.kind == Node::MEMSET || critical[i]) {
      s = 0;  // memset and the critical path stay on the main stream
    } else {
      // 1. a side stream whose tail is a predecessor (prefer the latest one)
      int best = -1;
      for (int q = 1; q < NS - 1; ++q)
        if (tail[q] >= 0 && std::find(preds[i].begin(), preds[i].end(), tail[q]) != preds[i].end() && tail[q] > best) {
          best = tail[q];
          s = q;
        }
      // 2. a side stream whose tail is an ancestor anyway (stream order adds no false dependency)
      for (int q = 1; q < NS - 1 && s < 0; ++q)
        if (tail[q] >= 0 && anc[i][tail[q]]) s = q;
      // 3. an unused side stream
      for (int q = 1; q < NS - 1 && s < 0; ++q)
        if (tail[q] < 0) s = q;
      // 4. the side stream whose tail is the oldest node (a false dependency; only when all streams are busy)
      if (s < 0) {
        s = 1;
        for (int q = 2; q < NS - 1; ++q)
          if (tail[q] < tail[s]) s = q;
      }
    }
    if (!forked[s]) {
      EGB_CUDA(cudaStreamWaitEvent(streams[s], start, 0));
      forked[s] = true;
    }
    for (int pj : preds[i])
      if (stream_of[pj] != s) EGB_CUDA(cudaStreamWaitEvent(streams[s], c.fork_events[pj], 0));
    if (nd.kind == Node::HEAD) nd.head.w_early = head_may_read_weights_early(tail[s] >= 0 ? &plan.nodes[tail[s]] : nullptr, nd.head) ? 1 : 0;
    launch_node(m, nd, streams[s]);
    EGB_CUDA(cudaEventRecord(c.fork_events[i], streams[s]));
    tail[s] = i;
    stream_of[i] = s;
  }
  for (int q = 1; q < NS; ++q)
    if (forked[q] && tail[q] >= 0) EGB_CUDA(cudaStreamWaitEvent(c.stream, c.fork_events[tail[q]], 0));
}

void Model::run(Plan& plan) {
  Context& c = *ctx;
  bool has_random = false, uses_epoch = false;
  for (auto& n : plan.nodes) {
    has_random = has_random || n.kind == Node::RANDOM;
    uses_epoch = uses_epoch || n.uses_epoch;
  }
  if (uses_epoch && plan.epoch_built != epoch) build_nodes(plan);
  const bool graphable = use_graphs && !c.timing && !has_random && plan.nodes.size() >= 3;
  if (graphable) {
    if (!plan.graph_valid) {
      if (plan.graph_exec) {
        cudaGraphExecDestroy(plan.graph_exec);
        plan.graph_exec = nullptr;
      }
      const size_t launches_before = c.launches;
      cudaGraph_t graph = nullptr;
      EGB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
      try {
        capture_levels(*this, plan);
      } catch (...) {
        cudaStreamEndCapture(c.stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      EGB_CUDA(cudaStreamEndCapture(c.stream, &graph));
      cudaError_t e = cudaGraphInstantiate(&plan.graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) fail(EGB_ERR_GPU, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      c.launches = launches_before;  // capture is not execution
      plan.graph_valid = true;
    }
    EGB_CUDA(cudaGraphLaunch(plan.graph_exec, c.stream));
    c.launches += plan.launches_per_run;
  } else {
    for (auto& n : plan.nodes) launch_node(*this, n, c.stream);
  }
  if (has_random) rng_counter++;
  plan.runs++;
  last_plan = &plan;
}

// ------------------------------------------------------------------ model construction

std::unique_ptr<Model> new_model(Context& ctx, std::shared_ptr<Program> prog, uint64_t seed) {
  if (prog->f64) fail(EGB_ERR_GENERATOR, "float64 models are not supported by the B200 backend (fp32 only)");
  if (!prog->compiled) compile_program(*prog);
  auto m = std::make_unique<Model>();
  m->ctx = &ctx;
  m->prog = prog;
  m->seed = seed;
  SplitMix rng(seed * 0x2545F4914F6CDD1Dull + 0x1234567ull);
  auto alloc_state = [&](int id, bool random) {
    const TensorDef& td = prog->tdef(id);
    for (auto d : td.shape)
      if (d < 0) fail(EGB_ERR_SHAPE, "tensor%d (%s) needs a static shape", id - 1, td.name.c_str());
    DevTensor t;
    t.shape = td.shape;
    t.bytes = (size_t)shape_len(td.shape) * 4;
    t.owned = true;
    EGB_CUDA(cudaMalloc(&t.ptr, std::max<size_t>(t.bytes, 4)));
    std::vector<float> host((size_t)shape_len(td.shape), 0.0f);
    if (random)
      for (auto& v : host) v = (float)(td.range_lo + (td.range_hi - td.range_lo) * rng.uniform());
    if (t.bytes) EGB_CUDA(cudaMemcpy(t.ptr, host.data(), t.bytes, cudaMemcpyHostToDevice));
    m->state[id] = t;
  };
  for (int id : prog->params) alloc_state(id, true);
  for (int id : prog->caches) alloc_state(id, false);
  return m;
}

}  // namespace egb
