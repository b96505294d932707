// codegen.cpp — walks the Trees graph and instantiates the sm_100a kernel templates (jit_templates.cuh).
//
// What the reference does per tree (one scalar work-item per output element, global id 0 = slowest dimension,
// OpenCLKernelBuilder.scala:135-221) is re-planned here for the GPU:
//   * the index space is linearised with the LAST dimension fastest and processed as 128-bit vectors by a
//     grid-stride loop over chunks (several independent vectors in flight per thread);
//   * every affine view (OpenCLKernelBuilder.scala:348-411) with integer coefficients is folded on the host into
//     `base + sum_x coef_x * g_x` over the flat source buffer, and bounds tests that interval analysis proves can
//     never fail are dropped; the rest keep the reference's two-sided test -> padding semantics;
//   * an unrolled chain `e_0 + e_1 + ... + e_{n-1}` (how users write per-axis sums and matmul, README.md:301-343,
//     benchmarks.scala:174-193) whose terms differ only by an affine step in their views is RE-ROLLED into a real
//     reduction over a new index t; `Tensor.join` of such terms is re-rolled into one more output dimension;
//   * if the reduced operand is a not-yet-evaluated inline tensor (`definition_root`), its closure is composed into
//     the reduction instead of being materialised (matmul2 would need an i*j*k intermediate, Tensors.scala:978-1003);
//     `sum_t A[i,t] * B[t,k]` is lowered to the tcgen05 3xTF32 contraction.
#include "codegen.h"

#include <mutex>
#include <unordered_map>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>

namespace cc {

// ---- planning knobs ------------------------------------------------------------------------------------------------------------------
// The CC_TUNE_* / CC_NO_* / CC_FUSE_* / CC_BATCHED_* environment switches shape the plan, but the kernel cache is keyed by the structure
// of the tree alone. Reading them live would let a switch flipped after the first compile change new plans while the cache keeps serving
// old ones for the same structure. They are therefore SAMPLED: once on first use, and again whenever the kernel cache is cleared
// (cc_kernel_cache_clear, cc_init) — the moment at which a changed switch can take effect consistently.
namespace {
std::mutex g_knob_mu;
std::unordered_map<std::string, std::string> g_knobs;
bool g_knobs_sampled = false;
const char* const kKnobNames[] = {"CC_BATCHED_CONTRACTION", "CC_FUSE_COL_STAGE", "CC_NO_OP_LOOPS", "CC_NO_STENCIL_TILE", "CC_TUNE_CONTRACTION_MIN_MACS",
                                  "CC_TUNE_GRID_MULT", "CC_TUNE_MIN_BLOCKS", "CC_TUNE_MIN_REROLL_TERMS", "CC_TUNE_STENCIL_RT", "CC_TUNE_T_GRID_MULT",
                                  "CC_TUNE_U", "CC_DISABLE_CONTRACTION", "CC_REDUCE_TILE_OWNER", "CC_TUNE_TILE_P", "CC_SMALL_N_MMA", "CC_TUNE_STENCIL_CTAS", "CC_TUNE_SMALL_N_PAIR"};
void sample_knobs_locked() {
  g_knobs.clear();
  for (const char* name : kKnobNames)
    if (const char* v = getenv(name)) g_knobs[name] = v;
  g_knobs_sampled = true;
}
}  // namespace

void plan_knobs_refresh() {
  std::lock_guard<std::mutex> lock(g_knob_mu);
  sample_knobs_locked();
}

const char* plan_knob(const char* name) {
  std::lock_guard<std::mutex> lock(g_knob_mu);
  if (!g_knobs_sampled) sample_knobs_locked();
  auto it = g_knobs.find(name);
  return it == g_knobs.end() ? nullptr : it->second.c_str();  // (the map only changes inside plan_knobs_refresh)
}


double java_decimal_round(double v) {
  if (!std::isfinite(v)) return v;
  double s = v * 1000.0;
  if (std::fabs(s) >= 9e15) return v;
  double f = std::floor(s);
  double res = std::fma(v, 1000.0, -f);  // exact residual in [0,1) up to one rounding
  while (res < 0) {
    f -= 1;
    res = std::fma(v, 1000.0, -f);
  }
  while (res >= 1) {
    f += 1;
    res = std::fma(v, 1000.0, -f);
  }
  if (res > 0.5)
    f += 1;
  else if (res == 0.5 && std::fmod(f, 2.0) != 0.0)
    f += 1;
  return f / 1000.0;
}

namespace {

struct Load {
  int arg = -1;
  std::vector<int64_t> src_shape;
  float padding = 0.f;
  std::vector<double> M;  // rows x (nd + 1), coefficients as the generated code uses them (DecimalFormat-rounded)
  int rows = 0;
  // analysis
  bool integer = true;
  std::vector<char> row_integer;
  std::vector<int64_t> coef;  // per index dim, in floats of the flat source
  int64_t base = 0;
  std::vector<char> need_lo, need_hi;  // per source dim
  int64_t max_abs_off = 0;
  bool reuse = false;  // some index dimension does not move the address: the same element is read from many index points
  bool any_check() const {
    for (size_t i = 0; i < need_lo.size(); ++i)
      if (need_lo[i] || need_hi[i]) return true;
    return false;
  }
};

struct Op {
  uint32_t kind = 0;
  int a = -1, b = -1;
  float lit = 0.f;
  int load = -1;
};

constexpr uint32_t K_ACC = 0xACC;  // pseudo-op of the epilogue: the value the reduction produced for this output element

struct Program {
  std::vector<int64_t> dims;  // index space: output dims, then the re-rolled reduction indices (outermost first) if any
  std::vector<Op> ops;        // the term (evaluated for every point of the index space)
  std::vector<Load> loads;
  std::vector<int> results;
  int shifted_loads = 0;      // loads running along the lanes at a constant offset off the 16-byte grid (dense stencil windows)
  int64_t tuple_inner = 1;    // several results (un-rolled join): the result index sits before the last dims whose product this is
  // reductions only
  int n_red = 0;              // number of trailing reduction dims
  uint32_t red_monoid = K_PLUS;  // what the re-rolled chain folds with: K_PLUS / K_TIMES / K_MIN / K_MAX
  std::vector<Op> post_ops;   // epilogue applied once per output element to the folded value (K_ACC); may use loads
  int post_result = -1;
  std::vector<char> load_in_post;  // per load: referenced by the epilogue (1) or by the term (0)
  bool trivial_post() const { return post_ops.size() == 1 && post_ops[0].kind == K_ACC; }
};

int64_t product(const std::vector<int64_t>& v) {
  int64_t p = 1;
  for (int64_t x : v) p *= x;
  return p;
}

void analyze_load(Load& L, const std::vector<int64_t>& dims) {
  const int nd = (int)dims.size();
  const int cols = nd + 1;
  L.rows = (int)L.src_shape.size();
  L.row_integer.assign(L.rows, 1);
  L.need_lo.assign(L.rows, 0);
  L.need_hi.assign(L.rows, 0);
  L.coef.assign(nd, 0);
  L.base = 0;
  L.integer = true;
  for (double& m : L.M) m = java_decimal_round(m);
  std::vector<int64_t> stride(L.rows, 1);
  for (int y = L.rows - 2; y >= 0; --y) stride[y] = stride[y + 1] * L.src_shape[y + 1];
  for (int y = 0; y < L.rows; ++y) {
    bool any = false;
    for (int x = 0; x < cols; ++x) {
      double m = L.M[(size_t)y * cols + x];
      if (m != 0.0) any = true;
      if (m != std::floor(m) || std::fabs(m) > 4e15) L.row_integer[y] = 0;
    }
    if (!L.row_integer[y]) {
      L.integer = false;
      L.need_lo[y] = L.need_hi[y] = 1;
      continue;
    }
    if (!any) continue;  // index 0, no test (K:383-385)
    int64_t lo = (int64_t)L.M[(size_t)y * cols + nd], hi = lo;
    for (int x = 0; x < nd; ++x) {
      int64_t m = (int64_t)L.M[(size_t)y * cols + x];
      int64_t ext = m * (dims[x] - 1);
      if (dims[x] == 0) ext = 0;
      lo += std::min<int64_t>(0, ext);
      hi += std::max<int64_t>(0, ext);
    }
    L.need_lo[y] = lo < 0;
    L.need_hi[y] = hi >= L.src_shape[y];
  }
  if (L.integer) {
    int64_t mx = 0;
    for (int y = 0; y < L.rows; ++y) L.base += (int64_t)L.M[(size_t)y * cols + nd] * stride[y];
    mx = std::llabs(L.base);
    for (int x = 0; x < nd; ++x) {
      for (int y = 0; y < L.rows; ++y) L.coef[x] += (int64_t)L.M[(size_t)y * cols + x] * stride[y];
      mx += std::llabs(L.coef[x]) * std::max<int64_t>(0, dims[x] - 1);
    }
    // row indices themselves are bounded by the same expression with stride 1
    L.max_abs_off = mx;
    for (int x = 0; x < nd; ++x)
      if (dims[x] > 1 && L.coef[x] == 0) L.reuse = true;
  }
}

// Moves index dimension `from` to position `to` (the dimensions in between shift by one): the same program over a permuted
// index space, i.e. a permuted output layout. Used to put a re-rolled join index where Tensor.join(tensors, dimension) wants it.
void move_index_dim(Program& p, int from, int to) {
  if (from == to) return;
  const int nd = (int)p.dims.size();
  std::vector<int> order;  // new position -> old position
  for (int x = 0; x < nd; ++x)
    if (x != from) order.push_back(x);
  order.insert(order.begin() + to, from);
  std::vector<int64_t> dims(nd);
  for (int x = 0; x < nd; ++x) dims[x] = p.dims[order[x]];
  for (Load& L : p.loads) {
    const int rows = (int)L.src_shape.size(), cols = nd + 1;
    std::vector<double> M(L.M.size());
    for (int y = 0; y < rows; ++y) {
      for (int x = 0; x < nd; ++x) M[(size_t)y * cols + x] = L.M[(size_t)y * cols + order[x]];
      M[(size_t)y * cols + nd] = L.M[(size_t)y * cols + nd];
    }
    L.M = std::move(M);
    L.reuse = false;
    analyze_load(L, dims);
  }
  p.dims = std::move(dims);
}

// ---- program construction ---------------------------------------------------------------------------------------

typedef std::unordered_map<uint32_t, std::vector<double>> StepMap;

struct Builder {
  const Tree& t;
  Program prog;
  int nd_base;                                                  // rank of the closure being exported
  std::vector<const StepMap*> exts;                             // per re-rolled index: Transform node -> extra matrix column
  std::unordered_map<uint32_t, int> memo;                       // node -> op index (ExportContext, R:220)
  std::map<uint32_t, int>* arg_of_param;                        // param node -> plan arg (shared across programs)
  std::vector<uint32_t>* arg_nodes;
  bool in_post = false;                                         // exporting the epilogue of a reduction (prog.post_ops)

  std::vector<Op>& target() { return in_post ? prog.post_ops : prog.ops; }
  int push(const Op& op) {
    target().push_back(op);
    return (int)target().size() - 1;
  }
  // switch to the epilogue: `acc_node` (the top of the re-rolled chain) becomes the folded value
  void begin_post(uint32_t acc_node) {
    in_post = true;
    memo.clear();
    Op op;
    op.kind = K_ACC;
    memo[acc_node] = push(op);
  }

  int arg_for(uint32_t param_node) {
    auto it = arg_of_param->find(param_node);
    if (it != arg_of_param->end()) return it->second;
    int a = (int)arg_nodes->size();
    arg_nodes->push_back(param_node);
    (*arg_of_param)[param_node] = a;
    return a;
  }

  int make_load(uint32_t extract_node) {
    const Node& ex = t.nodes[extract_node];
    const Node& arr = t.nodes[ex.kids[0]];
    const int nd = (int)prog.dims.size();
    Load L;
    if (arr.kind == K_PARAM) {
      // ArrayParameter.extract (K:435-455): indexed by the kernel's own ids, same rank as the kernel
      CC_REQUIRE((int)arr.shape.size() == nd_base, CC_ERR_BAD_TREE,
                 "direct Extract of a rank-%d array inside a rank-%d kernel", (int)arr.shape.size(), nd_base);
      L.arg = arg_for(ex.kids[0]);
      L.padding = arr.value;
      for (int32_t s : arr.shape) L.src_shape.push_back(s);
      L.M.assign((size_t)nd_base * (nd + 1), 0.0);
      for (int y = 0; y < nd_base; ++y) L.M[(size_t)y * (nd + 1) + y] = 1.0;
    } else {
      const Node& p = t.nodes[arr.kids[0]];
      CC_REQUIRE((int)arr.cols == nd_base + 1, CC_ERR_BAD_TREE, "transform has %u columns inside a rank-%d kernel",
                 arr.cols, nd_base);
      L.arg = arg_for(arr.kids[0]);
      L.padding = p.value;
      for (int32_t s : p.shape) L.src_shape.push_back(s);
      CC_REQUIRE(nd == nd_base + (int)exts.size(), CC_ERR_BAD_TREE, "index space rank %d != closure rank %d + %zu re-rolled indices", nd, nd_base,
                 exts.size());
      L.M.assign((size_t)arr.rows * (nd + 1), 0.0);
      for (uint32_t y = 0; y < arr.rows; ++y) {
        for (int x = 0; x < nd_base; ++x) L.M[(size_t)y * (nd + 1) + x] = arr.matrix[(size_t)y * arr.cols + x];
        for (size_t k = 0; k < exts.size(); ++k) {
          auto it = exts[k]->find(ex.kids[0]);
          if (it != exts[k]->end()) L.M[(size_t)y * (nd + 1) + nd_base + k] = it->second[y];
        }
        L.M[(size_t)y * (nd + 1) + nd] = arr.matrix[(size_t)y * arr.cols + nd_base];
      }
    }
    analyze_load(L, prog.dims);
    prog.loads.push_back(std::move(L));
    prog.load_in_post.push_back(in_post ? 1 : 0);
    return (int)prog.loads.size() - 1;
  }

  // iterative post-order export with identity memo
  int export_node(uint32_t root) {
    std::vector<std::pair<uint32_t, bool>> stack{{root, false}};
    while (!stack.empty()) {
      auto [i, ready] = stack.back();
      stack.pop_back();
      if (memo.count(i)) continue;
      const Node& nd = t.nodes[i];
      if (nd.kind == K_LITERAL) {
        Op op;
        op.kind = K_LITERAL;
        op.lit = nd.value;
        memo[i] = push(op);
      } else if (nd.kind == K_EXTRACT) {
        Op op;
        op.kind = K_EXTRACT;
        op.load = make_load(i);
        memo[i] = push(op);
      } else if (is_unary(nd.kind) || is_binary(nd.kind)) {
        if (!ready) {
          stack.push_back({i, true});
          for (size_t k = nd.kids.size(); k-- > 0;)
            if (!memo.count(nd.kids[k])) stack.push_back({nd.kids[k], false});
        } else {
          Op op;
          op.kind = nd.kind;
          op.a = memo.at(nd.kids[0]);
          if (nd.kids.size() > 1) op.b = memo.at(nd.kids[1]);
          memo[i] = push(op);
        }
      } else {
        fail(CC_ERR_BAD_TREE, strprintf("%s is only allowed at the root of a tree", kind_name(nd.kind)));
      }
    }
    return memo.at(root);
  }
};

// ---- congruence (re-rolling) --------------------------------------------------------------------------------------

// Are subtrees a and b the same expression up to the constant column of their Transforms? Records const(b) - const(a)
// per Transform node of a.
bool congruent(const Tree& t, uint32_t a0, uint32_t b0, std::unordered_map<uint32_t, std::vector<double>>& delta,
               std::unordered_map<uint32_t, uint32_t>* node_map = nullptr) {
  std::unordered_map<uint32_t, uint32_t> local_seen;
  std::unordered_map<uint32_t, uint32_t>& seen = node_map ? *node_map : local_seen;
  std::vector<std::pair<uint32_t, uint32_t>> stack{{a0, b0}};
  while (!stack.empty()) {
    auto [a, b] = stack.back();
    stack.pop_back();
    auto it = seen.find(a);
    if (it != seen.end()) {
      if (it->second != b) return false;
      continue;
    }
    seen[a] = b;
    const Node& na = t.nodes[a];
    const Node& nb = t.nodes[b];
    if (na.kind != nb.kind) return false;
    switch (na.kind) {
      case K_LITERAL:
        if (memcmp(&na.value, &nb.value, 4) != 0) return false;
        break;
      case K_PARAM:
        if (a != b) return false;
        break;
      case K_TRANSFORM: {
        if (na.kids[0] != nb.kids[0] || na.rows != nb.rows || na.cols != nb.cols) return false;
        std::vector<double> d(na.rows, 0.0);
        for (uint32_t y = 0; y < na.rows; ++y) {
          for (uint32_t x = 0; x + 1 < na.cols; ++x)
            if (na.matrix[(size_t)y * na.cols + x] != nb.matrix[(size_t)y * na.cols + x]) return false;
          d[y] = nb.matrix[(size_t)y * na.cols + na.cols - 1] - na.matrix[(size_t)y * na.cols + na.cols - 1];
        }
        delta[a] = std::move(d);
        break;
      }
      default:
        if (na.kids.size() != nb.kids.size()) return false;
        for (size_t k = 0; k < na.kids.size(); ++k) stack.push_back({na.kids[k], nb.kids[k]});
    }
  }
  return true;
}

// elements e_0..e_{n-1} -> per-Transform step such that const(e_i) = const(e_0) + i * step. Empty result = not congruent.
bool reroll(const Tree& t, const std::vector<uint32_t>& elems, std::unordered_map<uint32_t, std::vector<double>>& step) {
  if (elems.size() < 2) return false;
  if (!congruent(t, elems[0], elems[1], step)) return false;
  bool any = false;
  for (auto& kv : step)
    for (double d : kv.second)
      if (d != 0.0) any = true;
  if (!any) return false;
  for (size_t i = 2; i < elems.size(); ++i) {
    std::unordered_map<uint32_t, std::vector<double>> di;
    if (!congruent(t, elems[0], elems[i], di)) return false;
    if (di.size() != step.size()) return false;
    for (auto& kv : di) {
      auto it = step.find(kv.first);
      if (it == step.end()) return false;
      for (size_t y = 0; y < kv.second.size(); ++y)
        if (kv.second[y] != (double)i * it->second[y]) return false;
    }
  }
  return true;
}

// The monoids of MonoidPrograms (Tensors.scala:308-311). Users fold a split with any of them: `t.split(axis).reduce(_ + _)`
// (README.md:301-310), and equally `.reduce(Tensor.max)`, `.reduce(Tensor.min)`, `.reduce(_ * _)`.
inline bool is_monoid(uint32_t k) { return k == K_PLUS || k == K_TIMES || k == K_MIN || k == K_MAX; }

// left-leaning chain (((e0 op e1) op e2) op ...) of root's own operator -> [e0, e1, ...]
std::vector<uint32_t> plus_chain(const Tree& t, uint32_t root) {
  std::vector<uint32_t> elems;
  uint32_t cur = root;
  const uint32_t kind = t.nodes[root].kind;
  while (t.nodes[cur].kind == kind && is_monoid(kind)) {
    elems.push_back(t.nodes[cur].kids[1]);
    cur = t.nodes[cur].kids[0];
  }
  elems.push_back(cur);
  std::reverse(elems.begin(), elems.end());
  return elems;
}

// every leaf, left to right, of the maximal tree of root's own operator: ((e0 op e1) op (e2 op e3)) -> [e0, e1, e2, e3]
// (`parts.par.reduce(_ + _)`, reduceRight, hand-written pairwise sums). A node reached twice (a shared partial sum) is a leaf.
std::vector<uint32_t> monoid_tree_leaves(const Tree& t, uint32_t root) {
  std::vector<uint32_t> leaves;
  const uint32_t kind = t.nodes[root].kind;
  if (!is_monoid(kind)) return {root};
  std::unordered_map<uint32_t, int> visits;
  {
    std::vector<uint32_t> st{root};
    while (!st.empty()) {
      uint32_t i = st.back();
      st.pop_back();
      if (++visits[i] > 1 || t.nodes[i].kind != kind) continue;
      st.push_back(t.nodes[i].kids[0]);
      st.push_back(t.nodes[i].kids[1]);
    }
  }
  std::vector<uint32_t> st{root};
  while (!st.empty()) {
    uint32_t i = st.back();
    st.pop_back();
    if (t.nodes[i].kind == kind && (i == root || visits[i] == 1)) {
      st.push_back(t.nodes[i].kids[1]);
      st.push_back(t.nodes[i].kids[0]);
    } else {
      leaves.push_back(i);
    }
  }
  return leaves;
}

// shorter chains stay unrolled in one elementwise kernel, as the reference runs them (CC_TUNE_MIN_REROLL_TERMS: A/B knob)
const size_t kMinRerollTerms = [] {
  const char* e = plan_knob("CC_TUNE_MIN_REROLL_TERMS");
  return (size_t)(e ? std::max(2, atoi(e)) : 8);
}();

// A re-rolled reduction found inside an expression: `top` is the root of a left-leaning Plus chain whose terms are all
// congruent to terms[0], with Transform constants affine in a multi-index over `levels` (outermost first):
//   const(term i) = const(term 0) + sum_j digit_j(i) * steps[j]      (i = ((d_0 * n_1 + d_1) * n_2 + d_2) ...)
// One level is the plain per-axis sum / matmul pattern (README.md:301-343); three levels is the convolution of
// benchmarks.scala:526-545 (kernel row, kernel column, input channel).
struct Chain {
  uint32_t top = 0;
  std::vector<uint32_t> terms;
  std::vector<int64_t> levels;
  std::vector<StepMap> steps;  // keyed by the Transform nodes of terms[0]
  uint32_t monoid = K_PLUS;    // the chain's operator
};

bool nested_reroll(const Tree& t, Chain& ch) {
  const size_t n = ch.terms.size();
  if (n < 2) return false;
  std::vector<StepMap> D(n);
  std::vector<uint32_t> keys;
  for (size_t i = 1; i < n; ++i) {
    if (!congruent(t, ch.terms[0], ch.terms[i], D[i])) return false;
    if (i == 1)
      for (auto& kv : D[1]) keys.push_back(kv.first);
    else if (D[i].size() != keys.size())
      return false;
  }
  std::sort(keys.begin(), keys.end());
  size_t width = 0;
  for (uint32_t k : keys) width += D[1].at(k).size();
  std::vector<std::vector<double>> F(n, std::vector<double>(width, 0.0));  // flattened deltas, F[0] = 0
  for (size_t i = 1; i < n; ++i) {
    size_t o = 0;
    for (uint32_t k : keys) {
      auto it = D[i].find(k);
      if (it == D[i].end()) return false;
      for (double v : it->second) F[i][o++] = v;
    }
  }
  auto is_multiple = [&](const std::vector<double>& x, const std::vector<double>& step, double q) {
    for (size_t j = 0; j < width; ++j)
      if (x[j] != q * step[j]) return false;
    return true;
  };
  std::vector<int64_t> levels;               // innermost first while building
  std::vector<std::vector<double>> fsteps;
  size_t stride = 1;
  while (stride < n) {
    const size_t count = n / stride;
    const std::vector<double>& step = F[stride];
    size_t m = count;
    for (size_t q = 1; q < count; ++q)
      if (!is_multiple(F[q * stride], step, (double)q)) {
        m = q;
        break;
      }
    if (m < 2 || count % m != 0) return false;
    levels.push_back((int64_t)m);
    fsteps.push_back(step);
    stride *= m;
    if (levels.size() > 6) return false;
  }
  bool any = false;
  for (auto& st : fsteps)
    for (double v : st)
      if (v != 0.0) any = true;
  if (!any) return false;
  // every term must satisfy the digit formula
  for (size_t i = 1; i < n; ++i) {
    size_t rem = i;
    std::vector<double> want(width, 0.0);
    for (size_t j = 0; j < levels.size(); ++j) {
      const double digit = (double)(rem % (size_t)levels[j]);
      rem /= (size_t)levels[j];
      for (size_t x = 0; x < width; ++x) want[x] += digit * fsteps[j][x];
    }
    if (want != F[i]) return false;
  }
  ch.levels.assign(levels.rbegin(), levels.rend());
  ch.steps.clear();
  for (size_t j = levels.size(); j-- > 0;) {
    StepMap sm;
    size_t o = 0;
    for (uint32_t k : keys) {
      const size_t rows = D[1].at(k).size();
      sm[k] = std::vector<double>(fsteps[j].begin() + (long)o, fsteps[j].begin() + (long)(o + rows));
      o += rows;
    }
    ch.steps.push_back(std::move(sm));
  }
  return true;
}

// Finds the longest re-rollable Plus chain inside the float term `root` (searching through unary / binary nodes only).
bool find_chain(const Tree& t, uint32_t root, Chain& best) {
  std::vector<uint32_t> tops;
  std::unordered_map<uint32_t, char> seen, inner;
  std::vector<uint32_t> stack{root};
  while (!stack.empty()) {
    uint32_t i = stack.back();
    stack.pop_back();
    if (seen.count(i)) continue;
    seen[i] = 1;
    const Node& nd = t.nodes[i];
    if (is_monoid(nd.kind)) {
      std::vector<uint32_t> terms = plus_chain(t, i);
      // candidates: long left folds, and the ROOTS of other same-operator trees (their inner nodes are not candidates of their own)
      if (terms.size() >= kMinRerollTerms || (!inner.count(i) && terms.size() >= 2 && t.nodes[terms.back()].kind == nd.kind)) tops.push_back(i);
      for (uint32_t term : terms) {
        if (t.nodes[term].kind == nd.kind) inner[term] = 1;
        stack.push_back(term);  // the chain's own Plus nodes are not candidates
      }
    } else if (is_unary(nd.kind) || is_binary(nd.kind)) {
      for (uint32_t k : nd.kids) stack.push_back(k);
    }
  }
  size_t best_len = 0;
  for (uint32_t top : tops) {
    Chain c;
    c.top = top;
    c.monoid = t.nodes[top].kind;
    c.terms = plus_chain(t, top);
    if (c.terms.size() <= best_len && t.nodes[c.terms.back()].kind != c.monoid) continue;
    if (c.terms.size() < kMinRerollTerms || !nested_reroll(t, c)) {
      // `chain + other` parses as one longer left-leaning chain: keep the leading terms that are congruent to the first one
      // (their Plus node sits further down the left spine); the rest becomes part of the epilogue
      size_t m = 1;
      for (; m < c.terms.size(); ++m) {
        StepMap d;
        if (!congruent(t, c.terms[0], c.terms[m], d)) break;
      }
      bool ok = false;
      if (m != c.terms.size() && m >= kMinRerollTerms && m > best_len) {
        Chain part = c;
        uint32_t spine = top;
        for (size_t k = c.terms.size(); k > m; --k) spine = t.nodes[spine].kids[0];
        part.top = spine;
        part.terms.resize(m);
        part.levels.clear();
        part.steps.clear();
        if (nested_reroll(t, part)) {
          c = std::move(part);
          ok = true;
        }
      }
      if (!ok) {
        // not a left fold: any tree of the same operator (pairwise / parallel reduce, reduceRight) folds its leaves left to right
        Chain tree;
        tree.top = top;
        tree.monoid = c.monoid;
        tree.terms = monoid_tree_leaves(t, top);
        if (tree.terms.size() < kMinRerollTerms || tree.terms.size() <= best_len || !nested_reroll(t, tree)) continue;
        c = std::move(tree);
      }
    }
    // a small dense window over ONE source (3x3 / 5x5 box sums, max pooling: bare translated views, two or more window axes) is
    // faster unrolled -- its shifted loads share aligned vectors -- than as a reduction whose shift is a run-time value
    // (3x3 on 4096^2: 43 vs 72 us, 5x5: 96 vs 185 us)
    if (c.levels.size() >= 2 && c.terms.size() <= 32 && t.nodes[c.terms[0]].kind == K_EXTRACT) continue;
    best_len = c.terms.size();
    best = std::move(c);
  }
  if (best_len == 0) return false;
  // the epilogue (root with the chain replaced by its folded value) must not reach the Transforms the reduction steps apply to
  if (root != best.top) {
    std::unordered_map<uint32_t, char> s2;
    std::vector<uint32_t> st{root};
    while (!st.empty()) {
      uint32_t i = st.back();
      st.pop_back();
      if (i == best.top || s2.count(i)) continue;
      s2[i] = 1;
      const Node& nd = t.nodes[i];
      if (nd.kind == K_EXTRACT) {
        if (best.steps[0].count(nd.kids[0])) return false;
      } else {
        for (uint32_t k : nd.kids) st.push_back(k);
      }
    }
  }
  return true;
}

// When a join (Concatenate) has been re-rolled over its element index c, the reduction found in element 0 is also the
// reduction of every other element only if the step over c is the same for corresponding Transforms of every term.
bool join_step_uniform_over_terms(const Tree& t, const Chain& ch, const StepMap& step_c) {
  for (size_t i = 1; i < ch.terms.size(); ++i) {
    StepMap d;
    std::unordered_map<uint32_t, uint32_t> node_map;
    if (!congruent(t, ch.terms[0], ch.terms[i], d, &node_map)) return false;
    for (auto& kv : d) {
      auto a = step_c.find(kv.first);
      auto b = step_c.find(node_map.at(kv.first));
      if (a == step_c.end() || b == step_c.end() || a->second != b->second) return false;
    }
  }
  return true;
}

// ---- text helpers ---------------------------------------------------------------------------------------------------

std::string flit(float v) {
  if (std::isnan(v)) return "__int_as_float(0x7fc00000)";
  if (std::isinf(v)) return v > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
  uint32_t u;
  memcpy(&u, &v, 4);
  // bit-exact literal; the decimal form is kept in a comment for readability
  return strprintf("__uint_as_float(0x%08xu) /*%.9g*/", u, (double)v);
}

std::string dlit(double v) {
  return strprintf("%.17g", v);
}

struct Emit {
  std::string s;
  void operator()(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
    va_list ap;
    va_start(ap, fmt);
    va_list ap2;
    va_copy(ap2, ap);
    int n = vsnprintf(nullptr, 0, fmt, ap);
    va_end(ap);
    size_t old = s.size();
    s.resize(old + (size_t)n);
    vsnprintf(&s[old], (size_t)n + 1, fmt, ap2);
    va_end(ap2);
  }
};

const char* op_expr(uint32_t kind) {
  switch (kind) {
    case K_EXP: return "cc_exp(%s)";
    case K_LOG: return "cc_log(%s)";
    case K_ABS: return "fabsf(%s)";
    case K_TANH: return "cc_tanh(%s)";
    case K_SQRT: return "sqrtf(%s)";
    case K_NEG: return "(-%s)";
    case K_MIN: return "fminf(%s, %s)";
    case K_MAX: return "fmaxf(%s, %s)";
    case K_PLUS: return "(%s + %s)";
    case K_MINUS: return "(%s - %s)";
    case K_TIMES: return "(%s * %s)";
    case K_DIV: return "(%s / %s)";
    case K_PERCENT: return "fmodf(%s, %s)";
  }
  return "?";
}

// ---- iterated maps ------------------------------------------------------------------------------------------------------
//
// `(0 until n).foldLeft(x)(f)` (benchmarks.scala:100-108 `a * b + c` folded 100 times, :319-326 `tanh` folded 100 times) arrives
// as n copies of f's ops in the SSA list.  Emitted literally that is n inlined bodies per lane (NVRTC: 1.8 s for 100 x tanhf);
// a run of R >= 8 identical periods of P ops whose operands are loop invariants, values of the same period or values of the
// previous period is emitted as a counted loop with the previous-period values carried in registers.  Same operations in the
// same order, so the result does not change.
struct OpLoop {
  int s = 0, P = 0, R = 0;              // ops [s, s + P * R) are R repetitions of a P-op period
  std::vector<int> cls_a, cls_b;        // per position: -1 no operand, 0 invariant, 1 same period, 2 previous period
  std::vector<char> keep;               // per position: value needed after an iteration (carried or read after the loop)
  std::vector<int> init;                // per position: op holding the value carried into the first iteration (-1: none)
};

bool op_periodic(const std::vector<Op>& ops, int i, int P) {
  const Op& x = ops[(size_t)i];
  const Op& y = ops[(size_t)(i + P)];
  if (x.kind != y.kind || x.kind == K_EXTRACT || x.kind == K_ACC) return false;
  if (x.kind == K_LITERAL) return memcmp(&x.lit, &y.lit, 4) == 0;
  auto rel = [&](int a, int b) { return a < 0 ? b < 0 : (b == a || b == a + P); };
  return rel(x.a, y.a) && rel(x.b, y.b);
}

bool validate_loop(const std::vector<Op>& ops, const std::vector<int>& live, OpLoop& L) {
  const int s = L.s, P = L.P, R = L.R, end = s + P * R;
  L.cls_a.assign((size_t)P, -1);
  L.cls_b.assign((size_t)P, -1);
  L.keep.assign((size_t)P, 0);
  L.init.assign((size_t)P, -1);
  auto classify = [&](int j, int a0, int a1, int& cls) {  // operand of position j in repetition 0 / 1
    if (a1 < 0) return true;
    if (a1 == a0) {
      cls = 0;
      return a0 < s;
    }
    if (a1 >= s + P) {
      cls = 1;
      return true;
    }
    if (a1 < s) return false;  // reaches further back than one period
    cls = 2;
    const int jp = a1 - s;
    if (L.init[(size_t)jp] >= 0 && L.init[(size_t)jp] != a0) return false;
    if (a0 >= s) return false;
    L.init[(size_t)jp] = a0;
    L.keep[(size_t)jp] = 1;
    (void)j;
    return true;
  };
  for (int j = 0; j < P; ++j) {
    const Op& r0 = ops[(size_t)(s + j)];
    const Op& r1 = ops[(size_t)(s + P + j)];
    if (!classify(j, r0.a, r1.a, L.cls_a[(size_t)j]) || !classify(j, r0.b, r1.b, L.cls_b[(size_t)j])) return false;
  }
  // every later repetition follows the same classes
  for (int r = 1; r + 1 < R; ++r)
    for (int j = 0; j < P; ++j) {
      const Op& x = ops[(size_t)(s + r * P + j)];
      const Op& y = ops[(size_t)(s + (r + 1) * P + j)];
      auto same = [&](int a, int b, int cls) { return cls < 0 ? b < 0 : (cls == 0 ? b == a : b == a + P); };
      if (!same(x.a, y.a, L.cls_a[(size_t)j]) || !same(x.b, y.b, L.cls_b[(size_t)j])) return false;
    }
  // values read after the loop must belong to the last repetition
  auto outside_ref = [&](int a) {
    if (a < s || a >= end) return true;
    if (a < end - P) return false;
    L.keep[(size_t)(a - (end - P))] = 1;
    return true;
  };
  for (size_t i = (size_t)end; i < ops.size(); ++i) {
    if (ops[i].kind == K_LITERAL || ops[i].kind == K_EXTRACT || ops[i].kind == K_ACC) continue;
    if (!outside_ref(ops[i].a)) return false;
    if (ops[i].b >= 0 && !outside_ref(ops[i].b)) return false;
  }
  for (int a : live)
    if (!outside_ref(a)) return false;
  return true;
}

std::vector<OpLoop> find_op_loops(const std::vector<Op>& ops, const std::vector<int>& live) {
  std::vector<OpLoop> loops;
  const int n = (int)ops.size();
  constexpr int kMinReps = 8, kMaxPeriod = 64;
  if (n < kMinReps || n > 200000) return loops;
  if (const char* ev = plan_knob("CC_NO_OP_LOOPS"))  // A/B switch for tests: emit iterated maps unrolled, as the reference does
    if (atoi(ev) != 0) return loops;
  long budget = 4000000;
  int i = 0;
  while (i + kMinReps <= n && budget > 0) {
    OpLoop best;
    for (int P = 1; P <= kMaxPeriod && i + P * kMinReps <= n; ++P) {
      int m = 0;
      while (i + m + P < n && op_periodic(ops, i + m, P)) ++m;
      budget -= m + 1;
      const int R = m / P + 1;
      if (R < kMinReps || P * R <= best.P * best.R) continue;
      OpLoop c;
      c.s = i;
      c.P = P;
      c.R = R;
      if (validate_loop(ops, live, c)) best = std::move(c);
      budget -= n;
    }
    if (best.R > 0) {
      i = best.s + best.P * best.R;
      loops.push_back(std::move(best));
    } else {
      ++i;
    }
  }
  return loops;
}

// Emits the SSA body of the program for one lane; value names are <prefix><op>; loads read L<j>[lane].  `live` lists the ops the
// caller reads afterwards.
void emit_op_list(Emit& e, const std::vector<Op>& ops, const char* indent, const std::string& lane, const std::vector<int>& live,
                  const char* prefix = "_", const char* acc = nullptr) {
  const std::vector<OpLoop> loops = find_op_loops(ops, live);
  size_t next_loop = 0;
  auto one = [&](const Op& op, const std::string& name, const std::string& a, const std::string& b, const char* ind) {
    if (op.kind == K_ACC) {
      e("%sconst float %s = %s;\n", ind, name.c_str(), acc);
    } else if (op.kind == K_LITERAL) {
      e("%sconst float %s = %s;\n", ind, name.c_str(), flit(op.lit).c_str());
    } else if (op.kind == K_EXTRACT) {
      e("%sconst float %s = L%d[%s];\n", ind, name.c_str(), op.load, lane.c_str());
    } else {
      std::string fmt = op_expr(op.kind);
      std::string ex = is_unary(op.kind) ? strprintf(fmt.c_str(), a.c_str()) : strprintf(fmt.c_str(), a.c_str(), b.c_str());
      e("%sconst float %s = %s;\n", ind, name.c_str(), ex.c_str());
    }
  };
  for (size_t i = 0; i < ops.size();) {
    if (next_loop < loops.size() && (size_t)loops[next_loop].s == i) {
      const OpLoop& L = loops[next_loop++];
      const int last = L.s + L.P * (L.R - 1);
      for (int j = 0; j < L.P; ++j)
        if (L.keep[(size_t)j]) {
          if (L.init[(size_t)j] >= 0)
            e("%sfloat %s%d = %s%d;\n", indent, prefix, last + j, prefix, L.init[(size_t)j]);
          else
            e("%sfloat %s%d = 0.f;\n", indent, prefix, last + j);
        }
      const int unroll = L.P <= 4 ? 4 : (L.P <= 16 ? 2 : 1);
      e("%s#pragma unroll %d\n%sfor (int it_ = 0; it_ < %d; ++it_) {  // %d x a period of %d ops\n", indent, unroll, indent, L.R, L.R, L.P);
      const std::string ind2 = std::string(indent) + "  ";
      for (int j = 0; j < L.P; ++j) {
        const Op& op = ops[(size_t)(L.s + L.P + j)];  // repetition 1: its operand indices are in canonical position
        auto nm = [&](int a, int cls) -> std::string {
          if (cls == 0) return strprintf("%s%d", prefix, a);
          if (cls == 1) return strprintf("t%d_", a - (L.s + L.P));
          if (cls == 2) return strprintf("%s%d", prefix, last + (a - L.s));
          return "";
        };
        one(op, strprintf("t%d_", j), nm(op.a, L.cls_a[(size_t)j]), nm(op.b, L.cls_b[(size_t)j]), ind2.c_str());
      }
      for (int j = 0; j < L.P; ++j)
        if (L.keep[(size_t)j]) e("%s%s%d = t%d_;\n", ind2.c_str(), prefix, last + j, j);
      e("%s}\n", indent);
      i = (size_t)(L.s + L.P * L.R);
      continue;
    }
    const Op& op = ops[i];
    one(op, strprintf("%s%zu", prefix, i), strprintf("%s%d", prefix, op.a), strprintf("%s%d", prefix, op.b), indent);
    ++i;
  }
}

void emit_ops(Emit& e, const Program& p, const char* indent, const std::string& lane) {
  emit_op_list(e, p.ops, indent, lane, p.results);
}

struct LoadCtx {
  int V;                 // lanes
  int vdim;              // index dim the lanes run along (-1: none)
  const char* idx_type;  // "int" or "long long"
  // bounds-tested vector loads are issued unconditionally from a clamped address and the padding is selected afterwards: in the
  // straight-line body of an unrolled reduction a branch around the load would pin it next to its use, exposing its whole latency
  bool unconditional = false;
  // V == 1 only: fetch the element and its successor along the fastest source dimension as one 8-byte load into L[0], L[1] (the caller has
  // checked that the offset is even and that both share their bounds tests)
  bool pair = false;
};

// index expression of source row y for lane `lane` ("" = lane 0 / no lane term)
std::string row_index_expr(const Load& L, int y, int nd, int vdim, const std::string& lane) {
  const int cols = nd + 1;
  std::string s;
  if (L.row_integer[y]) {
    int64_t c = (int64_t)L.M[(size_t)y * cols + nd];
    s = strprintf("%lld", (long long)c);
    for (int x = 0; x < nd; ++x) {
      int64_t m = (int64_t)L.M[(size_t)y * cols + x];
      if (m == 0) continue;
      if (x == vdim && !lane.empty())
        s += strprintf(" + %lld * (g%d + %s)", (long long)m, x, lane.c_str());
      else
        s += strprintf(" + %lld * g%d", (long long)m, x);
    }
    return s;
  }
  // non-integer row: the reference evaluates `(int)(g_x * a + ... + c)` in double, left to right (K:363-386)
  bool first = true;
  for (int x = 0; x <= nd; ++x) {
    double m = L.M[(size_t)y * cols + x];
    if (m == 0.0) continue;
    std::string term;
    if (x == nd)
      term = dlit(m);
    else {
      std::string g = (x == vdim && !lane.empty()) ? strprintf("(double)(g%d + %s)", x, lane.c_str()) : strprintf("(double)g%d", x);
      term = (m == 1.0) ? g : g + " * " + dlit(m);
    }
    s += first ? term : " + " + term;
    first = false;
  }
  if (first) s = "0.0";
  return "(int)(" + s + ")";
}

// Emits code filling `float L<j>[V]` for load j. g<x> index variables are in scope.
void emit_load(Emit& e, const Program& p, int j, const LoadCtx& c, const char* indent) {
  const Load& L = p.loads[j];
  const int nd = (int)p.dims.size();
  const int V = c.V;
  std::string pad = flit(L.padding);
  // streaming (no L1 allocation) for data read once; cached for data reused across the index space (broadcasts, reductions)
  const char* LD1 = L.reuse ? "cc_ldc" : "cc_ldg";
  const char* LD4 = L.reuse ? "cc_ldc4" : "cc_ldg4";
  e("%sfloat L%d[%d];\n", indent, j, c.pair ? 2 : V);
  if (!L.integer) {
    // general path: per lane, per row index in the reference's arithmetic
    e("%s#pragma unroll\n%sfor (int l = 0; l < %d; ++l) {\n", indent, indent, V);
    std::string cond, off = "0";
    int64_t stride = 1;
    std::vector<int64_t> strides(L.rows, 1);
    for (int y = L.rows - 2; y >= 0; --y) strides[y] = strides[y + 1] * L.src_shape[y + 1];
    (void)stride;
    for (int y = 0; y < L.rows; ++y) {
      e("%s  const long long i%d = %s;\n", indent, y, row_index_expr(L, y, nd, c.vdim, "l").c_str());
      if (L.need_lo[y]) cond += strprintf("%si%d >= 0", cond.empty() ? "" : " && ", y);
      if (L.need_hi[y]) cond += strprintf("%si%d < %lld", cond.empty() ? "" : " && ", y, (long long)L.src_shape[y]);
      off += strprintf(" + i%d * %lldLL", y, (long long)strides[y]);
    }
    if (cond.empty())
      e("%s  L%d[l] = %s(p%d + (%s));\n", indent, j, LD1, L.arg, off.c_str());
    else
      e("%s  L%d[l] = (%s) ? %s(p%d + (%s)) : %s;\n", indent, j, cond.c_str(), LD1, L.arg, off.c_str(), pad.c_str());
    e("%s}\n", indent);
    return;
  }
  const int64_t coefV = (c.vdim >= 0 && V > 1) ? L.coef[c.vdim] : 0;
  // offset of lane 0
  std::string off = strprintf("(%s)%lld", c.idx_type, (long long)L.base);
  for (int x = 0; x < nd; ++x)
    if (L.coef[x] != 0) off += strprintf(" + (%s)%lld * g%d", c.idx_type, (long long)L.coef[x], x);
  e("%sconst %s o%d = %s;\n", indent, c.idx_type, j, off.c_str());
  // bounds tests: uniform rows vs lane-dependent rows
  std::string ucond;
  bool lane_checks = false;
  const int cols = nd + 1;
  for (int y = 0; y < L.rows; ++y) {
    if (!L.need_lo[y] && !L.need_hi[y]) continue;
    bool lane_dep = V > 1 && c.vdim >= 0 && L.M[(size_t)y * cols + c.vdim] != 0.0;
    if (lane_dep) {
      lane_checks = true;
      continue;
    }
    e("%sconst %s i%d_%d = %s;\n", indent, c.idx_type, j, y, row_index_expr(L, y, nd, c.vdim, "").c_str());
    if (L.need_lo[y]) ucond += strprintf("%si%d_%d >= 0", ucond.empty() ? "" : " && ", j, y);
    if (L.need_hi[y]) ucond += strprintf("%si%d_%d < %lld", ucond.empty() ? "" : " && ", j, y, (long long)L.src_shape[y]);
  }
  bool aligned = V == 4 && coefV == 1 && (L.base % 4 == 0);
  if (aligned)
    for (int x = 0; x < nd; ++x)
      if (x != c.vdim && L.coef[x] % 4 != 0) aligned = false;
  if (V == 1 && c.pair) {
    e("%s{\n%s  const bool in_ = %s;\n%s  const float2 t_ = cc_ldc2(p%d + (in_ ? o%d : (%s)0));\n", indent, indent, ucond.empty() ? "true" : ucond.c_str(), indent, L.arg, j,
      c.idx_type);
    e("%s  L%d[0] = in_ ? t_.x : %s;\n%s  L%d[1] = in_ ? t_.y : %s;\n%s}\n", indent, j, pad.c_str(), indent, j, pad.c_str(), indent);
    return;
  }
  if (V == 1) {
    if (ucond.empty())
      e("%sL%d[0] = %s(p%d + o%d);\n", indent, j, LD1, L.arg, j);
    else
      e("%sL%d[0] = (%s) ? %s(p%d + o%d) : %s;\n", indent, j, ucond.c_str(), LD1, L.arg, j, pad.c_str());
    return;
  }
  if (!lane_checks && aligned) {
    if (ucond.empty())
      e("%s%s(p%d + o%d, L%d);\n", indent, LD4, L.arg, j, j);
    else if (c.unconditional) {
      e("%sconst bool in%d = %s;\n%s%s(p%d + (in%d ? o%d : (%s)0), L%d);\n", indent, j, ucond.c_str(), indent, LD4, L.arg, j, j, c.idx_type, j);
      e("%s#pragma unroll\n%sfor (int l = 0; l < 4; ++l) L%d[l] = in%d ? L%d[l] : %s;\n", indent, indent, j, j, j, pad.c_str());
    } else
      e("%sif (%s) %s(p%d + o%d, L%d); else { L%d[0] = L%d[1] = L%d[2] = L%d[3] = %s; }\n", indent, ucond.c_str(), LD4,
        L.arg, j, j, j, j, j, j, pad.c_str());
    return;
  }
  if (!lane_checks && coefV == 0) {
    if (ucond.empty())
      e("%sL%d[0] = %s(p%d + o%d);\n", indent, j, LD1, L.arg, j);
    else
      e("%sL%d[0] = (%s) ? %s(p%d + o%d) : %s;\n", indent, j, ucond.c_str(), LD1, L.arg, j, pad.c_str());
    e("%sL%d[1] = L%d[2] = L%d[3] = L%d[0];\n", indent, j, j, j, j);
    return;
  }
  // Unit stride along the lanes, but shifted off the 16-byte grid by a constant (a translation along the fastest dimension) and /
  // or with bounds that differ per lane: read the (at most two) ALIGNED vectors that cover the four lanes and pick the lanes out of
  // them -- 1-2 vector loads instead of 4 scalar ones, and the windows of a stencil (x.translate(dy, dx) for dx = -1, 0, 1) share
  // their vectors. The addresses are clamped into the buffer instead of tested: a clamped vector only ever feeds lanes whose own
  // index is out of range, and those take the padding.
  {
    const int64_t TOTAL = product(L.src_shape);
    // (worth it from four such loads on: a 3x3 window 58 -> 43 us on 4096^2, 5x5 169 -> 96 us; with the two of a 5-point stencil the
    // extra bytes pulled through L1 cost more than the saved instructions: 160 -> 183 us)
    bool shifted = V == 4 && coefV == 1 && c.vdim >= 0 && TOTAL >= 8 && TOTAL % 4 == 0 && p.shifted_loads >= 4;
    for (int x = 0; shifted && x < nd; ++x)
      if (x != c.vdim && p.dims[x] > 1 && L.coef[x] % 4 != 0) shifted = false;
    if (shifted) {
      const int sft = (int)(((L.base % 4) + 4) % 4);
      const char* I = c.idx_type;
      e("%sconst %s oa%d = o%d - %d;\n", indent, I, j, j, sft);
      e("%sfloat A%d[4], B%d[4];\n", indent, j, j);
      e("%scc_ldc4(p%d + (oa%d < 0 ? (%s)0 : (oa%d > (%s)%lld ? (%s)%lld : oa%d)), A%d);\n", indent, L.arg, j, I, j, I, (long long)(TOTAL - 4), I,
        (long long)(TOTAL - 4), j, j);
      if (sft != 0)
        e("%scc_ldc4(p%d + (oa%d + 4 < 0 ? (%s)0 : (oa%d + 4 > (%s)%lld ? (%s)%lld : oa%d + 4)), B%d);\n", indent, L.arg, j, I, j, I, (long long)(TOTAL - 4), I,
          (long long)(TOTAL - 4), j, j);
      e("%s#pragma unroll\n%sfor (int l = 0; l < 4; ++l) {\n", indent, indent);
      std::string cond = ucond;
      for (int y = 0; y < L.rows; ++y) {
        if (!L.need_lo[y] && !L.need_hi[y]) continue;
        if (L.M[(size_t)y * cols + c.vdim] == 0.0) continue;
        e("%s  const %s k%d = %s;\n", indent, I, y, row_index_expr(L, y, nd, c.vdim, "l").c_str());
        if (L.need_lo[y]) cond += strprintf("%sk%d >= 0", cond.empty() ? "" : " && ", y);
        if (L.need_hi[y]) cond += strprintf("%sk%d < %lld", cond.empty() ? "" : " && ", y, (long long)L.src_shape[y]);
      }
      if (sft != 0)
        e("%s  const float raw = (l + %d < 4) ? A%d[(l + %d) & 3] : B%d[(l + %d) & 3];\n", indent, sft, j, sft, j, sft);
      else
        e("%s  const float raw = A%d[l];\n", indent, j);
      if (cond.empty())
        e("%s  L%d[l] = raw;\n", indent, j);
      else
        e("%s  L%d[l] = (%s) ? raw : %s;\n", indent, j, cond.c_str(), pad.c_str());
      e("%s}\n", indent);
      return;
    }
  }
  // per-lane scalar loads (strided / lane-dependent bounds on a non-unit stride); lanes a few floats apart share 32-byte sectors,
  // which a non-allocating load would fetch from L2 once per lane
  if (std::llabs(coefV) <= 8) LD1 = "cc_ldc";
  e("%s#pragma unroll\n%sfor (int l = 0; l < %d; ++l) {\n", indent, indent, V);
  std::string cond = ucond;
  for (int y = 0; y < L.rows; ++y) {
    if (!L.need_lo[y] && !L.need_hi[y]) continue;
    bool lane_dep = L.M[(size_t)y * cols + c.vdim] != 0.0;
    if (!lane_dep) continue;
    e("%s  const %s k%d = %s;\n", indent, c.idx_type, y, row_index_expr(L, y, nd, c.vdim, "l").c_str());
    if (L.need_lo[y]) cond += strprintf("%sk%d >= 0", cond.empty() ? "" : " && ", y);
    if (L.need_hi[y]) cond += strprintf("%sk%d < %lld", cond.empty() ? "" : " && ", y, (long long)L.src_shape[y]);
  }
  if (cond.empty())
    e("%s  L%d[l] = %s(p%d + o%d + (%s)%lld * l);\n", indent, j, LD1, L.arg, j, c.idx_type, (long long)coefV);
  else
    e("%s  L%d[l] = (%s) ? %s(p%d + o%d + (%s)%lld * l) : %s;\n", indent, j, cond.c_str(), LD1, L.arg, j, c.idx_type,
      (long long)coefV, pad.c_str());
  e("%s}\n", indent);
}

std::string param_list(int n_args, bool with_out, const char* out_name = "out") {
  std::string s;
  for (int i = 0; i < n_args; ++i) s += strprintf("%sconst float* __restrict__ p%d", i ? ", " : "", i);
  if (with_out) s += strprintf("%sfloat* __restrict__ %s", n_args ? ", " : "", out_name);
  return s;
}
std::string arg_pass(int n_args) {
  std::string s;
  for (int i = 0; i < n_args; ++i) s += strprintf("%sp%d", i ? ", " : "", i);
  return s;
}

// decode `e` (element index over dims[0..n)) into g0..g{n-1}; dims beyond `n` are not touched
void emit_decode(Emit& e, const std::vector<int64_t>& dims, int n, const char* idx_type, const char* src, const char* indent) {
  if (n == 0) return;
  e("%s%s rem_ = %s;\n", indent, idx_type, src);
  for (int x = n - 1; x >= 1; --x) {
    e("%sconst %s g%d = rem_ %% (%s)%lld; rem_ /= (%s)%lld;\n", indent, idx_type, x, idx_type, (long long)dims[x], idx_type,
      (long long)dims[x]);
  }
  e("%sconst %s g0 = rem_;\n", indent, idx_type);
}

bool program_is_flat(const Program& p) {
  // every load reads src[linear output index] with no test
  const int nd = (int)p.dims.size();
  std::vector<int64_t> stride(nd, 1);
  for (int x = nd - 2; x >= 0; --x) stride[x] = stride[x + 1] * p.dims[x + 1];
  for (const Load& L : p.loads) {
    if (!L.integer || L.any_check() || L.base != 0) return false;
    for (int x = 0; x < nd; ++x)
      if (p.dims[x] > 1 && L.coef[x] != stride[x]) return false;
  }
  return true;
}

const char* pick_idx_type(const Program& p, int64_t total) {
  int64_t mx = total;
  for (const Load& L : p.loads) {
    if (!L.integer) return "long long";
    mx = std::max(mx, L.max_abs_off);
    mx = std::max<int64_t>(mx, product(L.src_shape));
  }
  return mx >= (int64_t)1 << 30 ? "long long" : "int";
}

uint64_t count_flops(const Program& p) {
  uint64_t f = 0;
  for (const Op& op : p.ops)
    if (op.kind != K_LITERAL && op.kind != K_EXTRACT) ++f;
  return f;
}

// ---- elementwise kernel ------------------------------------------------------------------------------------------

void emit_elementwise(Plan& plan, const Program& p, int n_args, const DeviceProps& dev, int concat_fallback_n) {
  const int nd = (int)p.dims.size();
  const int64_t total = product(p.dims);
  const int nres = (int)p.results.size();
  const bool flat = program_is_flat(p);
  int V = (nd >= 1 && p.dims[nd - 1] % 4 == 0 && nres == 1) ? 4 : 1;
  // every load reads src[flat index]: the shape is irrelevant, vectorise over the flat index; the <= 3 leftover elements of an odd
  // element count are done by scalar code in block 0
  if (flat && nres == 1 && total >= 4) V = 4;
  const int64_t NV = total / V;
  const int64_t tail = total - NV * V;
  const char* IDX = pick_idx_type(p, total * std::max(1, nres));
  const int nloads = (int)p.loads.size();
  // Measured on B200 (scripts/gpu_sweep_c2.sh, profiles/r01_sweep_c2.log): one CTA per chunk (no persistence), 2 vectors in
  // flight per thread and a register cap that keeps 8 CTAs/SM resident beats a persistent grid by ~15 % (7.0 vs 6.1 TB/s on C2).
  int U = 2;
  if (nloads > 8) U = 1;
  // scalar lanes (a fastest dimension that is not a multiple of 4 under a view): 2 x 4 bytes per thread in flight cannot cover the
  // HBM latency (1001x1003x127 translate: 0.68 of HBM); take 8 elements per thread
  if (V == 1 && nloads <= 4) U = 8;
  while (U > 1 && NV < (int64_t)256 * U * dev.sm_count * 8) U /= 2;  // small tensors: more CTAs beats more vectors per thread
  int64_t grid_mult = (int64_t)1 << 40;
  int min_blocks = nloads * V * U <= 24 ? 8 : (nloads * V * U <= 40 ? 6 : 0);
  if (const char* e = plan_knob("CC_TUNE_U")) U = std::max(1, atoi(e));          // tuning knobs (scripts/gpu_sweep_c2.sh)
  if (const char* e = plan_knob("CC_TUNE_GRID_MULT")) grid_mult = std::max(1, atoi(e));
  if (const char* e = plan_knob("CC_TUNE_MIN_BLOCKS")) min_blocks = std::max(0, atoi(e));
  const int64_t chunk = (int64_t)256 * U;
  const int64_t nchunks = (NV + chunk - 1) / chunk;
  int64_t grid = std::min<int64_t>(nchunks, std::min<int64_t>((int64_t)dev.sm_count * grid_mult, 0x7fffffff));
  if (grid < 1) grid = 1;

  Emit e;
  e("// elementwise: dims=[");
  for (int x = 0; x < nd; ++x) e("%s%lld", x ? "," : "", (long long)p.dims[x]);
  e("] V=%d U=%d flat=%d idx=%s loads=%d ops=%zu grid=%lld\n", V, U, (int)flat, IDX, nloads, p.ops.size(), (long long)grid);
  e("struct Regs {");
  for (int j = 0; j < nloads; ++j) e(" float L%d[%d];", j, V);
  if (nloads == 0) e(" int unused_;");
  e(" };\n");
  // ld
  e("__device__ __forceinline__ void ld(const %s v%s%s, Regs& r) {\n", IDX, n_args ? ", " : "", param_list(n_args, false).c_str());
  if (nloads) {
    LoadCtx c{V, nd - 1, IDX};
    if (flat) {
      for (int j = 0; j < nloads; ++j) {
        if (V == 4)
          e("  cc_ldg4(p%d + v * 4, r.L%d);\n", p.loads[j].arg, j);
        else
          e("  r.L%d[0] = cc_ldg(p%d + v);\n", j, p.loads[j].arg);
      }
    } else {
      emit_decode(e, p.dims, nd, IDX, strprintf("v * %d", V).c_str(), "  ");
      for (int j = 0; j < nloads; ++j) {
        e("  {\n");
        emit_load(e, p, j, c, "    ");
        e("    #pragma unroll\n    for (int l = 0; l < %d; ++l) r.L%d[l] = L%d[l];\n  }\n", V, j, j);
      }
    }
  }
  e("}\n");
  // st
  e("__device__ __forceinline__ void st(const %s v, const Regs& r, float* __restrict__ out) {\n", IDX);
  e("  float o[%d][%d];\n", nres, V);
  e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V);
  for (int j = 0; j < nloads; ++j) e("    const float* L%d = r.L%d;\n", j, j);
  emit_ops(e, p, "    ", "l");
  for (int r = 0; r < nres; ++r) e("    o[%d][l] = _%d;\n", r, p.results[r]);
  e("  }\n");
  if (nres == 1) {
    if (V == 4)
      e("  cc_stg4(out + v * 4, o[0]);\n");
    else
      e("  out[v] = o[0][0];\n");
  } else {
    if (p.tuple_inner == 1) {
      for (int r = 0; r < nres; ++r) e("  out[v * %d + %d] = o[%d][0];\n", nres, r, r);
    } else {
      e("  const %s vo = v / (%s)%lld, vi = v - vo * (%s)%lld;\n", IDX, IDX, (long long)p.tuple_inner, IDX, (long long)p.tuple_inner);
      for (int r = 0; r < nres; ++r)
        e("  out[(vo * %d + %d) * (%s)%lld + vi] = o[%d][0];\n", nres, r, IDX, (long long)p.tuple_inner, r);
    }
  }
  e("}\n");
  if (min_blocks > 0)
    e("extern \"C\" __global__ void __launch_bounds__(256, %d) jit_kernel(%s) {\n", min_blocks, param_list(n_args, true).c_str());
  else
    e("extern \"C\" __global__ void __launch_bounds__(256) jit_kernel(%s) {\n", param_list(n_args, true).c_str());
  e("  const %s NV = %lld;\n  const %s nchunks = %lld;\n", IDX, (long long)NV, IDX, (long long)nchunks);
  e("  for (%s c = blockIdx.x; c < nchunks; c += gridDim.x) {\n", IDX);
  e("    const %s v0 = c * %lld + threadIdx.x;\n", IDX, (long long)chunk);
  e("    if ((c + 1) * %lld <= NV) {\n", (long long)chunk);
  e("      Regs r[%d];\n", U);
  e("      #pragma unroll\n      for (int u = 0; u < %d; ++u) ld(v0 + u * 256%s%s, r[u]);\n", U, n_args ? ", " : "", arg_pass(n_args).c_str());
  e("      #pragma unroll\n      for (int u = 0; u < %d; ++u) st(v0 + u * 256, r[u], out);\n", U);
  e("    } else {\n");
  e("      for (int u = 0; u < %d; ++u) {\n        const %s v = v0 + u * 256;\n        if (v < NV) { Regs r; ld(v%s%s, r); st(v, r, out); }\n      }\n", U, IDX,
    n_args ? ", " : "", arg_pass(n_args).c_str());
  e("    }\n  }\n");
  if (tail > 0) {
    e("  if (blockIdx.x == 0 && threadIdx.x < %lld) {\n    const %s i = (%s)%lld + threadIdx.x;\n", (long long)tail, IDX, IDX, (long long)(NV * V));
    for (int j = 0; j < nloads; ++j) e("    const float L%d[1] = {cc_ldg(p%d + i)};\n", j, p.loads[j].arg);
    emit_ops(e, p, "    ", "0");
    e("    out[i] = _%d;\n  }\n", p.results[0]);
  }
  e("}\n");
  plan.source += e.s;
  LaunchSpec ls;
  ls.entry = "jit_kernel";
  ls.grid[0] = (uint32_t)grid;
  ls.block[0] = 256;
  for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
  ls.args.push_back(ARG_OUT);
  if (total > 0) plan.launches.push_back(ls);
  (void)concat_fallback_n;
}

// ---- tiled-transpose elementwise kernel ---------------------------------------------------------------------------
//
// When a load is contiguous in the source along an output dimension d that is NOT the output's fastest dimension L
// (permute / transpose views, `join` of a split), a plain gather reads 4 bytes per 128-byte line. This template walks
// 32 x 32 tiles over (d, L): such loads are read with threads running along d (coalesced), staged through padded
// shared memory, and consumed with threads running along L, where every other load and the store are coalesced.
int transpose_dim(const Program& p) {
  const int nd = (int)p.dims.size();
  if (nd < 2 || p.results.size() != 1) return -1;
  const int L = nd - 1;
  int d = -1;
  for (const Load& ld : p.loads) {
    if (!ld.integer) return -1;
    if (ld.coef[L] == 0 || ld.coef[L] == 1 || ld.coef[L] == -1) continue;
    for (int x = 0; x < L; ++x)
      if (ld.coef[x] == 1 && p.dims[x] >= 32) {
        if (d < 0) d = x;
        break;
      }
  }
  if (d < 0 || p.dims[L] < 32) return -1;
  return d;
}

void emit_tiled_transpose(Plan& plan, const Program& p, int n_args, const DeviceProps& dev, int d) {
  const int nd = (int)p.dims.size();
  const int L = nd - 1;
  const int64_t total = product(p.dims);
  const char* IDX = pick_idx_type(p, total);
  const int nloads = (int)p.loads.size();
  std::vector<int> slot(nloads, -1);
  int nt = 0;
  for (int j = 0; j < nloads; ++j) {
    const Load& ld = p.loads[j];
    if (ld.coef[L] != 0 && ld.coef[L] != 1 && ld.coef[L] != -1 && ld.coef[d] == 1) slot[j] = nt++;
  }
  // 64 x 64 tiles (256-byte row segments on both sides, 16 independent loads in flight per thread) when the extents allow
  const int TS = (p.dims[d] >= 64 && p.dims[L] >= 64 && nt <= 3) ? 64 : 32;
  const int RY = 256 / TS;       // rows covered per pass
  const int NR = TS / RY;        // passes
  const int64_t tilesL = (p.dims[L] + TS - 1) / TS, tilesD = (p.dims[d] + TS - 1) / TS;
  int64_t outer = 1;
  for (int x = 0; x < nd; ++x)
    if (x != d && x != L) outer *= p.dims[x];
  const int64_t ntiles = tilesL * tilesD * outer;
  int64_t tgm = (int64_t)1 << 30;
  if (const char* ev = plan_knob("CC_TUNE_T_GRID_MULT")) tgm = std::max(1, atoi(ev));
  int64_t grid = std::min<int64_t>(ntiles, std::min<int64_t>((int64_t)dev.sm_count * tgm, 0x7fffffff));
  if (grid < 1) grid = 1;
  std::vector<int64_t> ostride(nd, 1);
  for (int x = nd - 2; x >= 0; --x) ostride[x] = ostride[x + 1] * p.dims[x + 1];

  Emit e;
  e("// tiled transpose: dims=[");
  for (int x = 0; x < nd; ++x) e("%s%lld", x ? "," : "", (long long)p.dims[x]);
  e("] tile %dx%d over (g%d, g%d), %d staged load(s) of %d, idx=%s grid=%lld\n", TS, TS, d, L, nt, nloads, IDX, (long long)grid);
  e("extern \"C\" __global__ void __launch_bounds__(256) jit_kernel(%s) {\n", param_list(n_args, true).c_str());
  e("  __shared__ float tile[%d][%d][%d];\n", nt, TS, TS + 1);
  e("  const int tx = threadIdx.x %% %d, ty = threadIdx.x / %d;\n", TS, TS);
  e("  for (%s t_ = blockIdx.x; t_ < (%s)%lld; t_ += gridDim.x) {\n", IDX, IDX, (long long)ntiles);
  e("    %s r_ = t_;\n    const %s tl = r_ %% (%s)%lld; r_ /= (%s)%lld;\n    const %s td = r_ %% (%s)%lld; r_ /= (%s)%lld;\n", IDX, IDX, IDX, (long long)tilesL, IDX,
    (long long)tilesL, IDX, IDX, (long long)tilesD, IDX, (long long)tilesD);
  for (int x = nd - 1; x >= 0; --x) {
    if (x == d || x == L) continue;
    e("    const %s g%d = r_ %% (%s)%lld; r_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
  }
  LoadCtx c{1, -1, IDX};
  // phase 1: threads run along d (the source-contiguous index of the staged loads)
  e("    #pragma unroll\n    for (int r = 0; r < %d; ++r) {\n", NR);
  e("      const %s g%d = td * %d + tx;\n      const %s g%d = tl * %d + ty + %d * r;\n", IDX, d, TS, IDX, L, TS, RY);
  e("      if (g%d < (%s)%lld && g%d < (%s)%lld) {\n", d, IDX, (long long)p.dims[d], L, IDX, (long long)p.dims[L]);
  for (int j = 0; j < nloads; ++j) {
    if (slot[j] < 0) continue;
    e("        {\n");
    emit_load(e, p, j, c, "          ");
    e("          tile[%d][ty + %d * r][tx] = L%d[0];\n        }\n", slot[j], RY, j);
  }
  e("      }\n    }\n    __syncthreads();\n");
  // phase 2: threads run along L (the output-contiguous index)
  e("    #pragma unroll\n    for (int r = 0; r < %d; ++r) {\n", NR);
  e("      const %s g%d = td * %d + ty + %d * r;\n      const %s g%d = tl * %d + tx;\n", IDX, d, TS, RY, IDX, L, TS);
  e("      if (g%d < (%s)%lld && g%d < (%s)%lld) {\n", d, IDX, (long long)p.dims[d], L, IDX, (long long)p.dims[L]);
  for (int j = 0; j < nloads; ++j) {
    if (slot[j] >= 0)
      e("        const float L%d[1] = {tile[%d][tx][ty + %d * r]};\n", j, slot[j], RY);
    else
      emit_load(e, p, j, c, "        ");
  }
  emit_ops(e, p, "        ", "0");
  std::string lin = "(long long)0";
  for (int x = 0; x < nd; ++x) lin += strprintf(" + (long long)g%d * %lldLL", x, (long long)ostride[x]);
  e("        __stcs(out + (%s), _%d);\n", lin.c_str(), p.results[0]);
  e("      }\n    }\n    __syncthreads();\n  }\n}\n");
  plan.source += e.s;
  LaunchSpec ls;
  ls.entry = "jit_kernel";
  ls.grid[0] = (uint32_t)grid;
  ls.block[0] = 256;
  for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
  ls.args.push_back(ARG_OUT);
  if (total > 0) plan.launches.push_back(ls);
}

// ---- stencil tile: dense windows over one source -------------------------------------------------------------------------
//
// An elementwise program that reads ONE same-shape source through many pure translations differing in the last two dimensions
// (box sums, max pooling, Laplacians: `x.translate(dy, dx)` for a window of (dy, dx)) re-reads every element once per window
// position through L1. Here a CTA stages the 16 x 128 output tile plus its halo into shared memory once -- bounds tests and the
// padding are applied while staging, with 128-bit loads -- and every window position is then a conflict-free shared-memory read.
// Other operands of the expression are loaded per output as in the elementwise template.
bool try_emit_stencil_tile(Plan& plan, const Program& p, int n_args, const DeviceProps& dev) {
  const int nd = (int)p.dims.size();
  if (plan_knob("CC_NO_STENCIL_TILE")) return false;
  if (nd < 2 || p.results.size() != 1) return false;
  const int64_t H = p.dims[nd - 2], W = p.dims[nd - 1];
  if (W % 4 != 0 || W < 128 || H < 8) return false;
  std::vector<int64_t> stride(nd, 1);
  for (int x = nd - 2; x >= 0; --x) stride[x] = stride[x + 1] * p.dims[x + 1];
  auto is_translation = [&](const Load& L) {
    if (!L.integer || (int)L.src_shape.size() != nd) return false;
    for (int x = 0; x < nd; ++x)
      if (L.src_shape[x] != p.dims[x] || L.coef[x] != stride[x]) return false;
    // the matrix itself must be the identity (coef == stride could also come from a degenerate mixture)
    for (int y = 0; y < nd; ++y)
      for (int x = 0; x < nd; ++x)
        if (L.M[(size_t)y * (nd + 1) + x] != (x == y ? 1.0 : 0.0)) return false;
    return true;
  };
  // the window source: the argument with the most translated loads
  std::vector<int> count((size_t)n_args, 0);
  for (const Load& L : p.loads)
    if (is_translation(L)) ++count[(size_t)L.arg];
  int wa = -1;
  for (int a = 0; a < n_args; ++a)
    if (count[(size_t)a] >= 6 && (wa < 0 || count[(size_t)a] > count[(size_t)wa])) wa = a;
  if (wa < 0) return false;
  const int nloads = (int)p.loads.size();
  std::vector<char> in_window((size_t)nloads, 0);
  int64_t ymin = 0, ymax = 0, xmin = 0, xmax = 0;
  std::vector<int64_t> lead_off;
  float padding = 0.f;
  bool first = true;
  for (int j = 0; j < nloads; ++j) {
    const Load& L = p.loads[j];
    if (L.arg != wa || !is_translation(L)) continue;
    std::vector<int64_t> off(nd);
    for (int y = 0; y < nd; ++y) off[y] = (int64_t)L.M[(size_t)y * (nd + 1) + nd];
    std::vector<int64_t> lo(off.begin(), off.end() - 2);
    if (first) {
      lead_off = lo;
      padding = L.padding;
      ymin = ymax = off[nd - 2];
      xmin = xmax = off[nd - 1];
      first = false;
    } else if (lo != lead_off || memcmp(&padding, &L.padding, 4) != 0) {
      continue;  // a translation along a leading dimension: read like any other operand
    }
    in_window[(size_t)j] = 1;
    ymin = std::min(ymin, off[nd - 2]), ymax = std::max(ymax, off[nd - 2]);
    xmin = std::min(xmin, off[nd - 1]), xmax = std::max(xmax, off[nd - 1]);
  }
  int nwin = 0;
  for (char c : in_window) nwin += c;
  // (from six views on: with the five of a 5-point stencil the L1-resident elementwise kernel already runs at the HBM rate)
  if (nwin < 6 || ymax - ymin > 16 || xmax - xmin > 32) return false;
  auto floor4 = [](int64_t v) { return v >= 0 ? v / 4 * 4 : -((-v + 3) / 4 * 4); };
  // 256 threads = 8 row groups x 32 column vectors; each thread owns RT rows of 4 columns -> tile of 8 * RT rows x 128 columns
  int RT = 4;
  if (const char* ev = plan_knob("CC_TUNE_STENCIL_RT")) RT = std::max(1, std::min(8, atoi(ev)));  // tuning knob
  const int TW = 128;
  const int64_t HX0 = floor4(xmin), HX1 = -floor4(-xmax);  // halo columns rounded outwards to the 16-byte grid
  const int SW = TW + (int)(HX1 - HX0);
  while (RT > 1 && 2 * (8 * RT + (int)(ymax - ymin)) * SW * 4 > 48 * 1024) RT /= 2;  // two tile buffers within the static shared-memory limit
  const int TH = 8 * RT;
  const int SH = TH + (int)(ymax - ymin);
  const int64_t total = product(p.dims);
  const char* IDX = pick_idx_type(p, total);
  int64_t lead = 1;
  for (int x = 0; x < nd - 2; ++x) lead *= p.dims[x];
  const int64_t tilesY = (H + TH - 1) / TH, tilesX = (W + TW - 1) / TW;
  const int64_t ntiles = lead * tilesY * tilesX;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ntiles, 0x7fffffff));
  Emit e;
  e("// stencil tile: dims=[");
  for (int x = 0; x < nd; ++x) e("%s%lld", x ? "," : "", (long long)p.dims[x]);
  e("] %d window loads of p%d, dy=[%lld,%lld] dx=[%lld,%lld], tile %dx%d + halo -> shared %dx%d, %d other loads, idx=%s grid=%lld\n", nwin, wa, (long long)ymin,
    (long long)ymax, (long long)xmin, (long long)xmax, TH, TW, SH, SW, nloads - nwin, IDX, (long long)std::min<int64_t>(ntiles, (int64_t)dev.sm_count * 2));
  // Persistent CTAs, two tile buffers: while a tile is being computed the next one is already on its way into the other buffer
  // (cp.async, 16 bytes each, L1 bypassed) — the copy engine-less equivalent of a TMA pipeline, with the halo's out-of-range vectors
  // (whole vectors: the tile origin, the halo and W are multiples of 4) written as padding by ordinary stores. Before round 2 a CTA
  // staged, synchronised, computed and exited: with 2 CTAs per SM (registers) the loads of one tile barely overlapped the arithmetic of
  // another (3 x 3: 0.67 of HBM, 5 x 5: 0.46).
  int ctas_per_sm = 2;
  if (const char* ev = plan_knob("CC_TUNE_STENCIL_CTAS")) ctas_per_sm = std::max(1, std::min(4, atoi(ev)));  // tuning knob
  const int64_t pgrid = std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)dev.sm_count * ctas_per_sm));
  e("extern \"C\" __global__ void __launch_bounds__(256, %d) jit_kernel(%s) {\n", ctas_per_sm, param_list(n_args, true).c_str());
  e("  __shared__ __align__(16) float tiles_[2][%d][%d];\n", SH, SW);
  const std::string pad = flit(padding);
  // stage(t, buffer): issue the copies of tile t
  e("  auto stage = [&](const %s t_, float (*tile)[%d]) {\n", IDX, SW);
  e("    %s r_ = t_;\n    const %s c0 = (r_ %% (%s)%lld) * %d; r_ /= (%s)%lld;\n    const %s r0 = (r_ %% (%s)%lld) * %d; r_ /= (%s)%lld;\n", IDX, IDX, IDX,
    (long long)tilesX, TW, IDX, (long long)tilesX, IDX, IDX, (long long)tilesY, TH, IDX, (long long)tilesY);
  for (int x = nd - 3; x >= 0; --x)
    e("    const %s g%d = r_ %% (%s)%lld; r_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
  // leading part of the window source's address, and whether the leading indices are in range at all
  std::string lead_ok = "true", lead_base = strprintf("(%s)0", IDX);
  for (int x = 0; x < nd - 2; ++x) {
    if (lead_off[(size_t)x] != 0)
      lead_ok += strprintf(" && g%d + %lld >= 0 && g%d + %lld < %lld", x, (long long)lead_off[(size_t)x], x, (long long)lead_off[(size_t)x], (long long)p.dims[x]);
    lead_base += strprintf(" + (g%d + (%s)%lld) * (%s)%lld", x, IDX, (long long)lead_off[(size_t)x], IDX, (long long)stride[x]);
  }
  e("    const bool lead_ok = %s;\n    const %s lead_base = %s;\n", lead_ok.c_str(), IDX, lead_base.c_str());
  // a thread keeps its column vector and walks down the rows (RPP rows per pass): one bounds test and two pointer bumps per copy, no
  // index division in the loop (the staging loop was a quarter of the 3 x 3 kernel's instructions)
  const int VPR = SW / 4, RPP = 256 / VPR;
  e("    const int sy0 = threadIdx.x / %d, sx = (threadIdx.x %% %d) * 4;\n", VPR, VPR);
  e("    const %s xx = c0 + sx + (%s)%lld;\n    const bool col_ok = lead_ok && xx >= 0 && xx < (%s)%lld;\n", IDX, IDX, (long long)HX0, IDX, (long long)W);
  e("    if (sy0 < %d) {\n      %s yy = r0 + sy0 + (%s)%lld;\n      const float* src = p%d + lead_base + yy * (%s)%lld + xx;\n      float* dst = &tile[sy0][sx];\n", RPP, IDX,
    IDX, (long long)ymin, wa, IDX, (long long)W);
  e("      #pragma unroll\n      for (int sy = sy0; sy < %d; sy += %d, yy += %d, src += (%s)%lld, dst += %d) {\n", SH, RPP, RPP, IDX, (long long)(RPP * W), RPP * SW);
  e("        if (col_ok && yy >= 0 && yy < (%s)%lld)\n          cc_cp_async16(dst, src);\n        else\n          *reinterpret_cast<float4*>(dst) = make_float4(%s, %s, %s, %s);\n", IDX,
    (long long)H, pad.c_str(), pad.c_str(), pad.c_str(), pad.c_str());
  e("      }\n    }\n    cc_cp_async_commit();\n  };\n");
  e("  int buf_ = 0;\n  if ((%s)blockIdx.x < (%s)%lld) stage((%s)blockIdx.x, tiles_[0]);\n", IDX, IDX, (long long)ntiles, IDX);
  e("  for (%s t_ = blockIdx.x; t_ < (%s)%lld; t_ += gridDim.x, buf_ ^= 1) {\n", IDX, IDX, (long long)ntiles);
  e("    float (*tile)[%d] = tiles_[buf_];\n", SW);
  e("    if (t_ + (%s)gridDim.x < (%s)%lld) {\n      stage(t_ + (%s)gridDim.x, tiles_[buf_ ^ 1]);\n      cc_cp_async_wait<1>();\n    } else {\n      cc_cp_async_wait<0>();\n    }\n    __syncthreads();\n",
    IDX, IDX, (long long)ntiles, IDX);
  e("    %s r_ = t_;\n    const %s c0 = (r_ %% (%s)%lld) * %d; r_ /= (%s)%lld;\n    const %s r0 = (r_ %% (%s)%lld) * %d; r_ /= (%s)%lld;\n", IDX, IDX, IDX,
    (long long)tilesX, TW, IDX, (long long)tilesX, IDX, IDX, (long long)tilesY, TH, IDX, (long long)tilesY);
  for (int x = nd - 3; x >= 0; --x)
    e("    const %s g%d = r_ %% (%s)%lld; r_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
  // ---- compute: a thread owns RT rows x 4 adjacent columns. It pulls the RT + (ymax - ymin) rows of covering aligned vectors out of
  // shared memory once (conflict-free 128-bit reads; a row is shared by the RT outputs above and below it, which is what takes the
  // shared-memory pipe off the critical path: ncu showed it 75 % busy with one row of outputs per thread) and every window position
  // is a static pick from those registers
  const int NC = 4 + (int)(HX1 - HX0);
  const int WY = (int)(ymax - ymin);
  e("    {\n      const int ly0 = (threadIdx.x / %d) * %d, lx = (threadIdx.x %% %d) * 4;\n", TW / 4, RT, TW / 4);
  e("      const %s g%d = c0 + lx;\n", IDX, nd - 1);
  e("      if (g%d < (%s)%lld && r0 + ly0 < (%s)%lld) {\n", nd - 1, IDX, (long long)W, IDX, (long long)H);
  // (reading only the columns some window position picks — the outer vectors shrunk to 8 or 4 bytes — was timed and is slower: lanes are 16
  // bytes apart, so a 4-byte read is a 4-way bank conflict and costs the same wavefronts as the whole vector: 3 x 3 0.92 -> 0.88 of HBM)
  for (int r = 0; r < RT + WY; ++r)
    e("        float R%d[%d];\n        #pragma unroll\n        for (int k = 0; k < %d; ++k) *reinterpret_cast<float4*>(&R%d[4 * k]) = *reinterpret_cast<const float4*>(&tile[ly0 + %d][lx + 4 * k]);\n",
      r, NC, NC / 4, r, r);
  LoadCtx c{4, nd - 1, IDX};
  std::string lin = strprintf("(%s)0", "long long");
  for (int x = 0; x < nd; ++x) lin += strprintf(" + (long long)g%d * %lldLL", x, (long long)stride[x]);
  for (int i = 0; i < RT; ++i) {
    e("        if (r0 + ly0 + %d < (%s)%lld) {\n          const %s g%d = r0 + ly0 + %d;\n", i, IDX, (long long)H, IDX, nd - 2, i);
    for (int j = 0; j < nloads; ++j) {
      if (in_window[(size_t)j]) {
        const Load& L = p.loads[j];
        const int64_t dy = (int64_t)L.M[(size_t)(nd - 2) * (nd + 1) + nd], dx = (int64_t)L.M[(size_t)(nd - 1) * (nd + 1) + nd];
        const int r = i + (int)(dy - ymin), cx = (int)(dx - HX0);
        e("          const float L%d[4] = {R%d[%d], R%d[%d], R%d[%d], R%d[%d]};\n", j, r, cx, r, cx + 1, r, cx + 2, r, cx + 3);
      } else {
        emit_load(e, p, j, c, "          ");
      }
    }
    e("          float o_[4];\n          #pragma unroll\n          for (int l = 0; l < 4; ++l) {\n");
    emit_ops(e, p, "            ", "l");
    e("            o_[l] = _%d;\n          }\n", p.results[0]);
    e("          cc_stg4(out + (%s), o_);\n        }\n", lin.c_str());
  }
  e("      }\n    }\n    __syncthreads();\n  }\n}\n");
  plan.source += e.s;
  LaunchSpec ls;
  ls.entry = "jit_kernel";
  ls.grid[0] = (uint32_t)pgrid;
  ls.block[0] = 256;
  for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
  ls.args.push_back(ARG_OUT);
  if (total > 0) plan.launches.push_back(ls);
  plan.note += "; dense window staged through shared memory (double-buffered cp.async)";
  return true;
}

// ---- reductions over the re-rolled index --------------------------------------------------------------------------

// out[g] = sum_t E(g, t).  dims = out dims + [T].
// out[g] = post(sum_t E(g, t)).  dims = out dims + the reduction dims (n_red of them, outermost first; t runs over their
// product in row-major order, which is the reference's left-to-right order of the chain).
void emit_reduce(Plan& plan, const Program& p, int n_args, const DeviceProps& dev) {
  const int nd = (int)p.dims.size();
  const int R = std::max(1, p.n_red);
  const int no = nd - R;
  std::vector<int64_t> odims(p.dims.begin(), p.dims.begin() + no);
  std::vector<int64_t> rdims(p.dims.begin() + no, p.dims.end());
  const int64_t T = product(rdims);
  const int64_t NOUT = product(odims);
  const char* IDX = pick_idx_type(p, std::max(NOUT, T));  // indices are output positions, reduction positions and source offsets
  const int nloads = (int)p.loads.size();
  const bool has_post = !p.post_ops.empty() && !p.trivial_post();
  auto in_post = [&](int j) { return j < (int)p.load_in_post.size() && p.load_in_post[j]; };
  // the fold: `+` stays a plain C `+` (it may contract with the term's multiply into an fma, as the reference's build allows);
  // min / max start from NaN, which fminf / fmaxf ignore, so a chain of NaNs still folds to NaN as the unrolled chain would
  const uint32_t mono = p.red_monoid;
  const char* ZERO = mono == K_PLUS ? "0.f" : (mono == K_TIMES ? "1.f" : "__int_as_float(0x7fc00000)");
  const char* MSTRUCT = mono == K_PLUS ? "cc_plus" : (mono == K_TIMES ? "cc_times" : (mono == K_MIN ? "cc_min_nan" : "cc_max_nan"));
  auto AP = [&](const std::string& a, const std::string& b) {
    switch (mono) {
      case K_TIMES: return "(" + a + " * " + b + ")";
      case K_MIN: return "fminf(" + a + ", " + b + ")";
      case K_MAX: return "fmaxf(" + a + ", " + b + ")";
      default: return "(" + a + " + " + b + ")";
    }
  };
  // choose the orientation: which index do neighbouring threads walk?
  bool any_t_contig = false, any_o_contig = false;
  for (int j = 0; j < nloads; ++j) {
    const Load& L = p.loads[j];
    if (!L.integer || in_post(j)) continue;
    if (std::llabs(L.coef[nd - 1]) == 1) any_t_contig = true;
    if (no >= 1 && std::llabs(L.coef[no - 1]) == 1) any_o_contig = true;
  }
  const bool rows = (any_t_contig && !any_o_contig) || no == 0 || NOUT < 64;
  std::string gs;  // ", g0, g1, ..." over the output dims
  for (int x = 0; x < no; ++x) gs += strprintf(", g%d", x);
  const std::string pass = (n_args ? ", " : "") + arg_pass(n_args);
  const std::string params = (n_args ? ", " : "") + param_list(n_args, false);
  Emit e;
  // decode of the flat reduction index into its digits g{no}..g{nd-1}
  auto emit_t_decode = [&](const char* src) {
    if (R == 1) {
      e("  const %s g%d = %s;\n", IDX, no, src);
      return;
    }
    e("  %s tr_ = %s;\n", IDX, src);
    for (int x = nd - 1; x > no; --x)
      e("  const %s g%d = tr_ %% (%s)%lld; tr_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
    e("  const %s g%d = tr_;\n", IDX, no);
  };
  // epilogue: acc[V] -> final values, in place
  auto emit_post_fn = [&](int V, int vdim) {
    if (!has_post) return;
    e("__device__ __forceinline__ void post(float (&acc)[%d]", V);
    for (int x = 0; x < no; ++x) e(", const %s g%d", IDX, x);
    e("%s) {\n", params.c_str());
    LoadCtx c{V, vdim, IDX};
    for (int j = 0; j < nloads; ++j)
      if (in_post(j)) emit_load(e, p, j, c, "  ");
    e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V);
    emit_op_list(e, p.post_ops, "    ", "l", {p.post_result}, "q", "acc[l]");
    e("    acc[l] = q%d;\n  }\n}\n", p.post_result);
  };
  std::string rd;
  for (int x = 0; x < R; ++x) rd += strprintf("%s%lld", x ? "x" : "", (long long)rdims[x]);
  if (!rows) {
    // --- column owner: a thread owns V adjacent outputs and walks t; T is split over blockIdx.y ----------------------
    const int V = (odims[no - 1] % 4 == 0) ? 4 : 1;
    const int64_t NV = NOUT / V;
    int64_t want = ((int64_t)dev.sm_count * 2048 * 2 + NV - 1) / NV;
    int64_t S = std::max<int64_t>(1, std::min<int64_t>(want, T / 32));
    S = std::min<int64_t>(S, 1024);
    if (NV * 2 >= (int64_t)dev.sm_count * 2048) S = 1;  // the outputs alone fill the machine: no partials round trip
    const int64_t TCH = (T + S - 1) / S;
    S = (T + TCH - 1) / TCH;
    // --- tile owner: small trailing output dimension (the filters of the reference's convolution benchmark, benchmarks.scala:463-556 at
    // :612-630's sizes; the 32 columns of its skinny matmuls). A thread owns ALL F outputs along it (F / 4 vectors), so a load that ignores
    // that dimension (the translated input) is issued once for F multiply-adds instead of once for 4, and when it runs along the
    // innermost reduction digit (the channels of an NHWC image) it is fetched as ONE 128-bit vector for four reduction steps. Same terms
    // in the same order per output as the column owner: bit-identical results.
    {
      const int64_t F = odims[no - 1];
      const int inner = nd - 1;
      const int NG = (int)(F / 4);
      std::vector<char> kvec((size_t)nloads, 0);
      // (CC_REDUCE_TILE_OWNER: 0 = never, 2 = whenever the shape allows, whatever the size -- the emulator tests; default: when the threads
      // alone fill the machine, which is decided before the column owner would split T over CTAs)
      const char* knob = plan_knob("CC_REDUCE_TILE_OWNER");
      const int mode = knob ? atoi(knob) : 1;
      bool ok = mode != 0 && V == 4 && F >= 4 && F <= 32 && T >= 4 && T <= 4096 && (mode == 2 || NOUT / F >= (int64_t)dev.sm_count * 128);
      bool any_invariant = false, any_kvec = false;
      const int cols = nd + 1;
      for (int j = 0; ok && j < nloads; ++j) {
        if (in_post(j)) continue;
        const Load& L = p.loads[j];
        if (!L.integer) {
          ok = false;
          break;
        }
        bool lane_free = L.coef[no - 1] == 0;
        for (int y = 0; y < L.rows; ++y)
          if (L.M[(size_t)y * cols + (no - 1)] != 0.0) lane_free = false;
        if (!lane_free) continue;
        any_invariant = true;
        bool kv = R >= 1 && p.dims[inner] % 4 == 0 && L.coef[inner] == 1 && L.base % 4 == 0;
        for (int x = 0; kv && x < nd; ++x)
          if (x != inner && L.coef[x] % 4 != 0) kv = false;
        for (int y = 0; kv && y < L.rows; ++y)
          if ((L.need_lo[y] || L.need_hi[y]) && L.M[(size_t)y * cols + inner] != 0.0) kv = false;
        if (kv) kvec[(size_t)j] = 1, any_kvec = true;
      }
      if (ok && ((NG >= 2 && any_invariant) || any_kvec)) {
        const int KV = any_kvec ? 4 : 1;
        // P positions along the next output dimension per thread: a load that ignores it (the weights) then serves P x 4 multiply-adds per
        // vector, and translated windows along it share their vectors between neighbouring positions (identical loads are merged by the
        // compiler). ncu on the 3 x 3 / depth 8 / batch 128 convolution: L1 write-back of the (uniform) weight vectors bounds P = 1.
        int P = 1;
        if (no >= 2 && any_kvec) {
          const char* pk = plan_knob("CC_TUNE_TILE_P");
          for (int cand : {4, 2})
            if (P == 1 && odims[no - 2] % cand == 0 && NG * cand <= 8 && (mode == 2 || NOUT / F / cand >= (int64_t)dev.sm_count * 128)) P = cand;
          if (pk) P = std::max(1, atoi(pk));
          if (odims[no - 2] % P != 0) P = 1;
        }
        const int64_t NT = NOUT / F / P;  // threads
        e("// axis reduction (tile owner): out dims=[");
        for (int x = 0; x < no; ++x) e("%s%lld", x ? "," : "", (long long)odims[x]);
        e("] T=%s F=%lld (%d vectors per thread) P=%d kvec=%d idx=%s epilogue=%d fold=%s\n", rd.c_str(), (long long)F, NG, P, KV, IDX, (int)has_post, kind_name(mono));
        std::string rgdecl, kvdecl;
        for (int x = no; x < nd; ++x) rgdecl += strprintf("%sconst %s g%d", x > no ? ", " : "", IDX, x);
        for (int j = 0; j < nloads; ++j)
          if (kvec[(size_t)j]) kvdecl += strprintf(", const float kv%d", j);
        // evt: the term at explicit reduction digits for the four lanes from output index g{no-1} on; k-vector loads arrive as values
        e("__device__ __forceinline__ void evt(%s", rgdecl.c_str());
        for (int x = 0; x < no; ++x) e(", const %s g%d", IDX, x);
        e("%s%s, float (&o)[4]) {\n", params.c_str(), kvdecl.c_str());
        LoadCtx c{4, no - 1, IDX};
        for (int j = 0; j < nloads; ++j) {
          if (in_post(j)) continue;
          if (kvec[(size_t)j])
            e("  float L%d[4];\n  L%d[0] = L%d[1] = L%d[2] = L%d[3] = kv%d;\n", j, j, j, j, j, j);
          else
            emit_load(e, p, j, c, "  ");
        }
        e("  #pragma unroll\n  for (int l = 0; l < 4; ++l) {\n");
        emit_op_list(e, p.ops, "    ", "l", p.results);
        e("    o[l] = _%d;\n  }\n}\n", p.results[0]);
        emit_post_fn(4, no - 1);
        e("extern \"C\" __global__ void __launch_bounds__(128) reduce_cols(%s) {\n", param_list(n_args, true, "dst").c_str());
        e("  const %s v = (%s)blockIdx.x * 128 + threadIdx.x;\n  if (v >= %lld) return;\n", IDX, IDX, (long long)NT);
        emit_decode(e, odims, no, IDX, strprintf("v * %lld", (long long)(F * P)).c_str(), "  ");  // (P divides the next dimension: a tile never wraps)
        if (P > 1) e("  const %s gb_ = g%d;\n", IDX, no - 2);
        e("  float acc[%d][4];\n  #pragma unroll\n  for (int q = 0; q < %d; ++q)\n    #pragma unroll\n    for (int l = 0; l < 4; ++l) acc[q][l] = %s;\n", NG * P, NG * P, ZERO);
        std::string ind = "  ";
        for (int x = no; x < nd; ++x) {
          const int step = x == inner ? KV : 1;
          if (T <= 96)
            e("%s#pragma unroll\n", ind.c_str());
          else if (x == inner)
            e("%s#pragma unroll %d\n", ind.c_str(), (int)std::max<int64_t>(1, std::min<int64_t>(8, p.dims[x]) / step));
          else
            e("%s#pragma unroll 1\n", ind.c_str());
          e("%sfor (%s g%d = 0; g%d < %lld; g%d += %d) {\n", ind.c_str(), IDX, x, x, (long long)p.dims[x], x, step);
          ind += "  ";
        }
        // (the k-vectors: in scope g{inner} is the first of the four reduction steps they cover; inside the position loop g{no-2} shadows
        // the tile's first position with the current one)
        LoadCtx ck{4, inner, IDX, true};
        std::string digits, outs, kvpass;
        for (int x = no; x < nd; ++x) digits += x == inner && KV > 1 ? strprintf("%sg%d + k", x > no ? ", " : "", x) : strprintf("%sg%d", x > no ? ", " : "", x);
        for (int x = 0; x + 1 < no; ++x) outs += strprintf(", g%d", x);
        outs += strprintf(", (%s)(4 * q)", IDX);
        for (int j = 0; j < nloads; ++j)
          if (kvec[(size_t)j]) kvpass += strprintf(", L%d[k]", j);
        auto open_positions = [&](std::string& in) {
          if (P == 1) return;
          e("%s#pragma unroll\n%sfor (int pp = 0; pp < %d; ++pp) {\n%s  const %s g%d = gb_ + pp;\n", in.c_str(), in.c_str(), P, in.c_str(), IDX, no - 2);
          in += "  ";
        };
        auto close_positions = [&](std::string& in) {
          if (P == 1) return;
          in.resize(in.size() - 2);
          e("%s}\n", in.c_str());
        };
        const char* ACC = P == 1 ? "acc[q]" : "acc[pp * %d + q]";
        const std::string accq = P == 1 ? std::string("acc[q]") : strprintf(ACC, NG);
        open_positions(ind);
        for (int j = 0; j < nloads; ++j)
          if (!in_post(j) && kvec[(size_t)j]) emit_load(e, p, j, ck, ind.c_str());
        e("%s#pragma unroll\n%sfor (int k = 0; k < %d; ++k) {\n", ind.c_str(), ind.c_str(), KV);
        e("%s  #pragma unroll\n%s  for (int q = 0; q < %d; ++q) {\n", ind.c_str(), ind.c_str(), NG);
        e("%s    float x[4];\n%s    evt(%s%s%s%s, x);\n", ind.c_str(), ind.c_str(), digits.c_str(), outs.c_str(), pass.c_str(), kvpass.c_str());
        e("%s    #pragma unroll\n%s    for (int l = 0; l < 4; ++l) %s[l] = %s;\n", ind.c_str(), ind.c_str(), accq.c_str(), AP(accq + "[l]", "x[l]").c_str());
        e("%s  }\n%s}\n", ind.c_str(), ind.c_str());
        close_positions(ind);
        for (int x = no; x < nd; ++x) {
          ind.resize(ind.size() - 2);
          e("%s}\n", ind.c_str());
        }
        open_positions(ind);
        e("%s#pragma unroll\n%sfor (int q = 0; q < %d; ++q) {\n", ind.c_str(), ind.c_str(), NG);
        if (has_post) e("%s  post(%s%s%s);\n", ind.c_str(), accq.c_str(), outs.c_str(), pass.c_str());
        if (P == 1)
          e("%s  cc_stg4(dst + v * %lld + 4 * q, acc[q]);\n%s}\n", ind.c_str(), (long long)F, ind.c_str());
        else
          e("%s  cc_stg4(dst + v * %lld + pp * %lld + 4 * q, %s);\n%s}\n", ind.c_str(), (long long)(F * P), (long long)F, accq.c_str(), ind.c_str());
        close_positions(ind);
        e("}\n");
        LaunchSpec ls;
        ls.entry = "reduce_cols";
        ls.grid[0] = (uint32_t)((NT + 127) / 128);
        ls.block[0] = 128;
        for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
        ls.args.push_back(ARG_OUT);
        plan.launches.push_back(ls);
        plan.note += "; a thread owns the whole trailing output dimension";
        plan.source += e.s;
        return;
      }
    }
    e("// axis reduction (column owner): out dims=[");
    for (int x = 0; x < no; ++x) e("%s%lld", x ? "," : "", (long long)odims[x]);
    e("] T=%s V=%d splits=%lld chunk=%lld idx=%s epilogue=%d fold=%s\n", rd.c_str(), V, (long long)S, (long long)TCH, IDX, (int)has_post, kind_name(mono));
    // evd: the term at explicit reduction digits; ev: the same from the flat reduction index (used when T is split)
    std::string rgs, rgdecl;
    for (int x = no; x < nd; ++x) {
      rgs += strprintf("%sg%d", x > no ? ", " : "", x);
      rgdecl += strprintf("%sconst %s g%d", x > no ? ", " : "", IDX, x);
    }
    e("__device__ __forceinline__ void evd(%s", rgdecl.c_str());
    for (int x = 0; x < no; ++x) e(", const %s g%d", IDX, x);
    e("%s, float (&o)[%d]) {\n", params.c_str(), V);
    LoadCtx c{V, no - 1, IDX};
    for (int j = 0; j < nloads; ++j)
      if (!in_post(j)) emit_load(e, p, j, c, "  ");
    e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V);
    emit_op_list(e, p.ops, "    ", "l", p.results);
    e("    o[l] = _%d;\n  }\n}\n", p.results[0]);
    e("__device__ __forceinline__ void ev(const %s t", IDX);
    for (int x = 0; x < no; ++x) e(", const %s g%d", IDX, x);
    e("%s, float (&o)[%d]) {\n", params.c_str(), V);
    emit_t_decode("t");
    e("  evd(%s%s%s, o);\n}\n", rgs.c_str(), gs.c_str(), pass.c_str());
    emit_post_fn(V, no - 1);
    // T split (S > 1): a CTA is 64 column vectors x 4 sub-splits; the sub-splits take interleaved rows and are combined through
    // shared memory in a fixed order, so the partials round trip through HBM is 4x smaller than with one split per CTA
    const int QS = 4;
    const int64_t Sb = S > 1 ? (S + QS - 1) / QS : 1;           // CTAs along y
    const int64_t TCHb = S > 1 ? (T + Sb - 1) / Sb : T;         // rows per CTA
    const int64_t Sy = S > 1 ? (T + TCHb - 1) / TCHb : 1;       // partials per output
    const int CW = S > 1 ? 64 : 256;                            // column vectors per CTA
    // (A register-tiled variant of this kernel — a thread owning P positions along the next-to-fastest output dimension, with a sliding
    // register window over the translated input — was built in round 1 and timed in round 2 (profiles/r02_knob_ab.json): 1.08x on the
    // 3x3 depth-8 convolution at P = 2, 0.82x at depth 16, 0.64x at P = 4 (registers 32 -> 61 -> 80, occupancy). Not a win; removed.)
    // The second stage runs inside reduce_cols: the last CTA to finish a block of columns (a self-resetting counter per blockIdx.x) folds
    // that block's partials, so the plan is ONE launch (timed in round 2, profiles/r02_knob_ab.json: 1.02-1.23x against the two-launch
    // plan, 16384x4096: 55.4 -> 45.2 us). CC_FUSE_COL_STAGE=0 brings the separate reduce_partials launch back.
    const int64_t gridx = (NV + CW - 1) / CW;
    bool fused = Sy > 1 && gridx <= kColCounters;
    if (const char* ev = plan_knob("CC_FUSE_COL_STAGE")) fused = fused && atoi(ev) != 0;
    // A partial sum of a sharded tensor (`shard.split(0).reduce(_ + _)`) is all-reduced when it is evaluated: when this kernel also writes the
    // final values (no separate second launch) the threads that hold them complete the all-reduce over the peer mailboxes themselves — the
    // compute step and its collective are ONE launch (`mb_` is null on ordinary launches).
    const bool coll = mono == K_PLUS && !has_post && V == 4 && NOUT <= 65536 && (Sy == 1 || fused);
    const char* COLL_PARAMS = coll ? ", const cc_peer_mailboxes* __restrict__ mb_, const unsigned long long epoch_" : "";
    e("extern \"C\" __global__ void __launch_bounds__(256) reduce_cols(%s%s%s) {\n", param_list(n_args, true, "dst").c_str(),
      fused ? ", float* __restrict__ out, unsigned* __restrict__ counters" : "", COLL_PARAMS);
    if (S == 1) {
      e("  const %s v = (%s)blockIdx.x * 256 + threadIdx.x;\n  if (v >= %lld) return;\n", IDX, IDX, (long long)NV);
    } else {
      e("  __shared__ float comb[%d][%d][%d];\n", QS - 1, CW, V);
      e("  const int cx = threadIdx.x %% %d, qy = threadIdx.x / %d;\n", CW, CW);
      e("  const %s v_ = (%s)blockIdx.x * %d + cx;\n  const bool live = v_ < %lld;\n  const %s v = live ? v_ : 0;\n", IDX, IDX, CW, (long long)NV, IDX);
    }
    emit_decode(e, odims, no, IDX, strprintf("v * %d", V).c_str(), "  ");
    if (S == 1) {
      // one thread folds the whole chain: nested loops over the reduction digits (bounds tests and address terms of the outer
      // digits hoist out of the inner loop), left to right = the reference's order; `0 + e_0` is exact
      e("  float acc[%d];\n  #pragma unroll\n  for (int l = 0; l < %d; ++l) acc[l] = %s;\n", V, V, ZERO);
      std::string ind = "  ";
      for (int x = no; x < nd; ++x) {
        if (T <= 96)
          e("%s#pragma unroll\n", ind.c_str());
        else if (x == nd - 1)
          e("%s#pragma unroll %d\n", ind.c_str(), (int)std::min<int64_t>(8, p.dims[x]));
        else
          e("%s#pragma unroll 1\n", ind.c_str());
        e("%sfor (%s g%d = 0; g%d < %lld; ++g%d) {\n", ind.c_str(), IDX, x, x, (long long)p.dims[x], x);
        ind += "  ";
      }
      e("%sfloat x[%d];\n%sevd(%s%s%s, x);\n", ind.c_str(), V, ind.c_str(), rgs.c_str(), gs.c_str(), pass.c_str());
      e("%s#pragma unroll\n%sfor (int l = 0; l < %d; ++l) acc[l] = %s;\n", ind.c_str(), ind.c_str(), V, AP("acc[l]", "x[l]").c_str());
      for (int x = no; x < nd; ++x) {
        ind.resize(ind.size() - 2);
        e("%s}\n", ind.c_str());
      }
      if (has_post) e("  post(acc%s%s);\n", gs.c_str(), pass.c_str());
      e("  float* d = dst + v * %d;\n", V);
    } else {
      e("  const int t0 = blockIdx.y * %lld;\n  const int t1 = min(%lld, t0 + %lld);\n", (long long)TCHb, (long long)T, (long long)TCHb);
      e("  float acc[%d];\n  #pragma unroll\n  for (int l = 0; l < %d; ++l) acc[l] = %s;\n", V, V, ZERO);
      e("  if (live) {\n    #pragma unroll 4\n    for (int t = t0 + qy; t < t1; t += %d) {\n      float x[%d];\n      ev(t%s%s, x);\n", QS, V, gs.c_str(), pass.c_str());
      e("      #pragma unroll\n      for (int l = 0; l < %d; ++l) acc[l] = %s;\n    }\n  }\n", V, AP("acc[l]", "x[l]").c_str());
      e("  if (qy > 0) {\n    #pragma unroll\n    for (int l = 0; l < %d; ++l) comb[qy - 1][cx][l] = acc[l];\n  }\n  __syncthreads();\n", V);
      const std::string combine = strprintf("  #pragma unroll\n  for (int q = 0; q < %d; ++q)\n    #pragma unroll\n    for (int l = 0; l < %d; ++l) acc[l] = %s;\n", QS - 1, V,
                                            AP("acc[l]", "comb[q][cx][l]").c_str());
      if (!fused) {
        e("  if (qy > 0 || !live) return;\n");
        e("%s", combine.c_str());
        if (Sy == 1 && has_post) e("  post(acc%s%s);\n", gs.c_str(), pass.c_str());  // the sub-splits of one CTA covered all of T
        e("  float* d = dst + (%s)blockIdx.y * %lld + v * %d;\n", "long long", (long long)NOUT, V);
      } else {
        // stage 1: this CTA's partial (plain stores: they are read back through L2 by another SM in a moment)
        e("  if (qy == 0 && live) {\n%s", combine.c_str());
        e("    float* d = dst + (long long)blockIdx.y * %lld + v * %d;\n    #pragma unroll\n    for (int l = 0; l < %d; ++l) d[l] = acc[l];\n  }\n", (long long)NOUT, V, V);
        e("  __threadfence();\n  __syncthreads();\n  __shared__ unsigned last_;\n");
        e("  if (threadIdx.x == 0) last_ = atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1 ? 1u : 0u;\n  __syncthreads();\n  if (!last_) return;\n  __threadfence();\n");
        // stage 2 (last CTA of this column block): the %lld partials, sub-split s takes partials s, s + QS, ...; combined in a fixed order
        e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) acc[l] = %s;\n", V, ZERO);
        e("  if (live) {\n    #pragma unroll 4\n    for (int s = qy; s < %lld; s += %d) {\n      const float* q_ = dst + (long long)s * %lld + v * %d;\n      float x[%d];\n", (long long)Sy, QS,
          (long long)NOUT, V, V);
        if (V == 4)
          e("      { const float4 t_ = __ldcg(reinterpret_cast<const float4*>(q_)); x[0] = t_.x; x[1] = t_.y; x[2] = t_.z; x[3] = t_.w; }\n");
        else
          e("      x[0] = __ldcg(q_);\n");
        e("      #pragma unroll\n      for (int l = 0; l < %d; ++l) acc[l] = %s;\n    }\n  }\n", V, AP("acc[l]", "x[l]").c_str());
        e("  if (qy > 0) {\n    #pragma unroll\n    for (int l = 0; l < %d; ++l) comb[qy - 1][cx][l] = acc[l];\n  }\n  __syncthreads();\n", V);
        e("  if (threadIdx.x == 0) counters[blockIdx.x] = 0u;  // ready for the next launch on this stream\n");
        e("  if (qy > 0 || !live) return;\n%s", combine.c_str());
        if (has_post) e("  post(acc%s%s);\n", gs.c_str(), pass.c_str());
        e("  float* d = out + v * %d;\n", V);
      }
    }
    if (coll) e("  if (mb_) cc_ll_allreduce4(acc, (unsigned long long)v * 4, mb_, (unsigned)epoch_);\n");
    if (V == 4)
      e("  cc_stg4(d, acc);\n");
    else
      e("  d[0] = acc[0];\n");
    e("}\n");
    LaunchSpec ls;
    ls.entry = "reduce_cols";
    ls.grid[0] = (uint32_t)((NV + CW - 1) / CW);
    ls.grid[1] = (uint32_t)Sy;
    ls.block[0] = 256;
    for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
    ls.args.push_back(Sy == 1 ? ARG_OUT : ARG_SCRATCH0);
    if (fused) {
      ls.args.push_back(ARG_OUT);
      ls.args.push_back(ARG_COL_COUNTERS);
      plan.scratch_floats.push_back((uint64_t)(Sy * NOUT));
      plan.note += "; second stage fused into reduce_cols (last CTA per column block)";
    }
    if (coll) {
      ls.args.push_back(ARG_PEER_MB);
      ls.args.push_back(ARG_PEER_EPOCH);
      plan.collective = 1;
    }
    plan.launches.push_back(ls);
    S = fused ? 1 : Sy;  // what the second-stage kernel folds
    if (S > 1) {
      plan.scratch_floats.push_back((uint64_t)(S * NOUT));
      e("extern \"C\" __global__ void __launch_bounds__(256) reduce_partials(const float* __restrict__ part, float* __restrict__ out%s) {\n",
        has_post ? params.c_str() : "");
      e("  const long long v = (long long)blockIdx.x * 256 + threadIdx.x;\n  if (v >= %lld) return;\n", (long long)NV);
      e("  float acc[%d];\n", V);
      if (V == 4) {
        e("  cc_ldg4(part + v * 4, acc);\n  #pragma unroll 8\n  for (int s = 1; s < %lld; ++s) {\n    float x[4];\n    cc_ldg4(part + (long long)s * %lld + v * 4, x);\n", (long long)S,
          (long long)NOUT);
        e("    #pragma unroll\n    for (int l = 0; l < 4; ++l) acc[l] = %s;\n  }\n", AP("acc[l]", "x[l]").c_str());
      } else {
        e("  acc[0] = part[v];\n  #pragma unroll 8\n  for (int s = 1; s < %lld; ++s) acc[0] = %s;\n", (long long)S,
          AP("acc[0]", strprintf("part[(long long)s * %lld + v]", (long long)NOUT)).c_str());
      }
      if (has_post) {
        emit_decode(e, odims, no, IDX, strprintf("(%s)v * %d", IDX, V).c_str(), "  ");
        e("  post(acc%s%s);\n", gs.c_str(), pass.c_str());
      }
      if (V == 4)
        e("  cc_stg4(out + v * 4, acc);\n}\n");
      else
        e("  out[v] = acc[0];\n}\n");
      LaunchSpec l2;
      l2.entry = "reduce_partials";
      l2.grid[0] = (uint32_t)((NV + 255) / 256);
      l2.block[0] = 256;
      l2.args = {ARG_SCRATCH0, ARG_OUT};
      if (has_post)
        for (int i = 0; i < n_args; ++i) l2.args.push_back(i);
      plan.launches.push_back(l2);
    }
  } else {
    // --- row owner: G threads share one output and stride over t in 128-bit vectors -----------------------------------
    const int V = (p.dims[nd - 1] % 4 == 0) ? 4 : 1;
    const int64_t TV = T / V;
    // one warp per output while a lane gets <= 16 vectors and the outputs alone fill the machine (256x512x1024 summed over
    // its last axis: 0.73 of HBM with 256 threads holding one vector each); a whole CTA per output for long rows
    const int G = (TV <= 64 || (TV <= 512 && NOUT >= (int64_t)dev.sm_count * 64)) ? 32 : 256;
    const int OPB = 256 / G;  // outputs per block
    e("// axis reduction (row owner): out dims=[");
    for (int x = 0; x < no; ++x) e("%s%lld", x ? "," : "", (long long)odims[x]);
    e("] T=%s V=%d threads/output=%d idx=%s epilogue=%d fold=%s\n", rd.c_str(), V, G, IDX, (int)has_post, kind_name(mono));
    e("__device__ __forceinline__ float ev(const %s t", IDX);
    for (int x = 0; x < no; ++x) e(", const %s g%d", IDX, x);
    e("%s) {\n", params.c_str());
    emit_t_decode("t");
    LoadCtx c{V, nd - 1, IDX};
    for (int j = 0; j < nloads; ++j)
      if (!in_post(j)) emit_load(e, p, j, c, "  ");
    e("  float o[%d];\n  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V, V);
    emit_op_list(e, p.ops, "    ", "l", p.results);
    e("    o[l] = _%d;\n  }\n", p.results[0]);
    if (V == 4)
      e("  return %s;\n}\n", AP(AP("o[0]", "o[1]"), AP("o[2]", "o[3]")).c_str());
    else
      e("  return o[0];\n}\n");
    emit_post_fn(1, -1);
    // the row sums of a sharded tensor stay a row block; when every rank wants all of them (`gather`) lane 0 of every output pushes its value
    // into all ranks' mailboxes and collects the other ranks' — the kernel writes the gathered vector ([ranks x outputs]) itself
    const bool collg = !has_post && NOUT <= 65536;
    e("extern \"C\" __global__ void __launch_bounds__(256) reduce_rows(%s%s) {\n", param_list(n_args, true).c_str(),
      collg ? ", const cc_peer_mailboxes* __restrict__ mb_, const unsigned long long epoch_" : "");
    e("  const int lane = threadIdx.x %% %d;\n", G);
    e("  const %s oidx = (%s)blockIdx.x * %d + threadIdx.x / %d;\n", IDX, IDX, OPB, G);
    e("  const bool live = oidx < %lld;\n  const %s oc = live ? oidx : 0;\n", (long long)NOUT, IDX);
    emit_decode(e, odims, no, IDX, "oc", "  ");
    e("  float acc = %s;\n  #pragma unroll 4\n  for (int tv = lane; tv < %lld; tv += %d) acc = %s;\n", ZERO, (long long)TV, G,
      AP("acc", strprintf("ev((%s)tv * %d%s%s)", IDX, V, gs.c_str(), pass.c_str())).c_str());
    if (mono == K_PLUS) {
      e(G == 32 ? "  acc = cc_warp_sum(acc);\n" : "  acc = cc_block_sum_256(acc);\n");
    } else if (G == 32) {
      e("  acc = cc_warp_fold<%s>(acc);\n", MSTRUCT);
    } else {
      e("  __shared__ float red_[32];\n  acc = cc_block_fold<%s>(acc, red_);\n", MSTRUCT);
    }
    if (has_post)
      e("  if (lane == 0 && live) {\n    float a1[1] = {acc};\n    post(a1%s%s);\n    out[oidx] = a1[0];\n  }\n}\n", gs.c_str(), pass.c_str());
    else if (collg)
      e("  if (lane == 0 && live) {\n    if (mb_) cc_ll_allgather1(acc, (unsigned long long)oidx, %lldull, out, mb_, (unsigned)epoch_);\n    else out[oidx] = acc;\n  }\n}\n",
        (long long)NOUT);
    else
      e("  if (lane == 0 && live) out[oidx] = acc;\n}\n");
    LaunchSpec ls;
    ls.entry = "reduce_rows";
    ls.grid[0] = (uint32_t)((NOUT + OPB - 1) / OPB);
    ls.block[0] = 256;
    for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
    ls.args.push_back(ARG_OUT);
    if (collg) {
      ls.args.push_back(ARG_PEER_MB);
      ls.args.push_back(ARG_PEER_EPOCH);
      plan.collective = 2;
    }
    plan.launches.push_back(ls);
  }
  plan.source += e.s;
}

// ---- general contraction: operand panels gathered by generated kernels -----------------------------------------------------
//
// A re-rolled reduction whose term is `load_a * load_b`, where load_a ignores the trailing output dims (the "N" dims) and load_b
// ignores the leading ones (the "M" dims), is a GEMM C[M, N] = A[M, K] * B[K, N] over gathered operands: A[m, k] = load_a at
// (m's output indices, k's reduction digits), B^T[n, k] likewise. The convolution of benchmarks.scala:463-556 is the case in point
// (M = batch x height x width pixels, N = filters, K = kernel row x kernel column x channel, load_a = the translated, zero-padded
// input: an implicit im2col). Two generated kernels write the K-major TF32 hi / lo panels the tcgen05 pipeline consumes straight
// from the affine maps (bounds tests and padding included), the pipeline runs on them, and a third generated kernel applies the
// epilogue (e.g. `bias +`) in place.
void emit_panel_kernel(Emit& e, const Program& p, const char* name, int load, const std::vector<int>& row_dims, int n_args) {
  const int nd = (int)p.dims.size();
  const int R = p.n_red, no = nd - R;
  int64_t rows = 1, K = 1;
  for (int x : row_dims) rows *= p.dims[x];
  for (int x = no; x < nd; ++x) K *= p.dims[x];
  const int64_t Kp = (K + 31) / 32 * 32;
  const char* IDX = pick_idx_type(p, std::max(rows * Kp, (int64_t)1));
  e("// operand panel %s: rows=%lld K=%lld (padded %lld), load %d\n", name, (long long)rows, (long long)K, (long long)Kp, load);
  e("extern \"C\" __global__ void __launch_bounds__(256) %s(%s%sfloat* __restrict__ hi, float* __restrict__ lo) {\n", name, param_list(n_args, false).c_str(),
    n_args ? ", " : "");
  e("  const %s v = (%s)blockIdx.x * 256 + threadIdx.x;\n  if (v >= (%s)%lld) return;\n", IDX, IDX, IDX, (long long)(rows * (Kp / 4)));
  e("  const %s row = v / %lld;\n  const %s k0 = (v - row * %lld) * 4;\n", IDX, (long long)(Kp / 4), IDX, (long long)(Kp / 4));
  // output indices of this row
  e("  %s rr_ = row;\n", IDX);
  for (size_t i = row_dims.size(); i-- > 0;) {
    const int x = row_dims[i];
    if (i == 0)
      e("  const %s g%d = rr_;\n", IDX, x);
    else
      e("  const %s g%d = rr_ %% (%s)%lld; rr_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
  }
  const bool vec = p.dims[nd - 1] % 4 == 0;  // the 4 k's of a thread share every digit but the innermost: one decode, one (vector) load
  if (vec) {
    e("  float x[4] = {0.f, 0.f, 0.f, 0.f};\n  if (k0 < (%s)%lld) {\n    %s kr_ = k0;\n", IDX, (long long)K, IDX);
    for (int x = nd - 1; x >= no; --x) {
      if (x == no)
        e("    const %s g%d = kr_;\n", IDX, x);
      else
        e("    const %s g%d = kr_ %% (%s)%lld; kr_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
    }
    LoadCtx c{4, nd - 1, IDX};
    emit_load(e, p, load, c, "    ");
    e("    #pragma unroll\n    for (int q = 0; q < 4; ++q) x[q] = L%d[q];\n  }\n", load);
    e("  float hv[4], lv[4];\n  #pragma unroll\n  for (int q = 0; q < 4; ++q) cc_split_tf32(x[q], hv[q], lv[q]);\n");
  } else {
    e("  float hv[4], lv[4];\n  #pragma unroll\n  for (int q = 0; q < 4; ++q) {\n    const %s k = k0 + q;\n    float x = 0.f;\n    if (k < (%s)%lld) {\n", IDX,
      IDX, (long long)K);
    e("      %s kr_ = k;\n", IDX);
    for (int x = nd - 1; x >= no; --x) {
      if (x == no)
        e("      const %s g%d = kr_;\n", IDX, x);
      else
        e("      const %s g%d = kr_ %% (%s)%lld; kr_ /= (%s)%lld;\n", IDX, x, IDX, (long long)p.dims[x], IDX, (long long)p.dims[x]);
    }
    LoadCtx c{1, -1, IDX};
    emit_load(e, p, load, c, "      ");
    e("      x = L%d[0];\n    }\n    cc_split_tf32(x, hv[q], lv[q]);\n  }\n", load);
  }
  e("  cc_stg4(hi + v * 4, hv);\n  cc_stg4(lo + v * 4, lv);\n}\n");
}

void emit_post_kernel(Emit& e, const Program& p, int n_args) {
  const int nd = (int)p.dims.size();
  const int no = nd - p.n_red;
  std::vector<int64_t> odims(p.dims.begin(), p.dims.begin() + no);
  const int64_t NOUT = product(odims);
  const int V = (no >= 1 && odims[no - 1] % 4 == 0) ? 4 : 1;
  const char* IDX = pick_idx_type(p, NOUT);
  e("// epilogue applied in place to the contraction's result\n");
  e("extern \"C\" __global__ void __launch_bounds__(256) post_kernel(%s%sfloat* __restrict__ out) {\n", param_list(n_args, false).c_str(), n_args ? ", " : "");
  e("  const %s v = (%s)blockIdx.x * 256 + threadIdx.x;\n  if (v >= (%s)%lld) return;\n", IDX, IDX, IDX, (long long)(NOUT / V));
  emit_decode(e, odims, no, IDX, strprintf("v * %d", V).c_str(), "  ");
  e("  float acc[%d];\n", V);
  if (V == 4)
    e("  { const float4 t_ = *reinterpret_cast<const float4*>(out + v * 4); acc[0] = t_.x; acc[1] = t_.y; acc[2] = t_.z; acc[3] = t_.w; }\n");
  else
    e("  acc[0] = out[v];\n");
  LoadCtx c{V, no - 1, IDX};
  for (size_t j = 0; j < p.loads.size(); ++j)
    if (j < p.load_in_post.size() && p.load_in_post[j]) emit_load(e, p, (int)j, c, "  ");
  e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V);
  emit_op_list(e, p.post_ops, "    ", "l", {p.post_result}, "q", "acc[l]");
  e("    acc[l] = q%d;\n  }\n", p.post_result);
  if (V == 4)
    e("  cc_stg4(out + v * 4, acc);\n}\n");
  else
    e("  out[v] = acc[0];\n}\n");
}

// ---- small-N contraction on warp-level tensor-core MMAs --------------------------------------------------------------------
//
// The reference's own convolution benchmark (benchmarks.scala:463-556 at :612-630's sizes: 128 x 32 x 32 pixels, 3 x 3 x 8 taps, 8 filters)
// and its skinny products are contractions with a tiny N (8 - 32 outputs per row) and a short K (<= 256): far too small for operand panels
// and the tcgen05 pipeline, and on the FMA pipe they are bound by operand delivery, not arithmetic — ncu on the generated reduction: L1
// write-back of the weight vectors 59 % busy, FMA pipe 17 %, every output re-reading all K weights. Here a warp keeps ALL of B (hi / lo TF32
// fragments for every k step and n tile) in registers for its whole life and streams rows of A through `mma.sync.m16n8k8.tf32` (3xTF32:
// a_lo b_hi + a_hi b_lo + a_hi b_hi, as the large contraction does), so A is read once — straight through its affine map, bounds tests and
// padding included: an implicit im2col — and B once per warp. K is permuted inside a k step (thread t holds k = 2t, 2t + 1 of both
// operands) so a thread's two A elements of a row are adjacent in memory.
// (B must fit in a warp's registers — k steps x n tiles x 4 registers <= 96 — or, as ready-made fragments, in 48 KB of shared memory; enough
// rows to fill the SMs with warps that each amortise loading it)
bool small_n_mma_fits(int64_t M, int64_t N, int64_t K) {
  if (const char* ev = plan_knob("CC_SMALL_N_MMA"))
    if (atoi(ev) == 0) return false;
  const int64_t KS = (K + 7) / 8, NT = (N + 7) / 8;
  // (odd N — the depth-3 convolutions, N = 3 and K = 27 — was timed too: 4.4 -> 7.9 us, the padded tiles and the runtime index divisions cost
  // more than the generic reduction's 27 multiply-adds per output)
  return N >= 4 && N % 2 == 0 && N <= 32 && K >= 8 && K <= 256 && KS * NT <= 96 && M >= 4096 && M * N * K >= ((int64_t)1 << 20) && M < ((int64_t)1 << 31) - 16;
}

bool try_small_n_mma(Plan& plan, const Program& p, int n_args, int la, int lb, int s, int64_t M, int64_t N, int64_t K) {
  if (!small_n_mma_fits(M, N, K)) return false;
  const int nd = (int)p.dims.size();
  const int R = p.n_red, no = nd - R;
  const int64_t KS = (K + 7) / 8, NT = (N + 7) / 8;
  if (!p.loads[la].integer || !p.loads[lb].integer) return false;
  const bool has_post = !p.post_ops.empty() && !p.trivial_post();
  const char* IDX = pick_idx_type(p, std::max(M * N, M * K));
  const std::string params = (n_args ? ", " : "") + param_list(n_args, false);
  const std::string pass = (n_args ? ", " : "") + arg_pass(n_args);
  Emit e;
  e("// small-N contraction %lldx%lldx%lld on warp-level MMAs (m16n8k8 TF32, 3xTF32): B in registers (%lld k steps x %lld n tiles), rows of A streamed "
    "through their affine map; idx=%s epilogue=%d\n", (long long)M, (long long)N, (long long)K, (long long)KS, (long long)NT, IDX, (int)has_post);
  // (from 25 fragments on B lives in shared memory instead, built once per CTA)
  auto decode = [&](const char* src, int first, int last, const char* indent) {  // flat index -> g{first..last-1} (row-major)
    e("%s%s r_%d = %s;\n", indent, IDX, first, src);
    for (int x = last - 1; x > first; --x)
      e("%sconst %s g%d = r_%d %% (%s)%lld; r_%d /= (%s)%lld;\n", indent, IDX, x, first, IDX, (long long)p.dims[x], first, IDX, (long long)p.dims[x]);
    e("%sconst %s g%d = r_%d;\n", indent, IDX, first, first);
  };
  LoadCtx c1{1, -1, IDX};
  // a thread's two k of a step (2t, 2t + 1) are neighbours along the innermost reduction digit: one 8-byte load when A runs along it
  // (with ONE n tile a k step is 3 MMAs, too few to cover a load: the kernel is latency-bound and four 4-byte loads in flight per step beat
  // two 8-byte ones — 3 x 3 / depth 8: 6.7 vs 7.9 us; from two n tiles on it is instruction-bound and the pair wins — 64 x 64 x 64 depth 16:
  // 45.9 -> 31.5 us, 65536 x 32 x 32: 7.3 -> 6.4)
  bool a_pair = NT >= 2 && K % 2 == 0 && p.dims[nd - 1] % 2 == 0 && p.loads[la].coef[nd - 1] == 1 && p.loads[la].base % 2 == 0;
  if (const char* ev = plan_knob("CC_TUNE_SMALL_N_PAIR")) a_pair = a_pair && atoi(ev) != 0;  // A/B knob
  {
    const Load& L = p.loads[la];
    for (int x = 0; a_pair && x < nd - 1; ++x)
      if (L.coef[x] % 2 != 0) a_pair = false;
    for (int y = 0; a_pair && y < L.rows; ++y)
      if ((L.need_lo[y] || L.need_hi[y]) && L.M[(size_t)y * (nd + 1) + (nd - 1)] != 0.0) a_pair = false;
  }
  if (a_pair) {
    LoadCtx c2{1, -1, IDX};
    c2.pair = true;
    e("__device__ __forceinline__ float2 ld_a2(const %s m, const %s k%s) {\n  if (m >= (%s)%lld || k >= (%s)%lld) return make_float2(0.f, 0.f);\n", IDX, IDX, params.c_str(),
      IDX, (long long)M, IDX, (long long)K);
    decode("m", 0, s, "  ");
    decode("k", no, nd, "  ");
    emit_load(e, p, la, c2, "  ");
    e("  return make_float2(L%d[0], L%d[1]);\n}\n", la, la);
  } else {
    e("__device__ __forceinline__ float ld_a(const %s m, const %s k%s) {\n  if (m >= (%s)%lld || k >= (%s)%lld) return 0.f;\n", IDX, IDX, params.c_str(), IDX,
      (long long)M, IDX, (long long)K);
    decode("m", 0, s, "  ");
    decode("k", no, nd, "  ");
    emit_load(e, p, la, c1, "  ");
    e("  return L%d[0];\n}\n", la);
  }
  e("__device__ __forceinline__ float ld_b(const %s n, const %s k%s) {\n  if (n >= (%s)%lld || k >= (%s)%lld) return 0.f;\n", IDX, IDX, params.c_str(), IDX,
    (long long)N, IDX, (long long)K);
  decode("n", s, no, "  ");
  decode("k", no, nd, "  ");
  emit_load(e, p, lb, c1, "  ");
  e("  return L%d[0];\n}\n", lb);
  if (has_post) {
    e("__device__ __forceinline__ float post1(const float acc0, const %s m, const %s n%s) {\n", IDX, IDX, params.c_str());
    decode("m", 0, s, "  ");
    decode("n", s, no, "  ");
    for (size_t j = 0; j < p.loads.size(); ++j)
      if (j < p.load_in_post.size() && p.load_in_post[j]) emit_load(e, p, (int)j, c1, "  ");
    emit_op_list(e, p.post_ops, "  ", "0", {p.post_result}, "q", "acc0");
    e("  return q%d;\n}\n", p.post_result);
  }
  const int64_t MT = (M + 15) / 16;
  e("extern \"C\" __global__ void __launch_bounds__(128) small_n_mma(%s) {\n", param_list(n_args, true, "dst").c_str());
  e("  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;\n");
  const bool b_shared = KS * NT > 24;
  if (!b_shared) {
    e("  unsigned bh[%lld][%lld][2], bl[%lld][%lld][2];\n", (long long)KS, (long long)NT, (long long)KS, (long long)NT);
    e("  #pragma unroll\n  for (int ks = 0; ks < %lld; ++ks)\n    #pragma unroll\n    for (int nt = 0; nt < %lld; ++nt)\n      #pragma unroll\n      for (int r = 0; r < 2; ++r) {\n",
      (long long)KS, (long long)NT);
    e("        float h, l;\n        cc_split_tf32(ld_b((%s)(nt * 8 + gid), (%s)(ks * 8 + 2 * tig + r)%s), h, l);\n", IDX, IDX, pass.c_str());
    e("        bh[ks][nt][r] = __float_as_uint(h);\n        bl[ks][nt][r] = __float_as_uint(l);\n      }\n");
  } else {
    // B too large for registers: the CTA builds every lane's fragments {hi0, hi1, lo0, lo1} once, in shared memory (conflict-free 128-bit reads)
    e("  __shared__ uint4 bfrag[%lld][32];\n", (long long)(KS * NT));
    e("  for (int i = threadIdx.x; i < %lld; i += 128) {\n    const int f = i >> 5, ln = i & 31, ks = f / %lld, nt = f %% %lld;\n    float h[2], l[2];\n", (long long)(KS * NT * 32),
      (long long)NT, (long long)NT);
    e("    #pragma unroll\n    for (int r = 0; r < 2; ++r) cc_split_tf32(ld_b((%s)(nt * 8 + (ln >> 2)), (%s)(ks * 8 + 2 * (ln & 3) + r)%s), h[r], l[r]);\n", IDX, IDX, pass.c_str());
    e("    bfrag[f][ln] = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(l[0]), __float_as_uint(l[1]));\n  }\n  __syncthreads();\n");
  }
  e("  for (%s tile = (%s)blockIdx.x * 4 + (threadIdx.x >> 5); tile < (%s)%lld; tile += (%s)gridDim.x * 4) {\n", IDX, IDX, IDX, (long long)MT, IDX);
  e("    const %s m0 = tile * 16 + gid, m1 = m0 + 8;\n", IDX);
  e("    float c[%lld][4];\n    #pragma unroll\n    for (int nt = 0; nt < %lld; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;\n", (long long)NT, (long long)NT);
  // A is written as "fetch up to 8 k steps, then use them"; ptxas sinks the loads to their uses regardless (two k steps stay in flight).
  // Two tiles per warp at a time (twice the loads in flight, two MMA chains) was timed on the 3 x 3 / depth 8 convolution: 6.74 -> 6.88 us
  // at 126 registers — not kept; what is left above the ~2.7 us launch floor is the legacy MMA rate and the scattered 32-byte pixel reads.
  const int64_t KC = std::min<int64_t>(KS, 8);
  e("    #pragma unroll\n    for (int kc = 0; kc < %lld; kc += %lld) {\n", (long long)KS, (long long)KC);
  if (a_pair) {
    e("      float2 x0[%lld], x1[%lld];\n      #pragma unroll\n      for (int j = 0; j < %lld; ++j) {\n        const %s k0 = (%s)((kc + j) * 8 + 2 * tig);\n", (long long)KC,
      (long long)KC, (long long)KC, IDX, IDX);
    e("        x0[j] = ld_a2(m0, k0%s);\n        x1[j] = ld_a2(m1, k0%s);\n      }\n", pass.c_str(), pass.c_str());
  } else {
    e("      float xa[%lld][4];\n      #pragma unroll\n      for (int j = 0; j < %lld; ++j) {\n        const %s k0 = (%s)((kc + j) * 8 + 2 * tig);\n", (long long)KC, (long long)KC,
      IDX, IDX);
    e("        xa[j][0] = ld_a(m0, k0%s);\n        xa[j][1] = ld_a(m1, k0%s);\n        xa[j][2] = ld_a(m0, k0 + 1%s);\n        xa[j][3] = ld_a(m1, k0 + 1%s);\n      }\n", pass.c_str(),
      pass.c_str(), pass.c_str(), pass.c_str());
  }
  e("      #pragma unroll\n      for (int j = 0; j < %lld; ++j) {\n        const int ks = kc + j;\n        if (ks >= %lld) break;\n", (long long)KC, (long long)KS);
  if (a_pair)
    e("        const float a[4] = {x0[j].x, x1[j].x, x0[j].y, x1[j].y};\n");
  else
    e("        const float a[4] = {xa[j][0], xa[j][1], xa[j][2], xa[j][3]};\n");
  e("      unsigned ah[4], al[4];\n      #pragma unroll\n      for (int i = 0; i < 4; ++i) {\n        float h, l;\n        cc_split_tf32(a[i], h, l);\n"
    "        ah[i] = __float_as_uint(h);\n        al[i] = __float_as_uint(l);\n      }\n");
  if (!b_shared)
    e("      #pragma unroll\n      for (int nt = 0; nt < %lld; ++nt) {\n        cc_mma_tf32_16x8x8(c[nt], al, bh[ks][nt]);\n        cc_mma_tf32_16x8x8(c[nt], ah, bl[ks][nt]);\n"
      "        cc_mma_tf32_16x8x8(c[nt], ah, bh[ks][nt]);\n      }\n      }\n    }\n", (long long)NT);
  else
    e("      #pragma unroll\n      for (int nt = 0; nt < %lld; ++nt) {\n        const uint4 bf = bfrag[ks * %lld + nt][lane];\n        const unsigned bh[2] = {bf.x, bf.y}, bl[2] = {bf.z, bf.w};\n"
      "        cc_mma_tf32_16x8x8(c[nt], al, bh);\n        cc_mma_tf32_16x8x8(c[nt], ah, bl);\n        cc_mma_tf32_16x8x8(c[nt], ah, bh);\n      }\n      }\n    }\n", (long long)NT, (long long)NT);
  e("    #pragma unroll\n    for (int nt = 0; nt < %lld; ++nt) {\n      const %s n0 = (%s)(nt * 8 + 2 * tig);\n      if (n0 >= (%s)%lld) continue;\n", (long long)NT, IDX, IDX, IDX,
    (long long)N);
  e("      #pragma unroll\n      for (int h = 0; h < 2; ++h) {\n        const %s m = h ? m1 : m0;\n        if (m >= (%s)%lld) continue;\n", IDX, IDX, (long long)M);
  if (has_post)
    e("        const float2 v = make_float2(post1(c[nt][2 * h], m, n0%s), post1(c[nt][2 * h + 1], m, n0 + 1%s));\n", pass.c_str(), pass.c_str());
  else
    e("        const float2 v = make_float2(c[nt][2 * h], c[nt][2 * h + 1]);\n");
  e("        __stcs(reinterpret_cast<float2*>(dst + (long long)m * %lld + n0), v);\n      }\n    }\n  }\n}\n", (long long)N);
  plan.source += e.s;
  LaunchSpec ls;
  ls.entry = "small_n_mma";
  ls.grid[0] = (uint32_t)std::min<int64_t>((MT + 3) / 4, 148 * 8);
  ls.block[0] = 128;
  for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
  ls.args.push_back(ARG_OUT);
  plan.launches.push_back(ls);
  plan.kind = PLAN_AXIS_REDUCE;
  plan.flops = 2ull * (uint64_t)M * (uint64_t)N * (uint64_t)K;
  plan.note += strprintf("; small-N contraction %lldx%lldx%lld on warp-level 3xTF32 MMAs, B in %s%s", (long long)M, (long long)N, (long long)K,
                         b_shared ? "shared memory" : "registers", has_post ? ", epilogue in the same kernel" : "");
  return true;
}

// Tries to lower the reduction program to gathered panels + the tcgen05 pipeline. Returns false if the shape of the term does not fit.
bool try_general_contraction(Plan& plan, const Program& p, int n_args) {
  const int nd = (int)p.dims.size();
  const int R = p.n_red, no = nd - R;
  if (R < 1 || no < 2 || p.ops.size() != 3 || p.results.size() != 1) return false;
  const Op& mul = p.ops[p.results[0]];
  if (mul.kind != K_TIMES || mul.a == mul.b || p.ops[mul.a].kind != K_EXTRACT || p.ops[mul.b].kind != K_EXTRACT) return false;
  int la = p.ops[mul.a].load, lb = p.ops[mul.b].load;
  auto uses = [&](const Load& L, int x) {
    for (int y = 0; y < L.rows; ++y)
      if (L.M[(size_t)y * (nd + 1) + x] != 0.0) return true;
    return false;
  };
  // Leading output dims BOTH operands depend on are batch dims: `C[b, i, k] = sum_t A[b, i, t] * B[b, t, k]` is `batch` independent
  // products over panels that carry the batch index in their rows: one tcgen05 pipeline launch per batch (timed in round 2,
  // profiles/r02_knob_ab.json: 4 x 2048 x 1024 x 2048 2.56 -> 0.22 ms, 8 x 512^3 1.4x, small batches unchanged — they stay below the
  // contraction threshold). CC_BATCHED_CONTRACTION=0 keeps such terms on the generic re-rolled reduction.
  int nb = 0;
  const char* batched_env = plan_knob("CC_BATCHED_CONTRACTION");
  if (!batched_env || atoi(batched_env) != 0)
    while (nb < no - 2 && uses(p.loads[la], nb) && uses(p.loads[lb], nb)) ++nb;
  auto split_point = [&](const Load& A, const Load& B) {
    // dims [nb, s) not used by B, dims [s, no) not used by A
    int s = nb;
    while (s < no && !uses(B, s)) ++s;
    for (int x = s; x < no; ++x)
      if (uses(A, x)) return -1;
    return (s >= nb + 1 && s < no) ? s : -1;
  };
  int s = split_point(p.loads[la], p.loads[lb]);
  if (s < 0) {
    std::swap(la, lb);
    s = split_point(p.loads[la], p.loads[lb]);
    if (s < 0) return false;
  }
  int64_t BATCH = 1, M = 1, N = 1, K = 1;
  for (int x = 0; x < nb; ++x) BATCH *= p.dims[x];
  for (int x = nb; x < s; ++x) M *= p.dims[x];
  for (int x = s; x < no; ++x) N *= p.dims[x];
  for (int x = no; x < nd; ++x) K *= p.dims[x];
  if (nb == 0 && try_small_n_mma(plan, p, n_args, la, lb, s, M, N, K)) return true;
  int64_t min_macs = (int64_t)1 << 25;
  if (const char* ev = plan_knob("CC_TUNE_CONTRACTION_MIN_MACS")) min_macs = atoll(ev);
  // the gathered panels cost 8 bytes of HBM traffic per (row, k) each way, so this pays off when N (the reuse of an A row) is large
  if (M * N * K < min_macs || N < 32 || K < 32 || M >= ((int64_t)1 << 31) || N >= ((int64_t)1 << 31) || K >= ((int64_t)1 << 31) - 32) return false;
  const int64_t Kp = (K + 31) / 32 * 32;
  if (BATCH * M * Kp >= ((int64_t)1 << 40) || BATCH * N * Kp >= ((int64_t)1 << 40) || BATCH > 65535) return false;
  std::vector<int> m_dims, n_dims;  // the rows of a panel: batch dims first, so batch b's rows are one contiguous block
  for (int x = 0; x < s; ++x) m_dims.push_back(x);
  for (int x = 0; x < nb; ++x) n_dims.push_back(x);
  for (int x = s; x < no; ++x) n_dims.push_back(x);
  Emit e;
  emit_panel_kernel(e, p, "panel_a", la, m_dims, n_args);
  emit_panel_kernel(e, p, "panel_b", lb, n_dims, n_args);
  const bool has_post = !p.post_ops.empty() && !p.trivial_post();
  if (has_post) emit_post_kernel(e, p, n_args);
  plan.source += e.s;
  auto add = [&](const char* entry, int64_t threads, std::vector<int> extra) {
    LaunchSpec ls;
    ls.entry = entry;
    ls.grid[0] = (uint32_t)std::max<int64_t>(1, (threads + 255) / 256);
    ls.block[0] = 256;
    for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
    for (int a : extra) ls.args.push_back(a);
    plan.launches.push_back(ls);
  };
  if (BATCH * M * (Kp / 4) + 255 >= ((int64_t)1 << 31) * 256 || BATCH * N * (Kp / 4) + 255 >= ((int64_t)1 << 31) * 256) return false;
  add("panel_a", BATCH * M * (Kp / 4), {ARG_SCRATCH0, ARG_SCRATCH0 - 1});
  add("panel_b", BATCH * N * (Kp / 4), {ARG_SCRATCH0 - 2, ARG_SCRATCH0 - 3});
  if (has_post) {
    const int V = (p.dims[no - 1] % 4 == 0) ? 4 : 1;
    add("post_kernel", BATCH * M * N / V, {ARG_OUT});
  }
  plan.kind = PLAN_CONTRACTION;
  plan.M = M;
  plan.N = N;
  plan.K = K;
  plan.gathered_panels = true;
  plan.batch = BATCH;
  plan.flops = 2ull * (uint64_t)BATCH * (uint64_t)M * (uint64_t)N * (uint64_t)K;
  plan.scratch_floats = {(uint64_t)(BATCH * M * Kp), (uint64_t)(BATCH * M * Kp), (uint64_t)(BATCH * N * Kp), (uint64_t)(BATCH * N * Kp)};
  plan.note += strprintf("; general contraction %lldx%lldx%lld over gathered operand panels -> tcgen05 3xTF32%s", (long long)M, (long long)N,
                         (long long)K, has_post ? " + in-place epilogue" : "");
  if (BATCH > 1) plan.note += strprintf(" (batch of %lld)", (long long)BATCH);
  return true;
}

// ---- whole-tensor fold (Tensor.sum and the other monoids) with the operand's closure fused in -----------------------------
//
// Same schedule as the precompiled reduce_sum_kernel (kernels_basic.cu): 512 threads, grid <= 4 CTAs per SM, U independent
// 128-bit vectors in flight per thread folded into U accumulators, warp-shuffle + shared-memory block fold, deterministic
// last-block-done second stage. With U == 4 and a 4-divisible element count the fold order is identical to reduce_sum_kernel,
// so `expr.sum` fused and `expr.doCache.sum` give the same bits.
const char* monoid_type(uint32_t m) {
  switch (m) {
    case K_PLUS: return "cc_plus";
    case K_TIMES: return "cc_times";
    case K_MIN: return "cc_min";
    case K_MAX: return "cc_max";
  }
  return "?";
}

void emit_full_reduce(Plan& plan, const Program& p, int n_args, const DeviceProps& dev, uint32_t monoid) {
  const int nd = (int)p.dims.size();
  const int64_t total = product(p.dims);
  const bool flat = program_is_flat(p);
  const int V = ((nd >= 1 && p.dims[nd - 1] % 4 == 0) || (flat && total >= 4)) ? 4 : 1;
  const int64_t NV = total / V;
  const int64_t tail = total - NV * V;  // flat programs only: <= 3 leftover elements, folded by one designated thread
  const char* IDX = pick_idx_type(p, total + (int64_t)4 * kFullReduceMaxBlocks * kFullReduceThreads);
  const int nloads = (int)p.loads.size();
  const int U = nloads * V <= 16 ? 4 : (nloads * V <= 32 ? 2 : 1);
  int64_t want = (NV + (int64_t)kFullReduceThreads * 4 - 1) / ((int64_t)kFullReduceThreads * 4);
  int64_t cap = std::min<int64_t>((int64_t)dev.sm_count * 4, kFullReduceMaxBlocks);
  const int64_t grid = std::max<int64_t>(1, std::min(want, cap));
  const char* M = monoid_type(monoid);

  Emit e;
  e("// whole-tensor fold: dims=[");
  for (int x = 0; x < nd; ++x) e("%s%lld", x ? "," : "", (long long)p.dims[x]);
  e("] monoid=%s V=%d U=%d flat=%d idx=%s loads=%d ops=%zu grid=%lld\n", M, V, U, (int)flat, IDX, nloads, p.ops.size(), (long long)grid);
  e("__device__ __forceinline__ void ev(const %s v%s%s, float (&o)[%d]) {\n", IDX, n_args ? ", " : "", param_list(n_args, false).c_str(), V);
  if (flat) {
    for (int j = 0; j < nloads; ++j) {
      e("  float L%d[%d];\n", j, V);
      if (V == 4)
        e("  cc_ldg4(p%d + v * 4, L%d);\n", p.loads[j].arg, j);
      else
        e("  L%d[0] = cc_ldg(p%d + v);\n", j, p.loads[j].arg);
    }
  } else {
    LoadCtx c{V, nd - 1, IDX};
    emit_decode(e, p.dims, nd, IDX, strprintf("v * %d", V).c_str(), "  ");
    for (int j = 0; j < nloads; ++j) emit_load(e, p, j, c, "  ");
  }
  e("  #pragma unroll\n  for (int l = 0; l < %d; ++l) {\n", V);
  emit_ops(e, p, "    ", "l");
  e("    o[l] = _%d;\n  }\n}\n", p.results[0]);
  e("extern \"C\" __global__ void __launch_bounds__(%d) reduce_all(%s%sfloat* __restrict__ out, float* __restrict__ partials, unsigned* __restrict__ counter) {\n",
    kFullReduceThreads, param_list(n_args, false).c_str(), n_args ? ", " : "");
  e("  typedef %s M;\n  __shared__ float smem[32];\n", M);
  e("  const %s NV = %lld;\n  const %s stride = (%s)gridDim.x * %d;\n  %s v = (%s)blockIdx.x * %d + threadIdx.x;\n", IDX, (long long)NV, IDX, IDX,
    kFullReduceThreads, IDX, IDX, kFullReduceThreads);
  e("  float a[%d][%d];\n  #pragma unroll\n  for (int u = 0; u < %d; ++u)\n    #pragma unroll\n    for (int l = 0; l < %d; ++l) a[u][l] = M::zero();\n", U, V, U, V);
  const std::string pass = (n_args ? ", " : "") + arg_pass(n_args);
  if (U > 1) {
    e("  for (; v + %d * stride < NV; v += %d * stride) {\n    float x[%d][%d];\n", U - 1, U, U, V);
    e("    #pragma unroll\n    for (int u = 0; u < %d; ++u) ev(v + u * stride%s, x[u]);\n", U, pass.c_str());
    e("    #pragma unroll\n    for (int u = 0; u < %d; ++u)\n      #pragma unroll\n      for (int l = 0; l < %d; ++l) a[u][l] = M::ap(a[u][l], x[u][l]);\n  }\n", U, V);
  }
  e("  for (; v < NV; v += stride) {\n    float x[%d];\n    ev(v%s, x);\n", V, pass.c_str());
  e("    #pragma unroll\n    for (int l = 0; l < %d; ++l) a[0][l] = M::ap(a[0][l], x[l]);\n  }\n", V);
  // fold the U accumulators per lane, then the lanes: ((x + y) + (z + w))
  e("  float q[%d];\n  #pragma unroll\n  for (int l = 0; l < %d; ++l) ", V, V);
  if (U == 4)
    e("q[l] = M::ap(M::ap(a[0][l], a[1][l]), M::ap(a[2][l], a[3][l]));\n");
  else if (U == 2)
    e("q[l] = M::ap(a[0][l], a[1][l]);\n");
  else
    e("q[l] = a[0][l];\n");
  if (V == 4)
    e("  float acc = M::ap(M::ap(q[0], q[1]), M::ap(q[2], q[3]));\n");
  else
    e("  float acc = q[0];\n");
  if (tail > 0) {
    // same place as reduce_sum_kernel's scalar tail: thread 0 of block 0, after its own vectors
    e("  if (blockIdx.x == 0 && threadIdx.x == 0) {\n    for (%s i = (%s)%lld; i < (%s)%lld; ++i) {\n", IDX, IDX, (long long)(NV * V), IDX, (long long)total);
    for (int j = 0; j < nloads; ++j) e("      const float L%d[1] = {cc_ldg(p%d + i)};\n", j, p.loads[j].arg);
    emit_ops(e, p, "      ", "0");
    e("      acc = M::ap(acc, _%d);\n    }\n  }\n", p.results[0]);
  }
  e("  acc = cc_block_fold<M>(acc, smem);\n  cc_fold_finish<M>(acc, out, partials, counter, smem);\n}\n");
  plan.source += e.s;
  LaunchSpec ls;
  ls.entry = "reduce_all";
  ls.grid[0] = (uint32_t)grid;
  ls.block[0] = kFullReduceThreads;
  for (int i = 0; i < n_args; ++i) ls.args.push_back(i);
  ls.args.push_back(ARG_OUT);
  ls.args.push_back(ARG_REDUCE_PARTIALS);
  ls.args.push_back(ARG_REDUCE_COUNTER);
  plan.launches.push_back(ls);
}

// ---- definition inlining -----------------------------------------------------------------------------------------

// M_outer: rows_p x (nd+1) maps the reduction's index space onto P's index space; M_def: rows_s x (rows_p+1) maps P's
// index space onto a source. Returns rows_s x (nd+1). (Same accumulation order as NDimensionalAffineTransform.scala:48-95.)
std::vector<double> compose(const std::vector<double>& M_def, int rows_s, const std::vector<double>& M_outer, int rows_p, int nd) {
  std::vector<double> out((size_t)rows_s * (nd + 1), 0.0);
  for (int y = 0; y < rows_s; ++y) {
    for (int x = 0; x < nd; ++x) {
      double acc = 0.0;
      for (int k = 0; k < rows_p; ++k) acc = acc + M_def[(size_t)y * (rows_p + 1) + k] * M_outer[(size_t)k * (nd + 1) + x];
      out[(size_t)y * (nd + 1) + x] = acc;
    }
    double acc = M_def[(size_t)y * (rows_p + 1) + rows_p];
    for (int k = 0; k < rows_p; ++k) acc = acc + M_def[(size_t)y * (rows_p + 1) + k] * M_outer[(size_t)k * (nd + 1) + nd];
    out[(size_t)y * (nd + 1) + nd] = acc;
  }
  return out;
}

}  // namespace

Plan make_plan(const Tree& t, const DeviceProps& dev) {
  Plan plan;
  const Node& root = t.nodes[t.root];
  std::vector<int64_t> odims(t.out_shape.begin(), t.out_shape.end());
  std::map<uint32_t, int> arg_of_param;
  std::vector<uint32_t> arg_nodes;
  std::vector<int32_t> ordinal_of_node(t.nodes.size(), -1);
  for (size_t p = 0; p < t.params.size(); ++p) ordinal_of_node[t.params[p]] = (int32_t)p;

  Program prog;
  bool is_reduce = false;
  int nd_base = (int)odims.size();

  if (root.kind == K_REDUCE) {
    CC_REQUIRE(product(odims) == 1, CC_ERR_BAD_TREE, "Reduce produces one float; output shape has %lld elements", (long long)product(odims));
    nd_base = (int)root.shape.size();
    Builder b{t, {}, nd_base, {}, {}, &arg_of_param, &arg_nodes};
    for (int32_t s : root.shape) b.prog.dims.push_back(s);
    b.prog.results.push_back(b.export_node(root.kids[0]));
    prog = std::move(b.prog);
    plan.note = strprintf("whole-tensor %s fold with the operand's closure fused in", kind_name(root.monoid));
  } else {
    // Tensor.join (Tensors.scala:577-598): a Concatenate root; its elements are evaluated over the head shape = out_shape minus
    // the last dim. Congruent elements are re-rolled into that last output dimension (index c).
    const bool joined = root.kind == K_CONCAT;
    StepMap step_c;
    bool rolled_c = false;
    uint32_t base = t.root;
    // where the element index lands in the output: last (Tensor.join(tensors), T:577-598) or `position` (join(tensors, dimension),
    // T:560-575: the reference joins last and gathers a permuted view in a second kernel)
    int join_pos = -1;
    std::vector<int64_t> final_odims = odims;
    if (joined) {
      CC_REQUIRE(!odims.empty(), CC_ERR_BAD_TREE, "Concatenate needs an output shape of rank >= 1");
      join_pos = root.position < 0 ? (int)odims.size() - 1 : root.position;
      CC_REQUIRE(join_pos < (int)odims.size() && odims[(size_t)join_pos] == (int64_t)root.kids.size(), CC_ERR_BAD_TREE,
                 "Concatenate of %zu elements at dimension %d does not match the output shape", root.kids.size(), join_pos);
      nd_base = (int)odims.size() - 1;
      // build over head dims + [element index]; moved into place afterwards
      odims.erase(odims.begin() + join_pos);
      odims.push_back((int64_t)root.kids.size());
      rolled_c = root.kids.size() >= 2 && reroll(t, std::vector<uint32_t>(root.kids.begin(), root.kids.end()), step_c);
      base = root.kids[0];
    }
    if (joined && !rolled_c) {
      Builder b{t, {}, nd_base, {}, {}, &arg_of_param, &arg_nodes};
      b.prog.dims.assign(odims.begin(), odims.end() - 1);
      for (uint32_t k : root.kids) b.prog.results.push_back(b.export_node(k));
      for (int x = join_pos; x < nd_base; ++x) b.prog.tuple_inner *= odims[(size_t)x];
      prog = std::move(b.prog);
      plan.note = "join as per-index tuple stores";
    } else {
      // A long left-leaning Plus chain of congruent terms inside the expression (`t.split(axis).reduce(_ + _)`, both matmul
      // formulations, the convolution of benchmarks.scala:526-545) is re-rolled into a reduction over 1..n nested indices; what
      // surrounds the chain (e.g. `bias + chain`) becomes the epilogue applied once per output element.
      Chain ch;
      const bool found = find_chain(t, base, ch) && (!rolled_c || join_step_uniform_over_terms(t, ch, step_c));
      std::vector<const StepMap*> exts;
      if (rolled_c) exts.push_back(&step_c);
      if (found)
        for (const StepMap& sm : ch.steps) exts.push_back(&sm);
      Builder b{t, {}, nd_base, exts, {}, &arg_of_param, &arg_nodes};
      b.prog.dims = odims;  // (head dims + the re-rolled element index c as the fastest output dimension when joined)
      if (found) {
        for (int64_t n : ch.levels) b.prog.dims.push_back(n);
        b.prog.n_red = (int)ch.levels.size();
        b.prog.red_monoid = ch.monoid;
        b.prog.results.push_back(b.export_node(ch.terms[0]));
        b.begin_post(ch.top);
        b.prog.post_result = b.export_node(base);
        is_reduce = true;
        std::string lv;
        for (size_t j = 0; j < ch.levels.size(); ++j) lv += strprintf("%s%lld", j ? " x " : "", (long long)ch.levels[j]);
        plan.note = strprintf("%s%s chain of %zu congruent terms re-rolled into a reduction over %s%s", rolled_c ? "join re-rolled into an output dimension; " : "",
                              kind_name(ch.monoid), ch.terms.size(), lv.c_str(), b.prog.trivial_post() ? "" : " with an elementwise epilogue");
      } else {
        b.prog.results.push_back(b.export_node(base));
        if (rolled_c) plan.note = "join re-rolled into an output dimension";
      }
      prog = std::move(b.prog);
      if (joined && join_pos != nd_base) move_index_dim(prog, nd_base, join_pos);
    }
    odims = final_odims;
  }

  if (is_reduce && prog.trivial_post() && prog.n_red == 1) {
    // compose the closure of an unevaluated inline operand into the reduction (looks through the fusion barrier,
    // SURVEY finding 2) when the view of it is integer and provably in range
    const int nd = (int)prog.dims.size();
    bool changed = false;
    Program np;
    np.dims = prog.dims;
    std::map<uint32_t, int> arg2;
    std::vector<uint32_t> nodes2;
    std::vector<int> remap(prog.ops.size(), -1);
    for (size_t i = 0; i < prog.ops.size(); ++i) {
      const Op& op = prog.ops[i];
      if (op.kind == K_EXTRACT) {
        const Load& L = prog.loads[op.load];
        uint32_t pnode = arg_nodes[L.arg];
        const Node& pn = t.nodes[pnode];
        if (pn.def_root >= 0 && L.integer && !L.any_check()) {
          // export the definition over P's own index space, then re-map every load through L.M
          std::map<uint32_t, int> a3;
          std::vector<uint32_t> n3;
          Builder db{t, {}, (int)pn.shape.size(), {}, {}, &a3, &n3};
          for (int32_t s : pn.shape) db.prog.dims.push_back(s);
          int res = db.export_node((uint32_t)pn.def_root);
          std::vector<int> dmap(db.prog.ops.size(), -1);
          for (size_t k = 0; k < db.prog.ops.size(); ++k) {
            Op o = db.prog.ops[k];
            if (o.kind == K_EXTRACT) {
              Load dl = db.prog.loads[o.load];
              uint32_t srcnode = n3[dl.arg];
              auto it = arg2.find(srcnode);
              if (it == arg2.end()) {
                arg2[srcnode] = (int)nodes2.size();
                nodes2.push_back(srcnode);
                it = arg2.find(srcnode);
              }
              Load nl;
              nl.arg = it->second;
              nl.src_shape = dl.src_shape;
              nl.padding = dl.padding;
              nl.M = compose(dl.M, (int)dl.src_shape.size(), L.M, (int)pn.shape.size(), nd);
              analyze_load(nl, np.dims);
              np.loads.push_back(std::move(nl));
              o.load = (int)np.loads.size() - 1;
            } else {
              if (o.a >= 0) o.a = dmap[o.a];
              if (o.b >= 0) o.b = dmap[o.b];
            }
            np.ops.push_back(o);
            dmap[k] = (int)np.ops.size() - 1;
          }
          remap[i] = dmap[res];
          changed = true;
          continue;
        }
        Load nl = L;
        auto it = arg2.find(pnode);
        if (it == arg2.end()) {
          arg2[pnode] = (int)nodes2.size();
          nodes2.push_back(pnode);
          it = arg2.find(pnode);
        }
        nl.arg = it->second;
        np.loads.push_back(std::move(nl));
        Op o = op;
        o.load = (int)np.loads.size() - 1;
        np.ops.push_back(o);
        remap[i] = (int)np.ops.size() - 1;
      } else {
        Op o = op;
        if (o.a >= 0) o.a = remap[o.a];
        if (o.b >= 0) o.b = remap[o.b];
        np.ops.push_back(o);
        remap[i] = (int)np.ops.size() - 1;
      }
    }
    if (changed) {
      np.results.push_back(remap[prog.results[0]]);
      np.n_red = prog.n_red;
      np.red_monoid = prog.red_monoid;
      np.post_ops = prog.post_ops;
      np.post_result = prog.post_result;
      np.load_in_post.assign(np.loads.size(), 0);
      prog = std::move(np);
      arg_nodes = nodes2;
      plan.note += "; inline operand composed into the reduction (never materialised)";
    }
  }

  const int n_args = (int)arg_nodes.size();
  for (uint32_t pn : arg_nodes) {
    CC_REQUIRE(ordinal_of_node[pn] >= 0, CC_ERR_BAD_TREE, "parameter node %u unreachable", pn);
    plan.arg_params.push_back((uint32_t)ordinal_of_node[pn]);
    int64_t n = 1;
    for (int32_t s : t.nodes[pn].shape) n *= s;
    plan.arg_min_floats.push_back((uint64_t)n);
  }
  plan.out_floats = (uint64_t)product(odims);
  const int64_t space = product(prog.dims);
  // algorithmic traffic: every distinct source once (or as much of it as the index space can touch) + the output once
  {
    std::vector<uint64_t> touched(n_args, 0);
    for (const Load& L : prog.loads) touched[L.arg] += (uint64_t)space;
    uint64_t bytes = plan.out_floats * 4;
    for (int a = 0; a < n_args; ++a) bytes += 4 * std::min<uint64_t>(touched[a], plan.arg_min_floats[a]);
    plan.algorithmic_bytes = bytes;
    plan.flops = count_flops(prog) * (uint64_t)space + ((is_reduce || root.kind == K_REDUCE) ? (uint64_t)space : 0);
  }

  // A source read through several views in one kernel (stencils: x + x.translate(...)) would be fetched from L2 once per view
  // with streaming loads (5-point stencil on 512^3: 0.27 ms, 0.62 of HBM); let those lines live in L1.
  {
    std::vector<int> uses((size_t)n_args, 0);
    for (const Load& L : prog.loads) ++uses[(size_t)L.arg];
    for (Load& L : prog.loads)
      if (L.integer && uses[(size_t)L.arg] > 1) L.reuse = true;
  }

  if (!prog.dims.empty())
    for (const Load& L : prog.loads)
      if (L.integer && L.coef[prog.dims.size() - 1] == 1 && L.base % 4 != 0) ++prog.shifted_loads;

  if (root.kind == K_REDUCE) {
    plan.kind = PLAN_FULL_REDUCE;
    emit_full_reduce(plan, prog, n_args, dev, root.monoid);
    return plan;
  }

  if (is_reduce) {
    plan.kind = PLAN_AXIS_REDUCE;
    // contraction: sum_t A[i,t] * B[t,k] with A [M,K] and B [K,N] row-major
    const int nd = (int)prog.dims.size();
    if (dev.contraction && prog.red_monoid == K_PLUS && nd == 3 && prog.n_red == 1 && prog.trivial_post() && prog.ops.size() == 3 && prog.loads.size() == 2 &&
        prog.ops[prog.results[0]].kind == K_TIMES) {
      const Op& mul = prog.ops[prog.results[0]];
      if (prog.ops[mul.a].kind == K_EXTRACT && prog.ops[mul.b].kind == K_EXTRACT && mul.a != mul.b) {
        const int64_t M = prog.dims[0], N = prog.dims[1], K = prog.dims[2];
        auto is_A = [&](const Load& L) {
          return L.integer && !L.any_check() && L.base == 0 && L.coef[0] == K && L.coef[1] == 0 && L.coef[2] == 1 &&
                 product(L.src_shape) == M * K;
        };
        auto is_B = [&](const Load& L) {
          return L.integer && !L.any_check() && L.base == 0 && L.coef[0] == 0 && L.coef[1] == 1 && L.coef[2] == N &&
                 product(L.src_shape) == K * N;
        };
        const Load& La = prog.loads[prog.ops[mul.a].load];
        const Load& Lb = prog.loads[prog.ops[mul.b].load];
        int a_arg = -1, b_arg = -1;
        if (is_A(La) && is_B(Lb)) a_arg = La.arg, b_arg = Lb.arg;
        if (is_A(Lb) && is_B(La)) a_arg = Lb.arg, b_arg = La.arg;
        // Any shape runs on the tensor cores (ragged edges are handled by the pipeline); tiny or very skinny products stay on
        // the generic reduction, which needs one launch and no hi/lo workspace traffic. Measured on B200
        // (scripts/gpu_gemm_shapes.py, profiles/r01_gemm_shapes.json): the three-launch pipeline costs ~15 us at least, the
        // generic kernel ~8-13 us; they cross near 2^24..2^25 multiply-adds (256^3: 15.5 vs 13.1 us, 8192x64x64: 15.3 vs
        // 16.7 us, 65536x32x32: 16.9 vs 29.1 us, 512^3: 20.9 vs 33.4 us).
        int64_t min_macs = (int64_t)1 << 25;
        if (const char* ev = plan_knob("CC_TUNE_CONTRACTION_MIN_MACS")) min_macs = atoll(ev);
        // (few columns and a short K: B fits in a warp's registers -- one generated kernel on warp-level MMAs, no workspace: 65536x32x32
        // 16.4 -> see profiles/r02_small_n_mma.md)
        const bool worth = M * N * K >= min_macs && N >= 32 && K >= 32 && !small_n_mma_fits(M, N, K);
        if (a_arg >= 0 && a_arg != b_arg && worth && M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31) && K < ((int64_t)1 << 31) - 32) {
          const int64_t Kp = (K + 31) / 32 * 32;  // gemm_padded_k
          plan.kind = PLAN_CONTRACTION;
          plan.M = M;
          plan.N = N;
          plan.K = K;
          // canonical argument order: A then B
          std::vector<uint32_t> ap{plan.arg_params[a_arg], plan.arg_params[b_arg]};
          std::vector<uint64_t> am{plan.arg_min_floats[a_arg], plan.arg_min_floats[b_arg]};
          plan.arg_params = ap;
          plan.arg_min_floats = am;
          plan.flops = 2ull * (uint64_t)M * (uint64_t)N * (uint64_t)K;
          plan.algorithmic_bytes = 4ull * (uint64_t)(M * K + K * N + M * N);
          plan.scratch_floats = {(uint64_t)(M * Kp), (uint64_t)(M * Kp), (uint64_t)(N * Kp), (uint64_t)(N * Kp)};
          plan.note += strprintf("; contraction %lldx%lldx%lld -> tcgen05 3xTF32", (long long)M, (long long)N, (long long)K);
          plan.source = "// contraction pattern: runs the precompiled TMA + tcgen05 3xTF32 pipeline (gemm_3xtf32.cu)\n";
          return plan;
        }
      }
    }
    if (dev.contraction && prog.red_monoid == K_PLUS && try_general_contraction(plan, prog, n_args)) return plan;
    emit_reduce(plan, prog, n_args, dev);
  } else {
    const int td = transpose_dim(prog);
    if (try_emit_stencil_tile(plan, prog, n_args, dev)) {
      plan.kind = PLAN_ELEMENTWISE;
    } else if (td >= 0) {
      plan.kind = PLAN_TILED_TRANSPOSE;
      emit_tiled_transpose(plan, prog, n_args, dev, td);
    } else {
      plan.kind = PLAN_ELEMENTWISE;
      emit_elementwise(plan, prog, n_args, dev, 0);
    }
  }
  return plan;
}

}  // namespace cc
