// ir.cpp — tree blob parsing, structural key, blob writer.
#include "ir.h"

#include <cstring>

namespace cc {

static thread_local std::string g_last_error;

std::string strprintf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  va_list ap2;
  va_copy(ap2, ap);
  int n = vsnprintf(nullptr, 0, fmt, ap);
  va_end(ap);
  std::string s((size_t)(n > 0 ? n : 0), '\0');
  if (n > 0) vsnprintf(&s[0], (size_t)n + 1, fmt, ap2);
  va_end(ap2);
  return s;
}
void set_last_error(const std::string& m) { g_last_error = m; }
const char* last_error_cstr() { return g_last_error.c_str(); }

const char* kind_name(uint32_t k) {
  switch (k) {
    case K_LITERAL: return "FloatLiteral";
    case K_PARAM: return "ArrayParameter";
    case K_TRANSFORM: return "Transform";
    case K_EXTRACT: return "Extract";
    case K_CONCAT: return "Concatenate";
    case K_EXP: return "Exp";
    case K_LOG: return "Log";
    case K_ABS: return "Abs";
    case K_TANH: return "Tanh";
    case K_SQRT: return "Sqrt";
    case K_NEG: return "UnaryMinus";
    case K_MIN: return "Min";
    case K_MAX: return "Max";
    case K_PLUS: return "Plus";
    case K_MINUS: return "Minus";
    case K_TIMES: return "Times";
    case K_DIV: return "Div";
    case K_PERCENT: return "Percent";
    case K_REDUCE: return "Reduce";
  }
  return "?";
}

namespace {
constexpr uint32_t kMagic = 0x31544343u;  // 'CCT1'

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  void need(size_t n) {
    if ((size_t)(end - p) < n) fail(CC_ERR_BAD_TREE, "tree blob truncated");
  }
  uint32_t u32() {
    need(4);
    uint32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  int32_t i32() { return (int32_t)u32(); }
  float f32() {
    need(4);
    float v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint64_t u64() {
    need(8);
    uint64_t v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  double f64() {
    need(8);
    double v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
};
}  // namespace

Tree parse_tree(const void* blob, uint64_t n_bytes) {
  CC_REQUIRE(blob && n_bytes >= 16, CC_ERR_BAD_TREE, "tree blob too small");
  Reader r{(const uint8_t*)blob, (const uint8_t*)blob + n_bytes};
  CC_REQUIRE(r.u32() == kMagic, CC_ERR_BAD_TREE, "bad tree blob magic");
  uint32_t n = r.u32();
  Tree t;
  t.root = r.u32();
  uint32_t out_rank = r.u32();
  CC_REQUIRE(n >= 1 && t.root < n, CC_ERR_BAD_TREE, "bad node count / root");
  CC_REQUIRE(out_rank <= 16, CC_ERR_BAD_TREE, "output rank %u too large", out_rank);
  for (uint32_t i = 0; i < out_rank; ++i) {
    int32_t s = r.i32();
    CC_REQUIRE(s >= 0, CC_ERR_BAD_TREE, "negative output dimension");
    t.out_shape.push_back(s);
  }
  t.nodes.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    Node& nd = t.nodes[i];
    nd.kind = r.u32();
    auto kid = [&]() {
      uint32_t k = r.u32();
      CC_REQUIRE(k < i, CC_ERR_BAD_TREE, "node %u references node %u (children must precede parents)", i, k);
      return k;
    };
    if (nd.kind == K_LITERAL) {
      nd.value = r.f32();
    } else if (nd.kind == K_PARAM) {
      nd.param_id = r.u64();
      nd.value = r.f32();
      uint32_t rank = r.u32();
      CC_REQUIRE(rank <= 16, CC_ERR_BAD_TREE, "parameter rank %u too large", rank);
      for (uint32_t d = 0; d < rank; ++d) {
        int32_t s = r.i32();
        CC_REQUIRE(s >= 0, CC_ERR_BAD_TREE, "negative parameter dimension");
        nd.shape.push_back(s);
      }
      nd.def_root = r.i32();
      CC_REQUIRE(nd.def_root < (int32_t)n, CC_ERR_BAD_TREE, "definition root out of range");
    } else if (nd.kind == K_TRANSFORM) {
      nd.kids.push_back(kid());
      nd.rows = r.u32();
      nd.cols = r.u32();
      CC_REQUIRE(nd.rows <= 16 && nd.cols >= 1 && nd.cols <= 17, CC_ERR_BAD_TREE, "bad matrix size");
      nd.matrix.resize((size_t)nd.rows * nd.cols);
      for (double& v : nd.matrix) v = r.f64();
      CC_REQUIRE(t.nodes[nd.kids[0]].kind == K_PARAM, CC_ERR_BAD_TREE,
                 "Transform must wrap an ArrayParameter (views are pre-composed on the host, Tensors.scala:979-989)");
      CC_REQUIRE(t.nodes[nd.kids[0]].shape.size() == nd.rows, CC_ERR_BAD_TREE, "matrix rows != source rank");
    } else if (nd.kind == K_EXTRACT) {
      nd.kids.push_back(kid());
      uint32_t ak = t.nodes[nd.kids[0]].kind;
      CC_REQUIRE(ak == K_PARAM || ak == K_TRANSFORM, CC_ERR_BAD_TREE, "Extract of a non-array node");
    } else if (nd.kind == K_CONCAT || nd.kind == K_CONCAT_AT) {
      uint32_t m = r.u32();
      CC_REQUIRE(m >= 1 && m <= (1u << 24), CC_ERR_BAD_TREE, "bad Concatenate length");
      if (nd.kind == K_CONCAT_AT) {
        uint32_t pos = r.u32();
        CC_REQUIRE(pos < 16, CC_ERR_BAD_TREE, "bad Concatenate position %u", pos);
        nd.position = (int32_t)pos;
        nd.kind = K_CONCAT;
      }
      for (uint32_t j = 0; j < m; ++j) nd.kids.push_back(kid());
    } else if (nd.kind == K_REDUCE) {
      nd.monoid = r.u32();
      CC_REQUIRE(nd.monoid == K_PLUS || nd.monoid == K_MIN || nd.monoid == K_MAX || nd.monoid == K_TIMES, CC_ERR_BAD_TREE,
                 "Reduce monoid must be Plus, Min, Max or Times (got %u)", nd.monoid);
      nd.kids.push_back(kid());
      uint32_t rank = r.u32();
      CC_REQUIRE(rank <= 16, CC_ERR_BAD_TREE, "Reduce rank %u too large", rank);
      for (uint32_t d = 0; d < rank; ++d) {
        int32_t s = r.i32();
        CC_REQUIRE(s >= 0, CC_ERR_BAD_TREE, "negative Reduce dimension");
        nd.shape.push_back(s);
      }
      CC_REQUIRE(i == t.root, CC_ERR_BAD_TREE, "Reduce is only allowed at the root of a tree");
    } else if (is_unary(nd.kind)) {
      nd.kids.push_back(kid());
    } else if (is_binary(nd.kind)) {
      nd.kids.push_back(kid());
      nd.kids.push_back(kid());
    } else {
      fail(CC_ERR_BAD_TREE, strprintf("unknown node kind %u", nd.kind));
    }
    if (nd.kind == K_CONCAT || nd.kind == K_REDUCE || is_unary(nd.kind) || is_binary(nd.kind)) {
      for (uint32_t k : nd.kids) {
        uint32_t kk = t.nodes[k].kind;
        CC_REQUIRE(kk != K_PARAM && kk != K_TRANSFORM && kk != K_CONCAT && kk != K_REDUCE, CC_ERR_BAD_TREE,
                   "%s operand must be a float term", kind_name(nd.kind));
      }
    }
  }
  CC_REQUIRE(r.p == r.end, CC_ERR_BAD_TREE, "trailing bytes after tree blob");
  uint32_t rk = t.nodes[t.root].kind;
  CC_REQUIRE(rk != K_PARAM && rk != K_TRANSFORM, CC_ERR_BAD_TREE, "root must be a value term");
  for (const Node& nd : t.nodes)
    if (nd.kind == K_PARAM && nd.def_root >= 0) {
      uint32_t dk = t.nodes[nd.def_root].kind;
      CC_REQUIRE(dk != K_PARAM && dk != K_TRANSFORM && dk != K_CONCAT && dk != K_REDUCE, CC_ERR_BAD_TREE, "definition must be a float term");
    }
  return t;
}

void canonicalize(Tree& t) {
  const uint32_t n = (uint32_t)t.nodes.size();
  // per-call index tables: inline for the small trees of throw-away expressions (no heap traffic), on the heap for long chains
  SmallVec<int32_t, 64> canon, param_ordinal;
  SmallVec<uint32_t, 64> order, stack;
  for (uint32_t i = 0; i < n; ++i) canon.push_back(-1), param_ordinal.push_back(-1);
  t.params.clear();
  t.params.reserve(8);
  auto dfs = [&](uint32_t root) {
    stack.push_back(root);
    while (!stack.empty()) {
      uint32_t i = stack.back();
      stack.pop_back();
      if (canon[i] >= 0) continue;
      canon[i] = (int32_t)order.size();
      order.push_back(i);
      const Node& nd = t.nodes[i];
      if (nd.kind == K_PARAM) {
        param_ordinal[i] = (int32_t)t.params.size();
        t.params.push_back(i);
      }
      for (size_t k = nd.kids.size(); k-- > 0;) stack.push_back(nd.kids[k]);
    }
  };
  dfs(t.root);
  t.n_main_params = (uint32_t)t.params.size();
  // definitions, in parameter order (parameters found inside definitions are appended and processed in turn)
  std::vector<std::pair<uint32_t, uint32_t>> defs;  // (param ordinal, def root)
  for (size_t p = 0; p < t.params.size(); ++p) {
    const Node& nd = t.nodes[t.params[p]];
    if (nd.def_root >= 0) {
      dfs((uint32_t)nd.def_root);
      defs.emplace_back((uint32_t)p, (uint32_t)nd.def_root);
    }
  }
  std::string& key = t.key;
  key.clear();
  key.reserve(order.size() * 12 + 64);
  auto put32 = [&](uint32_t v) { key.append((const char*)&v, 4); };
  auto putf = [&](float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    put32(u);
  };
  put32((uint32_t)t.out_shape.size());
  for (int32_t s : t.out_shape) put32((uint32_t)s);
  put32((uint32_t)order.size());
  for (uint32_t i : order) {
    const Node& nd = t.nodes[i];
    put32(nd.kind);
    switch (nd.kind) {
      case K_LITERAL: putf(nd.value); break;
      case K_PARAM:
        put32((uint32_t)param_ordinal[i]);
        putf(nd.value);
        put32((uint32_t)nd.shape.size());
        for (int32_t s : nd.shape) put32((uint32_t)s);
        put32(nd.def_root >= 0 ? 1u : 0u);
        break;
      case K_TRANSFORM:
        put32((uint32_t)canon[nd.kids[0]]);
        put32(nd.rows);
        put32(nd.cols);
        key.append((const char*)nd.matrix.data(), nd.matrix.size() * sizeof(double));
        break;
      case K_REDUCE:
        put32(nd.monoid);
        put32((uint32_t)canon[nd.kids[0]]);
        put32((uint32_t)nd.shape.size());
        for (int32_t s : nd.shape) put32((uint32_t)s);
        break;
      case K_CONCAT:
        put32((uint32_t)nd.kids.size());
        for (uint32_t k : nd.kids) put32((uint32_t)canon[k]);
        put32((uint32_t)(nd.position + 1));
        break;
      default:
        put32((uint32_t)nd.kids.size());
        for (uint32_t k : nd.kids) put32((uint32_t)canon[k]);
    }
  }
  for (auto& d : defs) {
    put32(0xDEF00000u | d.first);
    put32((uint32_t)canon[d.second]);
  }
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : key) {
    h ^= c;
    h *= 1099511628211ull;
  }
  t.hash = h;
}

// ---- writer -------------------------------------------------------------------------------------------------

void TreeWriter::u32(uint32_t v) { body_.append((const char*)&v, 4); }
void TreeWriter::f32(float v) { body_.append((const char*)&v, 4); }
void TreeWriter::f64(double v) { body_.append((const char*)&v, 8); }
uint32_t TreeWriter::begin(uint32_t kind) {
  offsets_.push_back(body_.size());
  def_field_.push_back(0);
  u32(kind);
  return (uint32_t)offsets_.size() - 1;
}
uint32_t TreeWriter::literal(float v) {
  uint32_t i = begin(K_LITERAL);
  f32(v);
  return i;
}
uint32_t TreeWriter::parameter(uint64_t id, float padding, const std::vector<int32_t>& shape, int32_t def_root) {
  uint32_t i = begin(K_PARAM);
  body_.append((const char*)&id, 8);
  f32(padding);
  u32((uint32_t)shape.size());
  for (int32_t s : shape) i32(s);
  def_field_[i] = body_.size();
  i32(def_root);
  return i;
}
void TreeWriter::set_definition(uint32_t param_node, int32_t def_root) {
  size_t off = def_field_.at(param_node);
  CC_REQUIRE(off != 0, CC_ERR_BAD_TREE, "set_definition on a non-parameter node");
  memcpy(&body_[off], &def_root, 4);
}
uint32_t TreeWriter::transform(uint32_t array, uint32_t rows, uint32_t cols, const double* m) {
  uint32_t i = begin(K_TRANSFORM);
  u32(array);
  u32(rows);
  u32(cols);
  for (uint32_t k = 0; k < rows * cols; ++k) f64(m[k]);
  return i;
}
uint32_t TreeWriter::extract(uint32_t array) {
  uint32_t i = begin(K_EXTRACT);
  u32(array);
  return i;
}
uint32_t TreeWriter::concatenate(const std::vector<uint32_t>& elements, int32_t position) {
  uint32_t i = begin(position < 0 ? K_CONCAT : K_CONCAT_AT);
  u32((uint32_t)elements.size());
  if (position >= 0) u32((uint32_t)position);
  for (uint32_t e : elements) u32(e);
  return i;
}
uint32_t TreeWriter::unary(uint32_t kind, uint32_t a) {
  uint32_t i = begin(kind);
  u32(a);
  return i;
}
uint32_t TreeWriter::binary(uint32_t kind, uint32_t a, uint32_t b) {
  uint32_t i = begin(kind);
  u32(a);
  u32(b);
  return i;
}
uint32_t TreeWriter::reduce(uint32_t monoid, uint32_t operand, const std::vector<int32_t>& operand_shape) {
  uint32_t i = begin(K_REDUCE);
  u32(monoid);
  u32(operand);
  u32((uint32_t)operand_shape.size());
  for (int32_t s : operand_shape) i32(s);
  return i;
}
std::string TreeWriter::finish(uint32_t root, const std::vector<int32_t>& out_shape) const {
  std::string out;
  out.reserve(16 + out_shape.size() * 4 + body_.size());
  auto put = [&](uint32_t v) { out.append((const char*)&v, 4); };
  put(kMagic);
  put((uint32_t)offsets_.size());
  put(root);
  put((uint32_t)out_shape.size());
  for (int32_t s : out_shape) put((uint32_t)s);
  out += body_;
  return out;
}

}  // namespace cc
