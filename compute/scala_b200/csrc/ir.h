// ir.h — the serialised form of the reference's expression trees (Trees.scala) as the backend sees them.
//
// Node kinds are 1:1 with the case classes reachable from the Tensor API:
// FloatLiteral R:373-380, ArrayParameter R:755-823, Transform R:676-690, Extract R:660-672,
// Concatenate R:953-973, unary R:384-470/620-658, binary R:472-618.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.h"

namespace cc {

enum Kind : uint32_t {
  K_LITERAL = 1,
  K_PARAM = 2,
  K_TRANSFORM = 3,
  K_EXTRACT = 4,
  K_CONCAT = 5,
  // Concatenate whose element index lands at output dimension `position` instead of last: Tensor.join(tensors, dimension)
  // (Tensors.scala:560-575) in ONE kernel; the reference materialises the join and gathers a permuted view of it. Parsed into
  // a K_CONCAT node with `position` set.
  K_CONCAT_AT = 6,
  K_EXP = 10,
  K_LOG = 11,
  K_ABS = 12,
  K_TANH = 13,
  K_SQRT = 14,
  K_NEG = 15,
  K_MIN = 20,
  K_MAX = 21,
  K_PLUS = 22,
  K_MINUS = 23,
  K_TIMES = 24,
  K_DIV = 25,
  K_PERCENT = 26,
  // Root-only: fold the operand over its whole index space with a monoid (K_PLUS / K_MIN / K_MAX / K_TIMES). The reference
  // has no such tree node — Tensor.sum materialises its operand and runs a hand-written program over the buffer
  // (Tensors.scala:303-393, 673-771); MonoidPrograms is generic over append / zero, and this node lets the backend fuse the
  // operand's closure into the reduction.
  K_REDUCE = 30,
};

inline bool is_unary(uint32_t k) { return k >= K_EXP && k <= K_NEG; }
inline bool is_binary(uint32_t k) { return k >= K_MIN && k <= K_PERCENT; }
const char* kind_name(uint32_t k);

struct Node {
  uint32_t kind = 0;
  float value = 0.f;           // literal value, or padding of a parameter
  uint64_t param_id = 0;       // identity of the producing tensor (Tensors.scala:1259)
  SmallVec<int32_t, 4> shape;  // parameter shape
  int32_t def_root = -1;       // optional closure of the producing (not yet evaluated) inline tensor
  uint32_t rows = 0, cols = 0; // transform matrix, row-major rows x cols (cols = view rank + 1)
  std::vector<double> matrix;
  SmallVec<uint32_t, 2> kids;  // operands / array / concatenate elements (inline for every node but a long Concatenate)
  int32_t position = -1;       // K_CONCAT: output dimension of the element index (-1 = last, Tensors.scala:577-598)
  uint32_t monoid = 0;         // K_REDUCE: K_PLUS / K_MIN / K_MAX / K_TIMES (shape = index space of the operand)
};

struct Tree {
  std::vector<Node> nodes;  // children before parents
  uint32_t root = 0;
  std::vector<int32_t> out_shape;

  // Filled by canonicalize():
  std::string key;                  // structural identity, parameters by first-visit ordinal (R:70-91, 152-177)
  uint64_t hash = 0;
  std::vector<uint32_t> params;     // node index per parameter ordinal: DFS pre-order of the main tree
                                    // (= parameterDescendants, Tensors.scala:230-251), then definitions' parameters
  uint32_t n_main_params = 0;
};

// Blob <-> Tree. parse validates every index / size and throws CC_ERR_BAD_TREE.
Tree parse_tree(const void* blob, uint64_t n_bytes);
void canonicalize(Tree& t);

// Incremental writer used by the host-side mirror (tensor.cpp) — the same bytes a JVM front end would write.
class TreeWriter {
 public:
  TreeWriter() {
    body_.reserve(512);
    offsets_.reserve(32);
    def_field_.reserve(32);
  }
  uint32_t literal(float v);
  uint32_t parameter(uint64_t id, float padding, const std::vector<int32_t>& shape, int32_t def_root = -1);
  uint32_t transform(uint32_t array, uint32_t rows, uint32_t cols, const double* m);
  uint32_t extract(uint32_t array);
  uint32_t concatenate(const std::vector<uint32_t>& elements, int32_t position = -1);
  uint32_t unary(uint32_t kind, uint32_t a);
  uint32_t binary(uint32_t kind, uint32_t a, uint32_t b);
  uint32_t reduce(uint32_t monoid, uint32_t operand, const std::vector<int32_t>& operand_shape);
  void set_definition(uint32_t param_node, int32_t def_root);
  std::string finish(uint32_t root, const std::vector<int32_t>& out_shape) const;
  uint32_t size() const { return (uint32_t)offsets_.size(); }

 private:
  uint32_t begin(uint32_t kind);
  void u32(uint32_t v);
  void i32(int32_t v) { u32((uint32_t)v); }
  void f32(float v);
  void f64(double v);
  std::string body_;
  std::vector<size_t> offsets_;
  std::vector<size_t> def_field_;  // byte offset of definition_root for parameter nodes (else 0)
};

}  // namespace cc
