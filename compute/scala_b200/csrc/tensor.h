// tensor.h — host-side mirror of the reference's lazy Tensor API (`object cuda`, the sibling of cpu.scala:103-117 /
// gpu.scala:15-27), written in C++ because the reference's own toolchain (Scala/JVM) is not available here.
// Same names, argument meaning and error behaviour as Tensors.scala:395-1442; everything device-side goes through the
// cc_* C ABI (include/compute_cuda.h), exactly as the Scala `cuda` object would.
#pragma once
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.h"
#include "ir.h"

namespace compute {
namespace cuda {

using Shape = std::vector<int32_t>;
class Tensor;
using TensorPtr = std::shared_ptr<Tensor>;

// PendingBuffer (Tensors.scala:253-299): a device buffer plus the event that completes it.
struct PendingBuffer {
  cc_buffer buffer = 0;
  cc_event event = 0;  // 0 = ReadyBuffer
};

// One slow action = one session: every tensor is evaluated at most once per session (`.shared`, Tensors.scala:1401-1403)
// and its buffer is released when the session ends (deterministic release, README.md:10).
class Session {
 public:
  ~Session();
  // a slow action touches a handful of tensors (the kernel's arguments): a flat table searched linearly beats a hash map's node
  // allocations; graphs with many non-inline inputs fall over to the index
  PendingBuffer* find(const Tensor* t);
  PendingBuffer* add(const Tensor* t, const PendingBuffer& p);

 private:
  struct Entry {
    const Tensor* tensor;
    PendingBuffer buffer;
  };
  cc::SmallVec<Entry, 8> done_;
  std::unordered_map<const Tensor*, size_t> index_;  // built once done_ outgrows kLinear
  static constexpr size_t kLinear = 16;
};

class Tensor : public std::enable_shared_from_this<Tensor> {
 public:
  Shape shape;
  float padding = 0.f;
  // how this tensor is spread over the ranks of the communicator: the whole tensor on every rank (also: no communicator),
  // this rank's block of rows of a tensor sharded along its leading axis, or this rank's additive contribution to a sum over that axis
  enum Distribution : uint8_t { kWhole = 0, kRowBlock = 1, kPartialSum = 2 };
  Distribution dist = kWhole;
  virtual ~Tensor();

  int64_t size() const;

  // ---- closures (Tensors.scala:1254-1260, 1394-1440) ----
  virtual bool is_inline() const = 0;
  // emits this tensor's closure (a float term) into `w`; `ctx` memoises per tensor so shared sub-graphs stay shared
  struct EmitCtx;
  uint32_t closure(EmitCtx& ctx) const;
  virtual void closure_operands(std::vector<const Tensor*>& out) const { (void)out; }
  virtual uint32_t emit_closure(EmitCtx& ctx, const std::vector<uint32_t>& operands) const = 0;

  // ---- evaluation ----
  // doBuffer (Tensors.scala:1401-1403 etc.): the returned buffer is retained for the caller
  PendingBuffer do_buffer(Session& s) const;
  // the same evaluation, but the handle stays owned by the session (valid until the session ends): what a kernel launch inside
  // the session needs for its arguments, without a retain / release pair per argument
  cc_buffer borrow_buffer(Session& s) const;
  // a tensor that owns its device buffer for its whole life (Tensor.apply, doCache) lends it without any reference counting: the
  // graph being evaluated keeps the tensor, and so the buffer, alive for the session
  virtual cc_buffer resident_buffer() const { return 0; }
  virtual PendingBuffer evaluate(Session& s) const = 0;
  // evaluate with the kernel storing straight into `out` (a wrapped, device-visible buffer — pinned host memory for small
  // read-backs); `*out_event` completes when `out` is written. false = this tensor has no kernel of its own to redirect
  virtual bool evaluate_into(Session& s, cc_buffer out, cc_event* out_event) const {
    (void)s, (void)out, (void)out_event;
    return false;
  }
  // compile only (no device work): the kernel this tensor's closure maps to
  cc_kernel compile_only() const;
  std::string tree_blob() const;  // the blob compile_only() hands to cc_compile_ex
  uint32_t emit_root_for_compile(EmitCtx& ctx) const;

  // ---- delayed operators (Tensors.scala:816-1074) ----
  TensorPtr broadcast(const Shape& new_shape);
  TensorPtr reshape(const Shape& new_shape);
  TensorPtr scale(const Shape& new_shape);
  TensorPtr translate(const std::vector<double>& offset);
  TensorPtr translate(const std::vector<double>& offset, const Shape& new_shape);
  TensorPtr permute(const std::vector<int32_t>& dimensions);
  TensorPtr transpose();
  std::vector<TensorPtr> split(int dimension);
  TensorPtr sum();
  // reduce(MonoidPrograms) (Tensors.scala:308-311, 673-766) for Plus / Min / Max / Times
  TensorPtr reduce(uint32_t monoid);
  virtual TensorPtr non_inline();
  TensorPtr do_cache();
  TensorPtr transform(const Shape& new_shape, const std::vector<double>& matrix1);
  // ---- sharding (include/compute_cuda.h: ct_shard / ct_gather) ----
  TensorPtr as_row_block();            // declare this (whole, non-scalar) tensor to be this rank's row block
  TensorPtr combined();                // partial sum -> whole: evaluates, then all-reduces across ranks (identity for anything else)
  TensorPtr gather(bool zero_copy);    // row block -> whole tensor on every rank

  // ---- slow actions (Tensors.scala:776-811, 1099-1118) ----
  std::vector<float> flat_array() const;
  void flat_array_into(float* host, uint64_t capacity) const;
  // evaluate into pinned host memory from cc_host_alloc; small results are stored there by the kernel itself (no copy command)
  void read_into_pinned(float* pinned, uint64_t n) const;
  std::string to_string() const;
};

// object Tensor (Tensors.scala:395-600)
TensorPtr from_host(const float* data, const Shape& shape, float padding = 0.f);       // apply
TensorPtr from_buffer(cc_buffer buf, const Shape& shape, float padding = 0.f);
TensorPtr scalar(float value, float padding = 0.f);
TensorPtr fill(float value, const Shape& shape, float padding = 0.f);
TensorPtr random(const Shape& shape, int32_t seed, float padding = 0.f);
TensorPtr random_normal(const Shape& shape, int32_t seed, float padding = 0.f);
TensorPtr unary(uint32_t kind, const TensorPtr& t);                                    // abs sqrt tanh exp log unary_-
TensorPtr binary(uint32_t kind, const TensorPtr& l, const TensorPtr& r);               // + - * / % min max
TensorPtr join(const std::vector<TensorPtr>& tensors);
TensorPtr join(const std::vector<TensorPtr>& tensors, int dimension);
Shape auto_broadcast_shape(const Shape& a, const Shape& b);                            // Tensors.scala:208-222
std::string java_float_to_string(float v);

int64_t live_tensors();

}  // namespace cuda
}  // namespace compute
