// tensor.cpp — host-side mirror of Tensors.scala's lazy Tensor API on top of the cc_* C ABI, plus its flat ct_* C view.
#include "tensor.h"

#include <atomic>
#include <charconv>
#include <cmath>
#include <cstring>
#include <mutex>

#include "ndat.h"

namespace compute {
namespace cuda {

using cc::Error;
using cc::fail;
using cc::strprintf;

namespace {
std::atomic<int64_t> g_live{0};

void check(int status) {
  if (status != CC_OK) throw Error(status, cc_last_error());
}

std::string shape_str(const Shape& s) {
  std::string r = "[";
  for (size_t i = 0; i < s.size(); ++i) r += (i ? "," : "") + std::to_string(s[i]);
  return r + "]";
}

int64_t product(const Shape& s) {
  int64_t p = 1;
  for (int32_t d : s) p *= d;
  return p;
}

void check_shape(const Shape& s) {
  for (int32_t d : s) CC_REQUIRE(d >= 0, CC_ERR_ILLEGAL_ARGUMENT, "negative dimension in shape %s", shape_str(s).c_str());
  CC_REQUIRE(s.size() <= 16, CC_ERR_ILLEGAL_ARGUMENT, "rank %zu > 16", s.size());
}

// deep expression chains (a 16384-term per-axis sum) must not recurse on destruction
thread_local std::vector<TensorPtr> g_graveyard;
thread_local bool g_draining = false;
void bury(TensorPtr&& p) {
  if (!p) return;
  g_graveyard.push_back(std::move(p));
  if (g_draining) return;
  g_draining = true;
  while (!g_graveyard.empty()) {
    TensorPtr victim = std::move(g_graveyard.back());
    g_graveyard.pop_back();
    victim.reset();
  }
  g_draining = false;
}
}  // namespace

int64_t live_tensors() { return g_live.load(); }

Session::~Session() {
  for (Entry& e : done_) {
    if (e.buffer.buffer) cc_buffer_release(e.buffer.buffer);
    if (e.buffer.event) cc_event_release(e.buffer.event);
  }
}
PendingBuffer* Session::find(const Tensor* t) {
  if (done_.size() <= kLinear) {
    for (Entry& e : done_)
      if (e.tensor == t) return &e.buffer;
    return nullptr;
  }
  auto it = index_.find(t);
  return it == index_.end() ? nullptr : &done_[it->second].buffer;
}
PendingBuffer* Session::add(const Tensor* t, const PendingBuffer& p) {
  done_.push_back(Entry{t, p});
  if (done_.size() == kLinear + 1)
    for (size_t i = 0; i < done_.size(); ++i) index_.emplace(done_[i].tensor, i);
  else if (done_.size() > kLinear + 1)
    index_.emplace(t, done_.size() - 1);
  return &done_.back().buffer;
}

// ---- closure emission ----------------------------------------------------------------------------------------------------

// tensor -> node index. Expressions evaluated once and thrown away (the reference's benchmarks rebuild theirs on every call) have a
// handful of nodes: a flat array searched linearly, with a hash index only once the graph outgrows it (chains of thousands of terms).
class NodeOfTensor {
 public:
  const uint32_t* find(const Tensor* t) const {
    if (entries_.size() <= kLinear) {
      for (const Entry& e : entries_)
        if (e.tensor == t) return &e.node;
      return nullptr;
    }
    auto it = index_.find(t);
    return it == index_.end() ? nullptr : &entries_[it->second].node;
  }
  bool count(const Tensor* t) const { return find(t) != nullptr; }
  uint32_t at(const Tensor* t) const {
    const uint32_t* n = find(t);
    CC_REQUIRE(n, CC_ERR_BAD_TREE, "tensor has no emitted node");
    return *n;
  }
  void set(const Tensor* t, uint32_t node) {  // t is not present yet
    entries_.push_back(Entry{t, node});
    if (entries_.size() == kLinear + 1)
      for (size_t i = 0; i < entries_.size(); ++i) index_.emplace(entries_[i].tensor, i);
    else if (entries_.size() > kLinear + 1)
      index_.emplace(t, entries_.size() - 1);
  }

 private:
  struct Entry {
    const Tensor* tensor;
    uint32_t node;
  };
  static constexpr size_t kLinear = 24;
  cc::SmallVec<Entry, 24> entries_;
  std::unordered_map<const Tensor*, size_t> index_;
};

struct Tensor::EmitCtx {
  cc::TreeWriter w;
  NodeOfTensor closures;
  NodeOfTensor param_nodes;
  std::vector<const Tensor*> params;  // in creation order
  EmitCtx() { params.reserve(8); }

  uint32_t param(const Tensor* t) {
    if (const uint32_t* known = param_nodes.find(t)) return *known;
    // ArrayParameter(id = the tensor itself, padding, shape) — Tensors.scala:1254-1260
    uint32_t n = w.parameter((uint64_t)(uintptr_t)t, t->padding, t->shape, -1);
    param_nodes.set(t, n);
    params.push_back(t);
    return n;
  }
};

Tensor::~Tensor() { g_live.fetch_sub(1); }

int64_t Tensor::size() const { return product(shape); }

uint32_t Tensor::closure(EmitCtx& ctx) const {
  struct Frame {
    const Tensor* t;
    bool expanded;
  };
  cc::SmallVec<Frame, 16> stack{Frame{this, false}};
  std::vector<const Tensor*> ops;
  std::vector<uint32_t> ids;
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    const Tensor* t = f.t;
    if (ctx.closures.count(t)) continue;
    ops.clear();
    t->closure_operands(ops);
    if (!f.expanded && !ops.empty()) {
      stack.push_back(Frame{t, true});
      for (size_t k = ops.size(); k-- > 0;)
        if (!ctx.closures.count(ops[k])) stack.push_back(Frame{ops[k], false});
      continue;
    }
    ids.clear();
    for (const Tensor* o : ops) ids.push_back(ctx.closures.at(o));
    ctx.closures.set(t, t->emit_closure(ctx, ids));
  }
  return ctx.closures.at(this);
}

namespace {

// ---- concrete tensors ------------------------------------------------------------------------------------------------------

struct Counted : Tensor {
  Counted() { g_live.fetch_add(1); }
};

struct PlanCache {
  std::mutex mu;
  cc_kernel kernel = 0;
  std::vector<const Tensor*> args;  // kept alive by the owning tensor's own sub-graph
  ~PlanCache() {
    if (kernel) cc_kernel_release(kernel);
  }
};
void resolve_plan(PlanCache& pc, Tensor::EmitCtx& ctx, uint32_t root, const Shape& out_shape);
PendingBuffer enqueue_plan(Session& s, const PlanCache& pc, const Shape& out_shape, cc_buffer out_override = 0, cc_event* out_event = nullptr,
                           bool allreduce = false);
bool plan_output_redirectable(const PlanCache& pc);

// InlineTensor (Tensors.scala:1400-1411)
struct InlineTensor : Counted {
  mutable PlanCache plan;
  bool is_inline() const override { return true; }
  PendingBuffer evaluate(Session& s) const override {
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        uint32_t root = closure(ctx);
        resolve_plan(plan, ctx, root, shape);
      }
    }
    // a partial sum (the fold of a row block's split(0) views) is evaluated COMBINED: the library runs the kernel and the all-reduce, as ONE
    // launch when the reduction can complete the exchange itself (cc_shard_launch_allreduce)
    return enqueue_plan(s, plan, shape, 0, nullptr, dist == kPartialSum);
  }
  bool evaluate_into(Session& s, cc_buffer out, cc_event* out_event) const override {
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        uint32_t root = closure(ctx);
        resolve_plan(plan, ctx, root, shape);
      }
    }
    if (!plan_output_redirectable(plan)) return false;
    enqueue_plan(s, plan, shape, out, out_event);
    return true;
  }
};

// FillTensor (Tensors.scala:1394-1397)
struct FillTensor final : InlineTensor {
  float value;
  uint32_t emit_closure(EmitCtx& ctx, const std::vector<uint32_t>&) const override { return ctx.w.literal(value); }
};

// derivedTensor (Tensors.scala:857-863) of a unary / binary float term
struct DerivedTensor final : InlineTensor {
  uint32_t kind;
  TensorPtr a, b;
  ~DerivedTensor() override {
    bury(std::move(a));
    bury(std::move(b));
  }
  void closure_operands(std::vector<const Tensor*>& out) const override {
    out.push_back(a.get());
    if (b) out.push_back(b.get());
  }
  uint32_t emit_closure(EmitCtx& ctx, const std::vector<uint32_t>& o) const override {
    return b ? ctx.w.binary(kind, o[0], o[1]) : ctx.w.unary(kind, o[0]);
  }
};

// TransformedTensor (Tensors.scala:1413-1428): closure = checkpoint.arrayTerm.transform(matrix).extract
struct TransformedTensor final : InlineTensor {
  TensorPtr checkpoint;
  std::vector<double> matrix;  // checkpoint rank x (own rank + 1)
  ~TransformedTensor() override { bury(std::move(checkpoint)); }
  uint32_t emit_closure(EmitCtx& ctx, const std::vector<uint32_t>&) const override {
    uint32_t p = ctx.param(checkpoint.get());
    uint32_t rows = (uint32_t)checkpoint->shape.size(), cols = (uint32_t)shape.size() + 1;
    return ctx.w.extract(ctx.w.transform(p, rows, cols, matrix.data()));
  }
};

// NonInlineTensor (Tensors.scala:1431-1440): closure = arrayTerm.extract
struct NonInlineTensor : Counted {
  bool is_inline() const override { return false; }
  uint32_t emit_closure(EmitCtx& ctx, const std::vector<uint32_t>&) const override { return ctx.w.extract(ctx.param(this)); }
  TensorPtr non_inline() override { return shared_from_this(); }
};

// Tensor.apply / CachedTensor: owns a device buffer
struct BufferTensor final : NonInlineTensor {
  cc_buffer buffer = 0;
  ~BufferTensor() override {
    if (buffer) cc_buffer_release(buffer);
  }
  cc_buffer resident_buffer() const override { return buffer; }
  PendingBuffer evaluate(Session&) const override {
    check(cc_buffer_retain(buffer));
    return {buffer, 0};
  }
};

// reshape (Tensors.scala:879-888) and InlineTensor.nonInline (:1405-1410): same doBuffer, new shape
struct AliasTensor final : NonInlineTensor {
  TensorPtr base;
  ~AliasTensor() override { bury(std::move(base)); }
  PendingBuffer evaluate(Session& s) const override { return base->do_buffer(s); }
};

struct RandomTensor final : NonInlineTensor {
  int32_t seed;
  bool normal;
  PendingBuffer evaluate(Session&) const override {
    cc_buffer out = 0;
    uint64_t n = (uint64_t)size();
    check(cc_buffer_alloc(n, &out));
    int st = normal ? cc_random_normal(out, n, seed, nullptr) : cc_random(out, n, seed, nullptr);
    if (st != CC_OK) {
      std::string m = cc_last_error();
      cc_buffer_release(out);
      throw Error(st, m);
    }
    return {out, 0};
  }
};

// Tensor.sum / reduce(MonoidPrograms) (Tensors.scala:303-393, 673-771). The reference always materialises the operand and
// folds the buffer; here an inline operand's closure is fused into the fold kernel (one pass over the inputs, no intermediate),
// and a materialised operand summed with Plus runs the precompiled reduce_sum_kernel — both fold in the same order.
struct ReduceTensor final : NonInlineTensor {
  TensorPtr base;
  uint32_t monoid = cc::K_PLUS;
  bool across_ranks = false;  // the operand is a row block: the fold is completed by an all-reduce of its one float (Plus only)
  mutable PlanCache plan;
  ~ReduceTensor() override { bury(std::move(base)); }
  PendingBuffer evaluate(Session& s) const override {
    if (monoid == cc::K_PLUS && !base->is_inline()) {
      PendingBuffer in = base->do_buffer(s);
      cc_buffer out = 0;
      int st = cc_buffer_alloc(1, &out);
      // a row block: ONE kernel folds the block and all-reduces the result over NVLink peer memory (NCCL if the mailboxes are not mapped)
      if (st == CC_OK)
        st = across_ranks ? cc_reduce_sum_allreduce(in.buffer, (uint64_t)base->size(), out, nullptr, 0, nullptr)
                          : cc_reduce_sum(in.buffer, (uint64_t)base->size(), out, nullptr, 0, nullptr);
      std::string m = st == CC_OK ? "" : cc_last_error();
      cc_buffer_release(in.buffer);
      if (st != CC_OK) {
        if (out) cc_buffer_release(out);
        throw Error(st, m);
      }
      return {out, 0};
    }
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        resolve_plan(plan, ctx, emit_root(ctx), shape);
      }
    }
    return enqueue_plan(s, plan, shape, 0, nullptr, across_ranks);
  }
  bool evaluate_into(Session& s, cc_buffer out, cc_event* out_event) const override {
    if (across_ranks) return false;  // the combine runs on a device buffer
    if (monoid == cc::K_PLUS && !base->is_inline()) {
      PendingBuffer in = base->do_buffer(s);
      int st = cc_reduce_sum(in.buffer, (uint64_t)base->size(), out, nullptr, 0, out_event);
      std::string m = st == CC_OK ? "" : cc_last_error();
      cc_buffer_release(in.buffer);
      if (st != CC_OK) throw Error(st, m);
      return true;
    }
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        resolve_plan(plan, ctx, emit_root(ctx), shape);
      }
    }
    if (!plan_output_redirectable(plan)) return false;
    enqueue_plan(s, plan, shape, out, out_event);
    return true;
  }
  uint32_t emit_root(EmitCtx& ctx) const { return ctx.w.reduce(monoid, base->closure(ctx), base->shape); }
};

// Tensor.join (Tensors.scala:577-598): one kernel over the head shape whose root is Concatenate(elements)
struct JoinTensor final : NonInlineTensor {
  std::vector<TensorPtr> tensors;
  int32_t position = -1;  // output dimension of the element index; -1 = last
  ~JoinTensor() override {
    for (auto& t : tensors) bury(std::move(t));
  }
  uint32_t emit_root(EmitCtx& ctx) const {
    std::vector<uint32_t> e;
    for (auto& t : tensors) e.push_back(t->closure(ctx));
    return ctx.w.concatenate(e, position);
  }
  mutable PlanCache plan;
  PendingBuffer evaluate(Session& s) const override {
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        uint32_t root = emit_root(ctx);
        resolve_plan(plan, ctx, root, shape);
      }
    }
    return enqueue_plan(s, plan, shape);
  }
  bool evaluate_into(Session& s, cc_buffer out, cc_event* out_event) const override {
    {
      std::lock_guard<std::mutex> lock(plan.mu);
      if (!plan.kernel) {
        EmitCtx ctx;
        uint32_t root = emit_root(ctx);
        resolve_plan(plan, ctx, root, shape);
      }
    }
    if (!plan_output_redirectable(plan)) return false;
    enqueue_plan(s, plan, shape, out, out_event);
    return true;
  }
};

std::string finish_blob(Tensor::EmitCtx& ctx, uint32_t root, const Shape& out_shape, bool attach_definitions) {
  if (attach_definitions) {
    // let the code generator look through the fusion barrier: an ArrayParameter produced by a not-yet-evaluated
    // InlineTensor carries that tensor's closure (one level deep)
    std::vector<const Tensor*> snapshot = ctx.params;
    for (const Tensor* p : snapshot)
      if (p->is_inline()) {
        uint32_t def = p->closure(ctx);
        ctx.w.set_definition(ctx.param_nodes.at(p), (int32_t)def);
      }
  }
  return ctx.w.finish(root, out_shape);
}

cc_kernel compile_closure(Tensor::EmitCtx& ctx, uint32_t root, const Shape& out_shape, bool attach_definitions,
                          std::vector<uint64_t>* ids) {
  std::string blob = finish_blob(ctx, root, out_shape, attach_definitions);
  cc_kernel k = 0;
  int n = 0;
  std::vector<uint64_t> tmp(ctx.params.size() + 1);
  check(cc_compile_ex(blob.data(), blob.size(), &k, tmp.data(), (int)tmp.size(), &n));
  tmp.resize((size_t)n);
  if (ids) *ids = std::move(tmp);
  return k;
}

// The graph under a tensor is immutable, so the kernel its closure compiles to and the tensors that kernel takes as
// arguments never change: resolve them once per tensor. (The reference re-hashes the whole tree on every evaluation,
// Tensors.scala:1293 — O(nodes) per slow action, which for a 16384-term per-axis sum is ~10 ms of host time.)
void resolve_plan(PlanCache& pc, Tensor::EmitCtx& ctx, uint32_t root, const Shape& out_shape) {
  std::vector<uint64_t> ids;
  cc_kernel k = compile_closure(ctx, root, out_shape, true, &ids);
  try {
    cc_kernel_info_t info;
    check(cc_kernel_info(k, &info));
    std::vector<const Tensor*> args;
    args.reserve((size_t)info.n_args);
    for (int i = 0; i < info.n_args; ++i) {
      int32_t ord = -1;
      check(cc_kernel_arg_param(k, i, &ord));
      CC_REQUIRE(ord >= 0 && (size_t)ord < ids.size(), CC_ERR_BAD_TREE, "kernel argument refers to unknown parameter %d", ord);
      args.push_back((const Tensor*)(uintptr_t)ids[(size_t)ord]);
    }
    pc.args = std::move(args);
    pc.kernel = k;
  } catch (...) {
    cc_kernel_release(k);
    throw;
  }
}

// The contraction pipeline stores through TMA maps / rewrites its output in place; every other plan writes `out` exactly once
// with plain stores, so `out` may be any device-visible memory.
bool plan_output_redirectable(const PlanCache& pc) {
  cc_kernel_info_t info;
  check(cc_kernel_info(pc.kernel, &info));
  return info.kind != 2;
}

// enqueueClosure (Tensors.scala:1291-1392)
PendingBuffer enqueue_plan(Session& s, const PlanCache& pc, const Shape& out_shape, cc_buffer out_override, cc_event* out_event, bool allreduce) {
  cc::SmallVec<cc_buffer, 16> args;
  cc_buffer out = 0;
  try {
    // upvalues.traverse(tree.id.asInstanceOf[Tensor].doBuffer) — Tensors.scala:1336-1340; the session owns the references
    for (const Tensor* t : pc.args) args.push_back(t->borrow_buffer(s));
    if (out_override)
      out = out_override;
    else
      check(cc_buffer_alloc((uint64_t)product(out_shape), &out));
    if (allreduce)
      check(cc_shard_launch_allreduce(pc.kernel, args.data(), (int)args.size(), out, nullptr, 0, out_event));
    else
      check(cc_launch(pc.kernel, args.data(), (int)args.size(), out, nullptr, 0, out_event));
  } catch (...) {
    if (out && !out_override) cc_buffer_release(out);
    throw;
  }
  return {out, 0};
}

// ---- sharded tensors: combine / gather nodes ------------------------------------------------------------------------------------------

int comm_world() {
  int world = 1, rank = 0;
  check(cc_comm_info(&world, &rank));
  return world;
}

// partial sum -> whole tensor: a fusion barrier whose evaluation is (inline partial expression, evaluated) + all-reduce
struct AllReduceTensor final : NonInlineTensor {
  TensorPtr base;  // an inline partial sum
  ~AllReduceTensor() override { bury(std::move(base)); }
  PendingBuffer evaluate(Session& s) const override { return base->do_buffer(s); }  // borrow_buffer() of a partial sum all-reduces it
};

// row block -> the whole tensor on every rank
struct GatherTensor final : NonInlineTensor {
  TensorPtr base;
  bool zero_copy = false;
  ~GatherTensor() override { bury(std::move(base)); }
  PendingBuffer evaluate(Session& s) const override {
    const int world = comm_world();
    if (world == 1) return base->do_buffer(s);
    const uint64_t n = (uint64_t)base->size();
    // equal blocks on every rank, checked once per size (collective, so every rank fails together instead of hanging in the exchange)
    if (!per_communicator().agreed.count(n)) {
      int equal = 0;
      check(cc_shard_agree(n, &equal));
      CC_REQUIRE(equal, CC_ERR_ILLEGAL_ARGUMENT,
                 "gather needs equal row blocks on every rank (this rank holds %llu floats): pad the leading axis to a multiple of the number of ranks",
                 (unsigned long long)n);
      per_communicator().agreed.insert(n);
    }
    int peer = 0;
    check(cc_comm_peer_enabled(&peer));
    const InlineTensor* inl = dynamic_cast<const InlineTensor*>(base.get());
    if (inl && peer) {
      // the block's own kernel: a contraction stores every accumulator tile into every rank's copy from its epilogue
      {
        std::lock_guard<std::mutex> lock(inl->plan.mu);
        if (!inl->plan.kernel) {
          EmitCtx ctx;
          uint32_t root = inl->closure(ctx);
          resolve_plan(inl->plan, ctx, root, inl->shape);
        }
      }
      cc_kernel_info_t info;
      check(cc_kernel_info(inl->plan.kernel, &info));
      if (info.kind == 2 && inl->shape.size() == 2 && inl->shape[1] % 4 == 0) {
        cc_buffer arena = symmetric_arena(n * (uint64_t)world);
        cc::SmallVec<cc_buffer, 16> args;
        for (const Tensor* t : inl->plan.args) args.push_back(t->borrow_buffer(s));
        int fused = 0;
        check(cc_shard_launch_allgather(inl->plan.kernel, args.data(), (int)args.size(), arena, nullptr, 0, nullptr, &fused));
        if (zero_copy) {
          check(cc_buffer_retain(arena));
          return {arena, 0};
        }
        return {copy_of(arena, n * (uint64_t)world), 0};
      }
      // any other kernel of the block's own: the library launches it and gathers — in ONE launch when the kernel can collect the other ranks'
      // values itself (a row-owner reduction: `shard.split(1).reduce(_ + _).gather`), else kernel + one-shot all-gather
      cc_buffer whole = 0;
      check(cc_buffer_alloc(n * (uint64_t)world, &whole));
      cc::SmallVec<cc_buffer, 16> args;
      for (const Tensor* t : inl->plan.args) args.push_back(t->borrow_buffer(s));
      const int st = cc_shard_launch_allgather(inl->plan.kernel, args.data(), (int)args.size(), whole, nullptr, 0, nullptr, nullptr);
      if (st != CC_OK) {
        const std::string m = cc_last_error();
        cc_buffer_release(whole);
        throw Error(st, m);
      }
      return {whole, 0};
    }
    PendingBuffer part = base->do_buffer(s);
    cc_buffer whole = 0;
    int st = cc_buffer_alloc(n * (uint64_t)world, &whole);
    if (st == CC_OK) st = cc_allgather(part.buffer, whole, n, nullptr, 0, nullptr);
    std::string m = st == CC_OK ? "" : cc_last_error();
    cc_buffer_release(part.buffer);
    if (st != CC_OK) {
      if (whole) cc_buffer_release(whole);
      throw Error(st, m);
    }
    return {whole, 0};
  }

  // What is cached per communicator (collective evaluations run in the same order on every rank, one at a time): the block sizes all
  // ranks agreed on, and one symmetric (peer-mapped) arena per size — cuMemAlloc + an IPC handle exchange, far too slow per evaluation;
  // cc_shard_launch_allgather's entry barrier protects an arena against the previous gather's readers.
  struct PerCommunicator {
    uint64_t generation = 0;
    std::unordered_set<uint64_t> agreed;
    std::unordered_map<uint64_t, cc_buffer> arenas;
  };
  static PerCommunicator& per_communicator() {
    static PerCommunicator pc;
    uint64_t now = 0;
    check(cc_comm_generation(&now));
    if (now != pc.generation) {  // a new communicator: the old arenas' memory went with the old one
      for (auto& kv : pc.arenas) cc_buffer_release(kv.second);
      pc.arenas.clear();
      pc.agreed.clear();
      pc.generation = now;
    }
    return pc;
  }
  static cc_buffer symmetric_arena(uint64_t n_floats) {
    PerCommunicator& pc = per_communicator();
    auto it = pc.arenas.find(n_floats);
    if (it != pc.arenas.end()) return it->second;
    cc_buffer b = 0;
    check(cc_comm_symmetric_alloc(n_floats, &b));
    pc.arenas.emplace(n_floats, b);
    return b;
  }
  static cc_buffer copy_of(cc_buffer src, uint64_t n_floats) {
    cc_buffer dst = 0;
    check(cc_buffer_alloc(n_floats, &dst));
    int st = cc_buffer_copy(dst, src, n_floats, nullptr, 0, nullptr);
    if (st != CC_OK) {
      std::string m = cc_last_error();
      cc_buffer_release(dst);
      throw Error(st, m);
    }
    return dst;
  }
};

template <class T>
std::shared_ptr<T> make(const Shape& shape, float padding) {
  check_shape(shape);
  auto t = std::make_shared<T>();
  t->shape = shape;
  t->padding = padding;
  return t;
}

}  // namespace

cc_buffer Tensor::borrow_buffer(Session& s) const {
  if (const cc_buffer own = resident_buffer()) return own;
  PendingBuffer* p = s.find(this);
  if (!p) {
    PendingBuffer fresh = evaluate(s);
    if (dist == kPartialSum && !is_inline()) {
      // (a partial sum is an inline expression — split(0) views folded with + — and InlineTensor::evaluate has already combined it; this is
      // the route for any other producer: `fresh` is this evaluation's own output)
      int st = cc_allreduce_sum(fresh.buffer, (uint64_t)size(), nullptr, 0, nullptr);
      if (st != CC_OK) {
        std::string m = cc_last_error();
        cc_buffer_release(fresh.buffer);
        throw Error(st, m);
      }
    }
    p = s.add(this, fresh);  // the session keeps evaluate()'s reference
  }
  return p->buffer;
}

PendingBuffer Tensor::do_buffer(Session& s) const {
  const cc_buffer b = borrow_buffer(s);
  check(cc_buffer_retain(b));
  return {b, 0};
}

uint32_t Tensor::emit_root_for_compile(EmitCtx& ctx) const {
  if (auto j = dynamic_cast<const JoinTensor*>(this)) return j->emit_root(ctx);
  if (auto r = dynamic_cast<const ReduceTensor*>(this)) return r->emit_root(ctx);
  return closure(ctx);
}

cc_kernel Tensor::compile_only() const {
  EmitCtx ctx;
  uint32_t root = emit_root_for_compile(ctx);
  return compile_closure(ctx, root, shape, true, nullptr);
}

std::string Tensor::tree_blob() const {
  EmitCtx ctx;
  uint32_t root = emit_root_for_compile(ctx);
  return finish_blob(ctx, root, shape, true);
}

// ---- factories ----------------------------------------------------------------------------------------------------------------

TensorPtr from_buffer(cc_buffer buf, const Shape& shape, float padding) {
  auto t = make<BufferTensor>(shape, padding);
  uint64_t n = 0;
  check(cc_buffer_length(buf, &n));
  CC_REQUIRE((int64_t)n >= t->size(), CC_ERR_ILLEGAL_ARGUMENT, "buffer of %llu floats is smaller than shape %s", (unsigned long long)n,
             shape_str(shape).c_str());
  check(cc_buffer_retain(buf));
  t->buffer = buf;
  return t;
}

TensorPtr from_host(const float* data, const Shape& shape, float padding) {
  auto t = make<BufferTensor>(shape, padding);
  check(cc_buffer_from_host(data, (uint64_t)t->size(), &t->buffer, nullptr));
  return t;
}

TensorPtr fill(float value, const Shape& shape, float padding) {
  auto t = make<FillTensor>(shape, padding);
  t->value = value;
  return t;
}
TensorPtr scalar(float value, float padding) { return fill(value, {}, padding); }

TensorPtr random(const Shape& shape, int32_t seed, float padding) {
  auto t = make<RandomTensor>(shape, padding);
  t->seed = seed;
  t->normal = false;
  return t;
}
TensorPtr random_normal(const Shape& shape, int32_t seed, float padding) {
  auto t = make<RandomTensor>(shape, padding);
  t->seed = seed;
  t->normal = true;
  return t;
}

Shape auto_broadcast_shape(const Shape& s1, const Shape& s2) {
  Shape out;
  const size_t n = std::max(s1.size(), s2.size());
  for (size_t i = 0; i < n; ++i) {
    if (i >= s1.size() || s1[i] == 1) {
      // (a unit dimension of the longer shape facing no dimension at all: the reference indexes past the shorter array here and dies
      // with ArrayIndexOutOfBoundsException, Tensors.scala:210-211)
      CC_REQUIRE(i < s2.size(), CC_ERR_ILLEGAL_ARGUMENT, "Failed to automatically broadcast between shape %s and %s", shape_str(s1).c_str(),
                 shape_str(s2).c_str());
      out.push_back(s2[i]);
    } else if (i >= s2.size() || s2[i] == 1)
      out.push_back(s1[i]);
    else if (s1[i] == s2[i])
      out.push_back(s1[i]);
    else
      fail(CC_ERR_ILLEGAL_ARGUMENT, strprintf("Failed to automatically broadcast between shape %s and %s", shape_str(s1).c_str(),
                                              shape_str(s2).c_str()));
  }
  return out;
}

TensorPtr unary(uint32_t kind, const TensorPtr& t0) {
  CC_REQUIRE(t0, CC_ERR_ILLEGAL_ARGUMENT, "null tensor");
  CC_REQUIRE(cc::is_unary(kind), CC_ERR_ILLEGAL_ARGUMENT, "not a unary operator: %u", kind);
  TensorPtr t = t0->combined();  // f(partial sum) needs the sum
  auto d = make<DerivedTensor>(t->shape, t->padding);
  d->kind = kind;
  d->a = t;
  d->dist = t->dist;
  return d;
}

TensorPtr binary(uint32_t kind, const TensorPtr& l, const TensorPtr& r) {
  CC_REQUIRE(l && r, CC_ERR_ILLEGAL_ARGUMENT, "null tensor");
  CC_REQUIRE(cc::is_binary(kind), CC_ERR_ILLEGAL_ARGUMENT, "not a binary operator: %u", kind);
  // partial sums stay partial only under + with another partial sum (the fold of split(0) slices); anything else needs the sum itself
  const bool partial = kind == cc::K_PLUS && l->dist == Tensor::kPartialSum && r->dist == Tensor::kPartialSum;
  TensorPtr lc = partial ? l : l->combined(), rc = partial ? r : r->combined();
  Shape ns = auto_broadcast_shape(lc->shape, rc->shape);
  TensorPtr bl = lc->broadcast(ns), br = rc->broadcast(ns);
  auto d = make<DerivedTensor>(bl->shape, bl->padding);
  d->kind = kind;
  d->a = bl;
  d->b = br;
  // a row block combined with a replicated operand (B of the row-sharded matmul, a constant) is a row block
  d->dist = partial ? Tensor::kPartialSum : (bl->dist == Tensor::kRowBlock || br->dist == Tensor::kRowBlock) ? Tensor::kRowBlock : Tensor::kWhole;
  return d;
}

TensorPtr join(const std::vector<TensorPtr>& tensors) {
  CC_REQUIRE(!tensors.empty(), CC_ERR_ILLEGAL_ARGUMENT, "join of an empty sequence");
  for (auto& t : tensors) {
    CC_REQUIRE(t, CC_ERR_ILLEGAL_ARGUMENT, "null tensor");
    CC_REQUIRE(t->shape == tensors[0]->shape, CC_ERR_ILLEGAL_ARGUMENT, "join of tensors with different shapes %s and %s",
               shape_str(tensors[0]->shape).c_str(), shape_str(t->shape).c_str());
  }
  Shape s = tensors[0]->shape;
  s.push_back((int32_t)tensors.size());
  auto j = make<JoinTensor>(s, tensors[0]->padding);
  j->tensors.reserve(tensors.size());
  for (auto& t : tensors) {
    j->tensors.push_back(t->combined());
    if (j->tensors.back()->dist == Tensor::kRowBlock) j->dist = Tensor::kRowBlock;  // the new dimension is the LAST one: the leading axis stays
  }
  return j;
}

// Tensors.scala:560-575 is `join(tensors).permute(...)`: a join kernel, then a gather of the permuted view. Here the join kernel
// itself stores with the element index at `dimension` (one kernel, one pass; node ConcatenateAt).
TensorPtr join(const std::vector<TensorPtr>& tensors, int dimension) {
  TensorPtr j = join(tensors);
  const int n = (int)j->shape.size();
  CC_REQUIRE(dimension >= 0 && dimension < n, CC_ERR_ILLEGAL_ARGUMENT, "join dimension %d out of range", dimension);
  if (n - 1 == dimension) return j;
  CC_REQUIRE(!(j->dist == Tensor::kRowBlock && dimension == 0), CC_ERR_UNSUPPORTED,
             "join at dimension 0 would put the new dimension in front of the sharded leading axis: gather first");
  Shape s = tensors[0]->shape;
  s.insert(s.begin() + dimension, (int32_t)tensors.size());
  auto at = make<JoinTensor>(s, tensors[0]->padding);
  at->tensors = static_cast<JoinTensor*>(j.get())->tensors;
  at->position = dimension;
  at->dist = j->dist;
  return at;
}

// ---- delayed operators ---------------------------------------------------------------------------------------------------------

namespace {
// What a view `matrix1` (rows = dimensions of `t`, columns = dimensions of the view + constant) of a row block is:
// still a row block if the view's dimension 0 IS t's dimension 0 and nothing else touches it; a partial contribution if it fixes
// t's dimension 0 to a constant (split(0): the local rows, to be folded with +); anything else would need rows of other ranks.
Tensor::Distribution view_of_row_block(const Tensor& t, const Shape& new_shape, const std::vector<double>& m) {
  const size_t rows = t.shape.size(), cols = new_shape.size() + 1;
  CC_REQUIRE(rows >= 1 && m.size() == rows * cols, CC_ERR_ILLEGAL_ARGUMENT, "transform matrix has %zu entries, expected %zu", m.size(), rows * cols);
  bool row0_identity = cols >= 2 && m[0] == 1.0, row0_constant = true;
  for (size_t c = 0; c < cols; ++c) {
    if (c != 0 && m[c] != 0.0) row0_identity = false;
    if (c + 1 < cols && m[c] != 0.0) row0_constant = false;
  }
  bool others_use_dim0 = false;
  for (size_t r = 1; r < rows; ++r)
    if (cols >= 2 && m[r * cols] != 0.0) others_use_dim0 = true;
  if (row0_identity && !others_use_dim0 && !new_shape.empty() && new_shape[0] == t.shape[0]) return Tensor::kRowBlock;
  if (row0_constant) return Tensor::kPartialSum;
  fail(CC_ERR_UNSUPPORTED, strprintf("this view of a row block %s mixes the sharded leading axis (it moves, shifts, scales or broadcasts over "
                                     "dimension 0): gather it or use a replicated tensor (SURVEY 8e)", shape_str(t.shape).c_str()));
  return Tensor::kWhole;
}
}  // namespace

TensorPtr Tensor::transform(const Shape& new_shape, const std::vector<double>& matrix1) {
  if (dist == kPartialSum) return combined()->transform(new_shape, matrix1);  // (padding would be added once per rank)
  const Distribution view_dist = dist == kRowBlock ? view_of_row_block(*this, new_shape, matrix1) : kWhole;
  // Tensors.scala:978-1003: views of views are composed on the host and keep the ORIGINAL checkpoint
  if (auto tt = dynamic_cast<TransformedTensor*>(this)) {
    auto t = make<TransformedTensor>(new_shape, padding);
    t->matrix = cc::ndat::pre_concatenate(matrix1, tt->matrix, new_shape.size());
    t->checkpoint = tt->checkpoint;
    t->dist = view_dist;
    return t;
  }
  auto t = make<TransformedTensor>(new_shape, padding);
  t->dist = view_dist;
  CC_REQUIRE(matrix1.size() == shape.size() * (new_shape.size() + 1), CC_ERR_ILLEGAL_ARGUMENT, "transform matrix has %zu entries, expected %zu",
             matrix1.size(), shape.size() * (new_shape.size() + 1));
  t->matrix = matrix1;
  t->checkpoint = shared_from_this();
  return t;
}

TensorPtr Tensor::broadcast(const Shape& new_shape) {
  if (new_shape == shape) return shared_from_this();
  if (auto f = dynamic_cast<FillTensor*>(this)) return fill(f->value, new_shape, padding);
  const size_t nl = new_shape.size(), l = shape.size();
  CC_REQUIRE(nl >= l, CC_ERR_ILLEGAL_ARGUMENT, "Cannot broadcast %s to %s", shape_str(shape).c_str(), shape_str(new_shape).c_str());
  std::vector<double> m((nl + 1) * l, 0.0);
  for (size_t i = 0; i < l; ++i) {
    if (shape[i] == new_shape[i])
      m[i * (nl + 1) + i] = 1.0;
    else if (shape[i] != 1)
      fail(CC_ERR_ILLEGAL_ARGUMENT, strprintf("Cannot broadcast %s to %s", shape_str(shape).c_str(), shape_str(new_shape).c_str()));
  }
  return transform(new_shape, m);
}

TensorPtr Tensor::reshape(const Shape& new_shape) {
  check_shape(new_shape);
  CC_REQUIRE(product(new_shape) == size(), CC_ERR_ILLEGAL_ARGUMENT, "cannot reshape %s to %s", shape_str(shape).c_str(),
             shape_str(new_shape).c_str());
  if (dist == kPartialSum) return combined()->reshape(new_shape);
  CC_REQUIRE(dist != kRowBlock || (!new_shape.empty() && new_shape[0] == shape[0]), CC_ERR_UNSUPPORTED,
             "reshape of a row block %s to %s changes the sharded leading axis: gather first", shape_str(shape).c_str(), shape_str(new_shape).c_str());
  auto a = make<AliasTensor>(new_shape, padding);
  a->base = shared_from_this();
  a->dist = dist;
  return a;
}

TensorPtr Tensor::non_inline() {
  if (dist == kPartialSum) return combined();
  auto a = make<AliasTensor>(shape, padding);
  a->base = shared_from_this();
  a->dist = dist;
  return a;
}

TensorPtr Tensor::as_row_block() {
  CC_REQUIRE(dist == kWhole, CC_ERR_ILLEGAL_ARGUMENT, "the tensor is already sharded");
  CC_REQUIRE(!shape.empty(), CC_ERR_ILLEGAL_ARGUMENT, "a scalar has no leading axis to shard");
  auto a = make<AliasTensor>(shape, padding);
  a->base = shared_from_this();
  a->dist = kRowBlock;
  return a;
}

TensorPtr Tensor::combined() {
  if (dist != kPartialSum) return shared_from_this();
  auto a = make<AllReduceTensor>(shape, padding);
  a->base = shared_from_this();
  return a;
}

TensorPtr Tensor::gather(bool zero_copy) {
  if (dist == kPartialSum) return combined();
  if (dist == kWhole) return shared_from_this();
  Shape whole = shape;
  whole[0] = (int32_t)((int64_t)shape[0] * comm_world());
  auto g = make<GatherTensor>(whole, padding);
  g->base = shared_from_this();
  g->zero_copy = zero_copy;
  return g;
}

TensorPtr Tensor::scale(const Shape& new_shape) {
  const size_t n = new_shape.size();
  CC_REQUIRE(n == shape.size(), CC_ERR_ILLEGAL_ARGUMENT, "scale to a different rank");
  std::vector<double> m(n * (n + 1), 0.0);
  for (size_t i = 0; i < n; ++i) m[i * (n + 1) + i] = (double)shape[i] / (double)new_shape[i];
  return transform(new_shape, m);
}

TensorPtr Tensor::translate(const std::vector<double>& offset) { return translate(offset, shape); }

TensorPtr Tensor::translate(const std::vector<double>& offset, const Shape& new_shape) {
  CC_REQUIRE(offset.size() == shape.size(), CC_ERR_ILLEGAL_ARGUMENT, "translate offset has %zu entries for a rank-%zu tensor", offset.size(),
             shape.size());
  std::vector<double> neg(offset.size());
  for (size_t i = 0; i < offset.size(); ++i) neg[i] = -offset[i];
  return transform(new_shape, cc::ndat::translate(neg));
}

TensorPtr Tensor::permute(const std::vector<int32_t>& dimensions) {
  const size_t n = shape.size();
  CC_REQUIRE(dimensions.size() == n, CC_ERR_ILLEGAL_ARGUMENT, "permute with %zu dimensions on a rank-%zu tensor", dimensions.size(), n);
  Shape ns(n);
  std::vector<double> m(n * (n + 1), 0.0);
  for (size_t nd = 0; nd < n; ++nd) {
    int32_t od = dimensions[nd];
    CC_REQUIRE(od >= 0 && (size_t)od < n, CC_ERR_ILLEGAL_ARGUMENT, "permute dimension %d out of range", od);
    ns[nd] = shape[(size_t)od];
    m[(size_t)od * (n + 1) + nd] = 1.0;
  }
  return transform(ns, m);
}

TensorPtr Tensor::transpose() {
  std::vector<int32_t> d(shape.size());
  for (size_t i = 0; i < d.size(); ++i) d[i] = (int32_t)(d.size() - 1 - i);
  return permute(d);
}

std::vector<TensorPtr> Tensor::split(int dimension) {
  const int n = (int)shape.size();
  CC_REQUIRE(dimension >= 0 && dimension < n, CC_ERR_ILLEGAL_ARGUMENT, "split dimension %d out of range for rank %d", dimension, n);
  Shape ns;
  for (int i = 0; i < n; ++i)
    if (i != dimension) ns.push_back(shape[(size_t)i]);
  std::vector<TensorPtr> out;
  out.reserve((size_t)shape[(size_t)dimension]);
  for (int32_t index = 0; index < shape[(size_t)dimension]; ++index) {
    std::vector<double> m((size_t)n * (size_t)n, 0.0);
    for (int i = 0; i < dimension; ++i) m[(size_t)i * n + i] = 1.0;
    m[(size_t)dimension * n + n - 1] = (double)index;
    for (int i = dimension + 1; i < n; ++i) m[(size_t)i * n + i - 1] = 1.0;
    out.push_back(transform(ns, m));
  }
  return out;
}

TensorPtr Tensor::reduce(uint32_t monoid) {
  CC_REQUIRE(monoid == cc::K_PLUS || monoid == cc::K_MIN || monoid == cc::K_MAX || monoid == cc::K_TIMES, CC_ERR_ILLEGAL_ARGUMENT,
             "reduce needs a monoid: Plus, Min, Max or Times (got %u)", monoid);
  if (dist == kPartialSum) return combined()->reduce(monoid);
  CC_REQUIRE(dist != kRowBlock || monoid == cc::K_PLUS, CC_ERR_UNSUPPORTED, "only + is combined across ranks: gather a row block before reducing it with another monoid");
  auto s = make<ReduceTensor>({}, padding);
  s->base = shared_from_this();
  s->monoid = monoid;
  s->across_ranks = dist == kRowBlock;
  return s;
}

TensorPtr Tensor::sum() { return reduce(cc::K_PLUS); }

TensorPtr Tensor::do_cache() {
  Session s;
  PendingBuffer p = do_buffer(s);  // (a partial sum comes back all-reduced: whole)
  auto t = make<BufferTensor>(shape, padding);
  t->buffer = p.buffer;
  t->dist = dist == kRowBlock ? kRowBlock : kWhole;
  return t;
}

// ---- slow actions ---------------------------------------------------------------------------------------------------------------

// Results up to this many floats are stored into pinned host memory by the producing kernel itself: a 4 KB read-back costs one
// launch + one wait (~11 us) instead of launch + copy command + wait (~18 us; scripts/gpu_call_overhead.py). Beyond it the copy
// engine's PCIe efficiency wins.
constexpr uint64_t kDirectToHostFloats = 16384;

void Tensor::read_into_pinned(float* pinned, uint64_t n) const {
  if (n > 0 && n <= kDirectToHostFloats && dist != kPartialSum) {
    uint64_t dptr = 0;
    cc_buffer wrapped = 0;
    if (cc_host_device_ptr(pinned, &dptr) == CC_OK && cc_buffer_wrap(dptr, n, &wrapped) == CC_OK) {
      cc_event ev = 0;
      bool direct = false;
      try {
        Session s;
        direct = evaluate_into(s, wrapped, &ev);
        if (direct) check(cc_event_wait(ev));
      } catch (...) {
        if (ev) cc_event_release(ev);
        cc_buffer_release(wrapped);
        throw;
      }
      if (ev) cc_event_release(ev);
      cc_buffer_release(wrapped);
      if (direct) return;
    }
  }
  flat_array_into(pinned, n);
}

void Tensor::flat_array_into(float* host, uint64_t capacity) const {
  const uint64_t n = (uint64_t)size();
  CC_REQUIRE(capacity >= n, CC_ERR_ILLEGAL_ARGUMENT, "flatArray needs room for %llu floats, got %llu", (unsigned long long)n,
             (unsigned long long)capacity);
  if (n > 0 && n <= kDirectToHostFloats && dist != kPartialSum) {
    // small result: let the kernel store into pooled pinned memory, then one memcpy into the caller's (pageable) array
    void* pinned = nullptr;
    if (cc_host_alloc(n * 4, &pinned) == CC_OK) {
      uint64_t dptr = 0;
      cc_buffer wrapped = 0;
      bool direct = false;
      if (cc_host_device_ptr(pinned, &dptr) == CC_OK && cc_buffer_wrap(dptr, n, &wrapped) == CC_OK) {
        cc_event ev = 0;
        try {
          Session s;
          direct = evaluate_into(s, wrapped, &ev);
          if (direct) check(cc_event_wait(ev));
        } catch (...) {
          if (ev) cc_event_release(ev);
          cc_buffer_release(wrapped);
          cc_host_free(pinned);
          throw;
        }
        if (ev) cc_event_release(ev);
        cc_buffer_release(wrapped);
      }
      if (direct) memcpy(host, pinned, (size_t)n * 4);
      cc_host_free(pinned);
      if (direct) return;
    }
  }
  Session s;
  PendingBuffer p = do_buffer(s);
  int st = cc_buffer_to_host(p.buffer, 0, host, n, nullptr, 0, nullptr);
  std::string m = st == CC_OK ? "" : cc_last_error();
  cc_buffer_release(p.buffer);
  if (st != CC_OK) throw Error(st, m);
}

std::vector<float> Tensor::flat_array() const {
  std::vector<float> v((size_t)size());
  flat_array_into(v.data(), v.size());
  return v;
}

std::string java_float_to_string(float x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
  if (x == 0.f) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
  std::string s(buf, r.ptr);  // [-]d[.ddd]e[+-]XX, shortest round-trip digits
  bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  int exp = atoi(s.c_str() + epos + 1);
  std::string digits;
  for (char c : mant)
    if (c != '.') digits.push_back(c);
  std::string out;
  const double a = std::fabs((double)x);
  if (a >= 1e-3 && a < 1e7) {
    if (exp >= 0) {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp + 1));
      while ((int)ip.size() < exp + 1) ip.push_back('0');
      std::string fp = digits.size() > (size_t)exp + 1 ? digits.substr((size_t)exp + 1) : "0";
      out = ip + "." + fp;
    } else {
      out = "0." + std::string((size_t)(-exp - 1), '0') + digits;
    }
  } else {
    out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp);
  }
  return neg ? "-" + out : out;
}

std::string Tensor::to_string() const {
  std::vector<float> flat = flat_array();
  std::string out;
  // Tensors.scala:776-811
  struct Rec {
    const Shape& shape;
    const std::vector<float>& a;
    std::string& out;
    void go(size_t dim, size_t begin, size_t count) {
      if (dim == shape.size()) {
        CC_REQUIRE(count == 1, CC_ERR_ILLEGAL_ARGUMENT, "shape does not match the data size");
        out += java_float_to_string(a[begin]);
        return;
      }
      out += "[";
      const size_t head = (size_t)shape[dim];
      const size_t g = head ? count / head : 0;
      for (size_t i = 0; i < head; ++i) {
        if (i) out += ",";
        go(dim + 1, begin + i * g, g);
      }
      out += "]";
    }
  } rec{shape, flat, out};
  rec.go(0, 0, flat.size());
  return out;
}

}  // namespace cuda
}  // namespace compute

// ---- flat C view ---------------------------------------------------------------------------------------------------------------------

using namespace compute::cuda;
using cc::guarded;

namespace {
// every handle given out is registered: a stale or made-up ct_tensor is an IllegalArgument, not a wild pointer dereference
// inside the caller's JVM
std::mutex g_handles_mu;
std::unordered_set<TensorPtr*> g_handles;

TensorPtr& ref(ct_tensor h) {
  CC_REQUIRE(h, CC_ERR_ILLEGAL_ARGUMENT, "null tensor handle");
  TensorPtr* p = (TensorPtr*)(uintptr_t)h;
  std::lock_guard<std::mutex> lock(g_handles_mu);
  CC_REQUIRE(g_handles.count(p), CC_ERR_ILLEGAL_ARGUMENT, "invalid tensor handle");
  return *p;
}
ct_tensor wrap(TensorPtr t) {
  TensorPtr* p = new TensorPtr(std::move(t));
  std::lock_guard<std::mutex> lock(g_handles_mu);
  g_handles.insert(p);
  return (ct_tensor)(uintptr_t)p;
}
Shape to_shape(const int32_t* s, int rank) {
  CC_REQUIRE(rank >= 0 && (rank == 0 || s), CC_ERR_ILLEGAL_ARGUMENT, "bad shape arguments");
  return Shape(s, s + rank);
}
}  // namespace

extern "C" {

int ct_from_host(const float* data, const int32_t* shape, int rank, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(from_host(data, to_shape(shape, rank), padding)); });
}
int ct_from_buffer(cc_buffer buf, const int32_t* shape, int rank, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(from_buffer(buf, to_shape(shape, rank), padding)); });
}
int ct_scalar(float value, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(scalar(value, padding)); });
}
int ct_fill(float value, const int32_t* shape, int rank, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(fill(value, to_shape(shape, rank), padding)); });
}
int ct_random(const int32_t* shape, int rank, int32_t seed, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(compute::cuda::random(to_shape(shape, rank), seed, padding)); });
}
int ct_random_normal(const int32_t* shape, int rank, int32_t seed, float padding, ct_tensor* out) {
  return guarded([&] { *out = wrap(random_normal(to_shape(shape, rank), seed, padding)); });
}
int ct_unary(int op, ct_tensor t, ct_tensor* out) {
  return guarded([&] { *out = wrap(unary((uint32_t)op, ref(t))); });
}
int ct_binary(int op, ct_tensor l, ct_tensor r, ct_tensor* out) {
  return guarded([&] { *out = wrap(binary((uint32_t)op, ref(l), ref(r))); });
}
int ct_broadcast(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->broadcast(to_shape(shape, rank))); });
}
int ct_reshape(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->reshape(to_shape(shape, rank))); });
}
int ct_scale(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->scale(to_shape(shape, rank))); });
}
int ct_translate(ct_tensor t, const double* offset, int n_offset, const int32_t* new_shape, int new_rank, ct_tensor* out) {
  return guarded([&] {
    CC_REQUIRE(n_offset >= 0 && (offset || n_offset == 0), CC_ERR_ILLEGAL_ARGUMENT, "bad offset");
    std::vector<double> off(offset, offset + n_offset);
    if (new_rank < 0)
      *out = wrap(ref(t)->translate(off));
    else
      *out = wrap(ref(t)->translate(off, to_shape(new_shape, new_rank)));
  });
}
int ct_permute(ct_tensor t, const int32_t* dims, int n, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->permute(std::vector<int32_t>(dims, dims + n))); });
}
int ct_transpose(ct_tensor t, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->transpose()); });
}
int ct_split(ct_tensor t, int dimension, ct_tensor* out, int capacity, int* out_count) {
  return guarded([&] {
    auto parts = ref(t)->split(dimension);
    if (out_count) *out_count = (int)parts.size();
    if (!out) return;
    CC_REQUIRE(capacity >= (int)parts.size(), CC_ERR_ILLEGAL_ARGUMENT, "split produces %zu tensors, capacity %d", parts.size(), capacity);
    for (size_t i = 0; i < parts.size(); ++i) out[i] = wrap(std::move(parts[i]));
  });
}
int ct_join(const ct_tensor* tensors, int n, ct_tensor* out) {
  return guarded([&] {
    std::vector<TensorPtr> v;
    for (int i = 0; i < n; ++i) v.push_back(ref(tensors[i]));
    *out = wrap(join(v));
  });
}
int ct_join_dim(const ct_tensor* tensors, int n, int dimension, ct_tensor* out) {
  return guarded([&] {
    std::vector<TensorPtr> v;
    for (int i = 0; i < n; ++i) v.push_back(ref(tensors[i]));
    *out = wrap(join(v, dimension));
  });
}
int ct_sum(ct_tensor t, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->sum()); });
}
int ct_reduce(ct_tensor t, int monoid, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->reduce((uint32_t)monoid)); });
}
int ct_non_inline(ct_tensor t, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->non_inline()); });
}
int ct_do_cache(ct_tensor t, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->do_cache()); });
}
int ct_rank(ct_tensor t, int* out) {
  return guarded([&] { *out = (int)ref(t)->shape.size(); });
}
int ct_shape(ct_tensor t, int32_t* out, int capacity) {
  return guarded([&] {
    const Shape& s = ref(t)->shape;
    CC_REQUIRE(capacity >= (int)s.size(), CC_ERR_ILLEGAL_ARGUMENT, "shape capacity too small");
    for (size_t i = 0; i < s.size(); ++i) out[i] = s[i];
  });
}
int ct_padding(ct_tensor t, float* out) {
  return guarded([&] { *out = ref(t)->padding; });
}
int ct_flat_buffer(ct_tensor t, float** out_host, uint64_t* out_n_floats) {
  return guarded([&] {
    CC_REQUIRE(out_host && out_n_floats, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    const uint64_t n = (uint64_t)ref(t)->size();
    void* host = nullptr;
    int st = cc_host_alloc(n * 4, &host);
    if (st != CC_OK) throw Error(st, cc_last_error());
    try {
      ref(t)->read_into_pinned((float*)host, n);
    } catch (...) {
      cc_host_free(host);
      throw;
    }
    *out_host = (float*)host;
    *out_n_floats = n;
  });
}
int ct_flat_buffer_release(float* host) { return cc_host_free(host); }
int ct_flat_array(ct_tensor t, float* host_out, uint64_t capacity) {
  return guarded([&] { ref(t)->flat_array_into(host_out, capacity); });
}
int ct_to_string(ct_tensor t, char* out, uint64_t capacity, uint64_t* out_needed) {
  return guarded([&] {
    std::string s = ref(t)->to_string();
    if (out_needed) *out_needed = s.size() + 1;
    if (out && capacity) {
      size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
      memcpy(out, s.data(), n);
      out[n] = 0;
    }
  });
}
int ct_do_buffer(ct_tensor t, cc_buffer* out, cc_event* out_event) {
  return guarded([&] {
    Session s;
    PendingBuffer p = ref(t)->do_buffer(s);
    *out = p.buffer;
    if (out_event) *out_event = 0;
  });
}
int ct_compile(ct_tensor t, cc_kernel* out) {
  return guarded([&] { *out = ref(t)->compile_only(); });
}
int ct_tree_blob(ct_tensor t, void* out, uint64_t capacity, uint64_t* out_needed) {
  return guarded([&] {
    std::string blob = ref(t)->tree_blob();
    if (out_needed) *out_needed = blob.size();
    if (out) {
      CC_REQUIRE(capacity >= blob.size(), CC_ERR_ILLEGAL_ARGUMENT, "blob of %zu bytes does not fit in %llu", blob.size(), (unsigned long long)capacity);
      memcpy(out, blob.data(), blob.size());
    }
  });
}
int ct_shard(ct_tensor local, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(local)->as_row_block()); });
}
int ct_distribution(ct_tensor t, int* out) {
  return guarded([&] { *out = (int)ref(t)->dist; });
}
int ct_gather(ct_tensor t, int zero_copy, ct_tensor* out) {
  return guarded([&] { *out = wrap(ref(t)->gather(zero_copy != 0)); });
}
int ct_release(ct_tensor t) {
  return guarded([&] {
    TensorPtr* p = &ref(t);
    {
      std::lock_guard<std::mutex> lock(g_handles_mu);
      g_handles.erase(p);
    }
    delete p;  // outside the lock: may cascade into a whole graph's destruction
  });
}
int ct_live_tensors(int64_t* out) {
  return guarded([&] { *out = live_tensors(); });
}

}  // extern "C"
