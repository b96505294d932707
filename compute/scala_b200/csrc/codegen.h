// codegen.h — Trees graph -> execution plan (pattern matching + CUDA C++ generated from the sm_100a templates).
// Replaces OpenCLKernelBuilder.generateKernelSourceCode (OpenCLKernelBuilder.scala:135-221) and the per-node
// emitters (:251-326, 348-455, 516-571, 632-652).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ir.h"

namespace cc {

enum PlanKind { PLAN_ELEMENTWISE = 0, PLAN_AXIS_REDUCE = 1, PLAN_CONTRACTION = 2, PLAN_TILED_TRANSPOSE = 3, PLAN_FULL_REDUCE = 4 };

enum {
  ARG_OUT = -1,
  ARG_SCRATCH0 = -2 /* -2-k = scratch k */,
  // PLAN_FULL_REDUCE: the runtime's shared partials buffer and its self-resetting block counter (serialised on stream 0)
  ARG_REDUCE_PARTIALS = -100,
  ARG_REDUCE_COUNTER = -101,
  // fused second stage of a split axis reduction (opt-in): self-resetting per-stream block counters, one per blockIdx.x (<= kColCounters)
  ARG_COL_COUNTERS = -102,
  // a reduction that can complete a collective itself (Plan::collective): pointer to the device copy of the peer mailboxes (null on ordinary
  // launches) and the collective's epoch as a 64-bit value
  ARG_PEER_MB = -103,
  ARG_PEER_EPOCH = -104
};
constexpr int kColCounters = 4096;
constexpr int kFullReduceThreads = 512;
constexpr int kFullReduceMaxBlocks = 148 * 4 * 2;  // = kReduceMaxBlocks of kernels_basic.cu: the partials buffer holds this many floats

struct LaunchSpec {
  std::string entry;  // __global__ name inside the generated module
  uint32_t grid[3] = {1, 1, 1};
  uint32_t block[3] = {1, 1, 1};
  uint32_t smem = 0;
  std::vector<int> args;  // >= 0: plan argument i; ARG_OUT; ARG_SCRATCH0 - k
};

struct Plan {
  int kind = PLAN_ELEMENTWISE;
  std::string source;                  // generated CUDA C++ (without the template prelude)
  std::vector<LaunchSpec> launches;    // empty for a plain PLAN_CONTRACTION (runs the precompiled tcgen05 pipeline)
  std::vector<uint32_t> arg_params;    // tree parameter ordinal of each buffer argument
  std::vector<uint64_t> arg_min_floats;  // minimum length each argument buffer must have
  std::vector<uint64_t> scratch_floats;
  uint64_t out_floats = 0;
  uint64_t algorithmic_bytes = 0;
  uint64_t flops = 0;
  int64_t M = 0, N = 0, K = 0;  // contraction
  // contraction whose K-major hi / lo operand panels are written by generated kernels (launches[0] = panel_a, [1] = panel_b,
  // optional [2] = post_kernel applied in place to the result) instead of the precompiled split kernels
  bool gathered_panels = false;
  // 1: the plan's last launch can all-reduce (sum) its output over the peer mailboxes itself; 2: it can all-gather it (the output buffer is then
  // the gathered one, [ranks x out_floats]). 0: neither.
  int collective = 0;
  int64_t batch = 1;  // gathered panels only: leading output dims both operands depend on = that many independent M x N x K products
  std::string note;             // human-readable description of the choices made (kept in the kernel source header)
};

struct DeviceProps {
  int sm_count = 148;
  int max_smem = 227 * 1024;
  bool contraction = true;  // lower sum_t A[i,t]*B[t,k] to the tcgen05 pipeline
};

Plan make_plan(const Tree& t, const DeviceProps& dev);

// The planning switches (CC_TUNE_*, CC_NO_*, CC_FUSE_COL_STAGE, CC_BATCHED_CONTRACTION, CC_DISABLE_CONTRACTION) as sampled from the
// environment on first use and at every plan_knobs_refresh() (called when the kernel cache is cleared or the runtime initialised);
// nullptr = unset. make_plan never reads the environment itself, so plans and the structural cache cannot disagree.
const char* plan_knob(const char* name);
void plan_knobs_refresh();

// K:14-32 — `new DecimalFormat()` (<= 3 fraction digits, HALF_EVEN) as applied to every affine coefficient before it
// is pasted into the kernel text; returns the value the generated code actually uses.
double java_decimal_round(double v);

}  // namespace cc
