// TEST INFRASTRUCTURE ONLY — never compiled into, imported by or shipped with the product (compute/scala_b200).
//
// A minimal CUDA execution model on host threads, just enough to compile the kernels the code generator emits (the generated text
// plus jit_templates.cuh with CC_HOST_EMULATION) with g++ and run them on small grids, so that the generator's indexing, bounds
// tests, padding, lane picks, shared-memory tiles, shuffles and fold orders can be checked against the CPU oracle where no GPU
// exists. One block runs at a time on blockDim host threads (a reusable barrier is __syncthreads); `__shared__` is a static.
// Nothing here is a fallback: the product path stays CUDA-only and fails loudly without a device.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3_ {
  unsigned x = 0, y = 0, z = 0;
};
struct float4 {
  float x, y, z, w;
};
struct uint2 {
  unsigned x, y;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct float2 {
  float x, y;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint4 {
  unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

inline thread_local uint3_ threadIdx, blockIdx;
inline uint3_ blockDim, gridDim;

namespace emu {
inline std::unique_ptr<std::barrier<>> block_barrier;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barriers;
inline float shfl_buf[1024];
inline int linear_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }

// Runs `kernel` for every block of the grid, one block at a time, on blockDim host threads.
inline void run_grid(const unsigned grid[3], const unsigned block[3], const std::function<void()>& kernel) {
  blockDim = uint3_{block[0], block[1], block[2]};
  gridDim = uint3_{grid[0], grid[1], grid[2]};
  const int nthreads = (int)(block[0] * block[1] * block[2]);
  block_barrier = std::make_unique<std::barrier<>>(nthreads);
  warp_barriers.clear();
  for (int w = 0; w < (nthreads + 31) / 32; ++w) warp_barriers.push_back(std::make_unique<std::barrier<>>(std::min(32, nthreads - 32 * w)));
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t] {
      threadIdx = uint3_{(unsigned)(t % (int)block[0]), (unsigned)((t / (int)block[0]) % (int)block[1]), (unsigned)(t / (int)(block[0] * block[1]))};
      for (unsigned bz = 0; bz < grid[2]; ++bz)
        for (unsigned by = 0; by < grid[1]; ++by)
          for (unsigned bx = 0; bx < grid[0]; ++bx) {
            blockIdx = uint3_{bx, by, bz};
            kernel();
            block_barrier->arrive_and_wait();  // the next block reuses the `__shared__` statics
          }
    });
  for (auto& th : pool) th.join();
}
}  // namespace emu

inline void __syncthreads() { emu::block_barrier->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const int t = emu::linear_tid();
  emu::shfl_buf[t] = v;
  emu::warp_barriers[(size_t)(t / 32)]->arrive_and_wait();
  const float r = emu::shfl_buf[t ^ lane_mask];
  emu::warp_barriers[(size_t)(t / 32)]->arrive_and_wait();
  return r;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// mma.sync.m16n8k8 (row x col, fp32 accumulate) for the 32 host threads of a warp: every lane publishes its fragments, then computes its own
// four accumulator elements from the whole tile. Fragment layout as in jit_templates.cuh (g = lane / 4, t = lane % 4):
// a = {A[g][t], A[g+8][t], A[g][t+4], A[g+8][t+4]}, b = {B[t][g], B[t+4][g]}, c = {C[g][2t], C[g][2t+1], C[g+8][2t], C[g+8][2t+1]}.
// Products in double, accumulated in k order and rounded once per call: exact on the exactly representable data the emulator tests use.
namespace emu {
inline unsigned mma_buf[1024][6];
}
inline void cc_emu_mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  const int t_ = emu::linear_tid(), warp0 = t_ / 32 * 32, lane = t_ % 32, g = lane / 4, t = lane % 4;
  for (int i = 0; i < 4; ++i) emu::mma_buf[t_][i] = a[i];
  for (int i = 0; i < 2; ++i) emu::mma_buf[t_][4 + i] = b[i];
  emu::warp_barriers[(size_t)(t_ / 32)]->arrive_and_wait();
  auto A = [&](int row, int k) {  // row in [0, 16), k in [0, 8)
    unsigned u = emu::mma_buf[warp0 + (row % 8) * 4 + (k % 4)][(row >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0)];
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
  };
  auto B = [&](int k, int n) {  // k in [0, 8), n in [0, 8)
    unsigned u = emu::mma_buf[warp0 + n * 4 + (k % 4)][4 + (k >= 4 ? 1 : 0)];
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
  };
  for (int i = 0; i < 4; ++i) {
    const int row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
    double acc = (double)c[i];
    for (int k = 0; k < 8; ++k) acc += A(row, k) * B(k, col);
    c[i] = (float)acc;
  }
  emu::warp_barriers[(size_t)(t_ / 32)]->arrive_and_wait();
}

inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline float __int_as_float(int i) { return __uint_as_float((unsigned)i); }
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
// the generator relies on these being plain IEEE operations that never contract with their producers
inline float __fadd_rn(float a, float b) {
  volatile float r = a + b;
  return r;
}
inline float __fmul_rn(float a, float b) {
  volatile float r = a * b;
  return r;
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldcv(const T* p) { return *p; }
inline void __stcs(float* p, float v) { *p = v; }
inline void __stcs(float2* p, float2 v) { *p = v; }
using std::min;
using std::max;
