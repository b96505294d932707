// kernels_basic.cu — precompiled sm_100a kernels for the hand-written programs of Tensors.scala:
//   full-sum reduction   (replaces sequentialReductionProgram T:313-351 / parallelReductionProgram T:358-392)
//   random               (T:432-443, Wang hash T:106-117)
//   random_normal        (T:398-429)
// Designed for B200: 128-bit streaming loads with 4 independent vectors in flight per thread, grid sized to the SM
// count, warp-shuffle + shared-memory block reduction, deterministic last-block-done second stage (no float atomics).
#include <cstdint>

#include "builtin_kernels.h"
#include "common.h"

namespace cc {
namespace {

constexpr int kReduceThreads = 512;
constexpr int kReduceMaxBlocks = 148 * 4 * 2;  // upper bound on partials for any sm_count <= 296

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? smem[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;  // valid in warp 0
}

// ---- peer mailbox helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch) {
  unsigned spins = 0;
  while (ld_acquire_sys(p) != epoch)
    if (++spins > (1u << 27)) __trap();  // a peer that never shows up must abort this launch, not hang the GPU
}
// LL slots (the "low latency" protocol NCCL uses for small messages): every float travels as an 8-byte {value, epoch} pair, so the
// payload carries its own ready flag -- no fence, no separate flag store, no extra NVLink round trip between data and flag. Two
// slots are written / polled per 16-byte access; each 8-byte half is a single transaction, so a half is either old or new.
__device__ __forceinline__ void ll_store2(uint2* p, float v0, float v1, unsigned e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(e), "r"(__float_as_uint(v1)), "r"(e) : "memory");
}
__device__ __forceinline__ void ll_store1(uint2* p, float v, unsigned e) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(e) : "memory");
}
__device__ __forceinline__ void ll_load2(const uint2* p, unsigned e, float& v0, float& v1) {
  unsigned a, b, c, d, spins = 0;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (++spins > (1u << 27)) __trap();  // a peer that never shows up must abort this launch, not hang the GPU
  } while (b != e || d != e);
  v0 = __uint_as_float(a);
  v1 = __uint_as_float(c);
}
__device__ __forceinline__ float ll_load1(const uint2* p, unsigned e) {
  unsigned a, b, spins = 0;
  do {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    if (++spins > (1u << 27)) __trap();
  } while (b != e);
  return __uint_as_float(a);
}
__device__ __forceinline__ uint2* mb_slots(const PeerMailboxes& mb, int rank) { return reinterpret_cast<uint2*>(mb.data[rank]); }

__device__ __forceinline__ size_t mb_data_index(int parity, int world, int src_rank, size_t i) {
  return ((size_t)parity * world + src_rank) * kPeerCapFloats + i;
}
__device__ __forceinline__ size_t mb_flag_index(int parity, int world, int src_rank, int block) {
  return ((size_t)parity * world + src_rank) * kPeerMaxBlocks + block;
}

template <bool kFused>
__global__ void __launch_bounds__(kReduceThreads) reduce_sum_kernel(const float* __restrict__ in, uint64_t n, float* __restrict__ out,
                                                                   float* __restrict__ partials, unsigned* __restrict__ counter, PeerMailboxes mb,
                                                                   unsigned epoch) {
  __shared__ float smem[32];
  __shared__ bool is_last;
  const uint64_t nvec = n >> 2;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  const uint64_t stride = (uint64_t)gridDim.x * kReduceThreads;
  uint64_t v = (uint64_t)blockIdx.x * kReduceThreads + threadIdx.x;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  // 4 independent 128-bit loads in flight per thread per iteration
  for (; v + 3 * stride < nvec; v += 4 * stride) {
    float4 x0 = ldg_stream4(in4 + v);
    float4 x1 = ldg_stream4(in4 + v + stride);
    float4 x2 = ldg_stream4(in4 + v + 2 * stride);
    float4 x3 = ldg_stream4(in4 + v + 3 * stride);
    a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
    a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
    a2.x += x2.x; a2.y += x2.y; a2.z += x2.z; a2.w += x2.w;
    a3.x += x3.x; a3.y += x3.y; a3.z += x3.z; a3.w += x3.w;
  }
  for (; v < nvec; v += stride) {
    float4 x0 = ldg_stream4(in4 + v);
    a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
  }
  float acc = ((a0.x + a1.x) + (a2.x + a3.x)) + ((a0.y + a1.y) + (a2.y + a3.y)) + (((a0.z + a1.z) + (a2.z + a3.z)) + ((a0.w + a1.w) + (a2.w + a3.w)));
  // scalar tail (n % 4 elements) — one designated thread
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (uint64_t i = nvec << 2; i < n; ++i) acc += in[i];
  acc = block_sum(acc, smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    const unsigned done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // second stage: fixed order over the (<= kReduceMaxBlocks) partials
  float p = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kReduceThreads) p += __ldcg(partials + i);
  p = block_sum(p, smem);
  if (!kFused) {
    if (threadIdx.x == 0) {
      out[0] = p;
      *counter = 0u;  // ready for the next launch on this stream
    }
    return;
  }
  // fused all-reduce: push this GPU's total as an LL slot {value, epoch} into slot [rank] of every peer's mailbox over NVLink, then
  // poll our own mailbox for every rank's slot and add the totals in rank order (same order, same bits on every rank)
  const int parity = (int)(epoch & 1u);
  if (threadIdx.x == 0) smem[0] = p;
  __syncthreads();
  if ((int)threadIdx.x < mb.world) ll_store1(mb_slots(mb, threadIdx.x) + mb_data_index(parity, mb.world, mb.rank, 0), smem[0], epoch);
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int r = 0; r < mb.world; ++r) total += ll_load1(mb_slots(mb, mb.rank) + mb_data_index(parity, mb.world, r, 0), epoch);
    out[0] = total;
    *counter = 0u;
  }
}

// one-shot in-place all-reduce of a vector: block b owns floats [b*1024, (b+1)*1024); every thread pushes its 4 floats as LL slots
// into every peer's mailbox and polls the same 4 slots of every rank in its own -- threads never synchronise with each other
__global__ void __launch_bounds__(256) peer_allreduce_kernel(float* __restrict__ v, uint64_t n, PeerMailboxes mb, unsigned epoch) {
  const int parity = (int)(epoch & 1u);
  const size_t i = (size_t)blockIdx.x * kPeerChunk + (size_t)threadIdx.x * 4;
  if (i >= n) return;
  const bool vec = i + 3 < n;
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  if (vec) {
    const float4 x = *reinterpret_cast<const float4*>(v + i);
    t[0] = x.x, t[1] = x.y, t[2] = x.z, t[3] = x.w;
  } else {
    for (int j = 0; j < 4; ++j)
      if (i + j < n) t[j] = v[i + j];
  }
  for (int peer = 0; peer < mb.world; ++peer) {  // NVLink stores (local for peer == rank)
    uint2* dst = mb_slots(mb, peer) + mb_data_index(parity, mb.world, mb.rank, i);
    ll_store2(dst, t[0], t[1], epoch);
    ll_store2(dst + 2, t[2], t[3], epoch);
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < mb.world; ++r) {
    const uint2* src = mb_slots(mb, mb.rank) + mb_data_index(parity, mb.world, r, i);
    float x0, x1, x2, x3;
    ll_load2(src, epoch, x0, x1);
    ll_load2(src + 2, epoch, x2, x3);
    acc[0] += x0, acc[1] += x1, acc[2] += x2, acc[3] += x3;
  }
  if (vec) {
    *reinterpret_cast<float4*>(v + i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
    for (int j = 0; j < 4; ++j)
      if (i + j < n) v[i + j] = acc[j];
  }
}

// one-shot all-gather: recv[r * n .. (r + 1) * n) = rank r's send[0 .. n). Same LL exchange as the all-reduce above with the
// rank-ordered sum replaced by a copy of every rank's slots into place.
__global__ void __launch_bounds__(256) peer_allgather_kernel(const float* __restrict__ send, float* __restrict__ recv, uint64_t n, PeerMailboxes mb,
                                                             unsigned epoch) {
  const int parity = (int)(epoch & 1u);
  const size_t i = (size_t)blockIdx.x * kPeerChunk + (size_t)threadIdx.x * 4;
  if (i >= n) return;
  const bool vec = (n & 3u) == 0 && i + 3 < n;  // n % 4 == 0 keeps every rank's slice of recv 16-byte aligned
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  if (vec) {
    const float4 x = *reinterpret_cast<const float4*>(send + i);
    t[0] = x.x, t[1] = x.y, t[2] = x.z, t[3] = x.w;
  } else {
    for (int j = 0; j < 4; ++j)
      if (i + j < n) t[j] = send[i + j];
  }
  for (int peer = 0; peer < mb.world; ++peer) {
    uint2* dst = mb_slots(mb, peer) + mb_data_index(parity, mb.world, mb.rank, i);
    ll_store2(dst, t[0], t[1], epoch);
    ll_store2(dst + 2, t[2], t[3], epoch);
  }
  for (int r = 0; r < mb.world; ++r) {
    const uint2* src = mb_slots(mb, mb.rank) + mb_data_index(parity, mb.world, r, i);
    float y[4];
    ll_load2(src, epoch, y[0], y[1]);
    ll_load2(src + 2, epoch, y[2], y[3]);
    float* dst = recv + (size_t)r * n + i;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (i + j < n) dst[j] = y[j];
    }
  }
}

// every rank raises its epoch flag in every peer's mailbox (after a system-scope fence, so everything this GPU wrote before —
// including the previous kernel's stores into peer memory — is visible first) and waits for all peers' flags in its own
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerMailboxes mb, unsigned epoch) {
  const int parity = (int)(epoch & 1u);
  if ((int)threadIdx.x < mb.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(mb.flags[peer] + mb_flag_index(parity, mb.world, mb.rank, 0), epoch);
    wait_flag(mb.flags[mb.rank] + mb_flag_index(parity, mb.world, peer, 0), epoch);
  }
}

__device__ __forceinline__ uint32_t wang_hash(uint32_t value) {  // T:106-117
  value = (value ^ 61u) ^ (value >> 16);
  value *= 9u;
  value ^= value << 4;
  value *= 0x27d4eb2du;
  value ^= value >> 15;
  return value;
}
__device__ __forceinline__ uint32_t xorshift(uint32_t seed) {  // T:403-407
  const uint32_t tmp1 = seed ^ (seed << 13);
  const uint32_t tmp2 = tmp1 ^ (tmp1 >> 17);
  return tmp2 ^ (tmp2 << 5);
}
__device__ __forceinline__ float u32_to_unit(uint32_t h) { return __uint2float_rn(h) * 2.3283064365386963e-10f; }  // h / 2^32

__global__ void __launch_bounds__(256) random_kernel(float* __restrict__ out, uint64_t n, uint32_t seed) {
  const uint64_t nvec = n >> 2;
  const uint64_t stride = (uint64_t)gridDim.x * 256;
  for (uint64_t v = (uint64_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += stride) {
    const uint32_t i = (uint32_t)(v << 2);
    float4 r;
    r.x = u32_to_unit(wang_hash((i + 0u) ^ seed));
    r.y = u32_to_unit(wang_hash((i + 1u) ^ seed));
    r.z = u32_to_unit(wang_hash((i + 2u) ^ seed));
    r.w = u32_to_unit(wang_hash((i + 3u) ^ seed));
    __stcs(reinterpret_cast<float4*>(out) + v, r);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const uint64_t i = (nvec << 2) + threadIdx.x;
    out[i] = u32_to_unit(wang_hash((uint32_t)i ^ seed));
  }
}

__global__ void __launch_bounds__(256) random_normal_kernel(float* __restrict__ out, uint64_t n, uint32_t seed) {
  const uint64_t npair = (n + 1) >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * 256;
  for (uint64_t p = (uint64_t)blockIdx.x * 256 + threadIdx.x; p < npair; p += stride) {
    const uint32_t i = (uint32_t)p;
    const uint32_t r1 = wang_hash(i ^ seed);
    const uint32_t r2 = xorshift(r1);
    const float u1 = u32_to_unit(r1);
    const float u2 = u32_to_unit(r2);
    const float r = sqrtf(-2.f * logf(u1));
    const float theta = (2.f * 3.14159274101257f) * u2;  // 2 * M_PI_F
    float s, c;
    sincosf(theta, &s, &c);
    const float z0 = r * c;
    const float z1 = r * s;
    if (2 * p + 1 < n) {
      __stcs(reinterpret_cast<float2*>(out) + p, make_float2(z0, z1));
    } else {
      out[2 * p] = z0;
    }
  }
}

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("%s launch failed: %s", what, cudaGetErrorString(e)));
}

}  // namespace

uint64_t reduce_sum_scratch_floats() { return kReduceMaxBlocks; }

void launch_reduce_sum(const float* in, uint64_t n, float* out, float* scratch, unsigned* counter, int sm_count, cudaStream_t stream) {
  const uint64_t nvec = n >> 2;
  uint64_t want = (nvec + (uint64_t)kReduceThreads * 4 - 1) / ((uint64_t)kReduceThreads * 4);
  uint64_t cap = (uint64_t)sm_count * 4;
  if (cap > (uint64_t)kReduceMaxBlocks) cap = kReduceMaxBlocks;
  unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
  reduce_sum_kernel<false><<<grid, kReduceThreads, 0, stream>>>(in, n, out, scratch, counter, PeerMailboxes{}, 0u);
  check_launch("reduce_sum");
}

size_t peer_mailbox_flag_offset(int world) { return (size_t)2 * world * kPeerCapFloats * sizeof(uint2); }  // LL slots: 8 bytes per float
size_t peer_mailbox_bytes(int world) { return peer_mailbox_flag_offset(world) + (size_t)2 * world * kPeerMaxBlocks * sizeof(unsigned); }

void launch_peer_allreduce(float* v, uint64_t n, const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream) {
  CC_REQUIRE(n <= (uint64_t)kPeerCapFloats, CC_ERR_UNSUPPORTED, "peer all-reduce carries at most %d floats", kPeerCapFloats);
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + kPeerChunk - 1) / kPeerChunk);
  peer_allreduce_kernel<<<blocks, 256, 0, stream>>>(v, n, mb, epoch);
  check_launch("peer_allreduce");
}

void launch_peer_allgather(const float* send, float* recv, uint64_t n_per_rank, const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream) {
  CC_REQUIRE(n_per_rank <= (uint64_t)kPeerCapFloats, CC_ERR_UNSUPPORTED, "peer all-gather carries at most %d floats per rank", kPeerCapFloats);
  if (n_per_rank == 0) return;
  const unsigned blocks = (unsigned)((n_per_rank + kPeerChunk - 1) / kPeerChunk);
  peer_allgather_kernel<<<blocks, 256, 0, stream>>>(send, recv, n_per_rank, mb, epoch);
  check_launch("peer_allgather");
}

void launch_peer_barrier(const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream) {
  peer_barrier_kernel<<<1, 32, 0, stream>>>(mb, epoch);
  check_launch("peer_barrier");
}

void launch_reduce_sum_allreduce(const float* in, uint64_t n, float* out, float* scratch, unsigned* counter, int sm_count,
                                 const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream) {
  const uint64_t nvec = n >> 2;
  uint64_t want = (nvec + (uint64_t)kReduceThreads * 4 - 1) / ((uint64_t)kReduceThreads * 4);
  uint64_t cap = (uint64_t)sm_count * 4;
  if (cap > (uint64_t)kReduceMaxBlocks) cap = kReduceMaxBlocks;
  unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
  reduce_sum_kernel<true><<<grid, kReduceThreads, 0, stream>>>(in, n, out, scratch, counter, mb, epoch);
  check_launch("reduce_sum_allreduce");
}

void launch_random(float* out, uint64_t n, int32_t seed, cudaStream_t stream) {
  uint64_t blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  random_kernel<<<(unsigned)blocks, 256, 0, stream>>>(out, n, (uint32_t)seed);
  check_launch("random");
}

void launch_random_normal(float* out, uint64_t n, int32_t seed, cudaStream_t stream) {
  uint64_t blocks = (((n + 1) >> 1) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  random_normal_kernel<<<(unsigned)blocks, 256, 0, stream>>>(out, n, (uint32_t)seed);
  check_launch("random_normal");
}

}  // namespace cc
