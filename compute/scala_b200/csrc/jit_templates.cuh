// jit_templates.cuh — hand-written sm_100a device templates every generated kernel is instantiated from.
// Embedded into libcompute_cuda.so as a string and prepended to the generated source before NVRTC
// (--gpu-architecture=sm_100a). Replaces the scalar, one-work-item-per-element OpenCL C the reference generates
// (OpenCLKernelBuilder.scala:135-221) — no OpenCL construct is translated; these are CUDA-native building blocks:
//   * 128-bit read-only streaming loads that do not allocate in L1, 128-bit streaming stores;
//   * float ops with the accuracy contract of the north star (<= 2 ulp): IEEE add/mul/div/sqrt, fma contraction on
//     (the reference builds with -cl-unsafe-math-optimizations and FP_CONTRACT ON, OpenCL.scala:1131-1135);
//   * warp-shuffle and shared-memory block reductions.

// ---- memory ---------------------------------------------------------------------------------------------------------
// (CC_HOST_EMULATION is defined only by tests/kernel_emulator, which compiles generated kernels for the host to check the code
// generator's arithmetic and indexing without a GPU; NVRTC never sees it. The four PTX wrappers get plain C++ bodies there.)
#ifdef CC_HOST_EMULATION
__device__ __forceinline__ float cc_ldg(const float* p) { return *p; }
__device__ __forceinline__ void cc_ldg4(const float* p, float (&v)[4]) { v[0] = p[0], v[1] = p[1], v[2] = p[2], v[3] = p[3]; }
__device__ __forceinline__ void cc_stg4(float* p, const float (&v)[4]) { p[0] = v[0], p[1] = v[1], p[2] = v[2], p[3] = v[3]; }
__device__ __forceinline__ void cc_split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  const float r = x - hi;
  unsigned u = __float_as_uint(r);
  u = (u + 0x1000u) & 0xffffe000u;  // round to nearest, ties away: cvt.rna.tf32
  lo = __uint_as_float(u);
}
#else
// Loads of kernel arguments: non-coherent (ld.global.nc) — a generated kernel never writes what it reads, and the runtime launches it
// with programmatic dependent launch (an early-resident grid) only when no command that could still be running writes its inputs
// (runtime.cpp: pdl_now), so the data is read-only for the grid's whole lifetime as PTX requires of .nc. CC_COHERENT_LOADS (opt-in through
// the environment variable of the same name) compiles them as ordinary coherent loads with the same cache hints, for A/B comparisons.
#ifdef CC_COHERENT_LOADS
#define CC_LD_NC ""
#else
#define CC_LD_NC ".nc"
#endif
__device__ __forceinline__ float cc_ldg(const float* p) {
  float v;
  asm("ld.global" CC_LD_NC ".L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// p must be 16-byte aligned
__device__ __forceinline__ void cc_ldg4(const float* p, float (&v)[4]) {
  asm("ld.global" CC_LD_NC ".L1::no_allocate.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
      : "l"(p));
}
#endif

// First statement of every generated entry point when the runtime launches with programmatic stream serialisation (CC_PDL=1):
// let the next kernel's CTAs be scheduled, then wait until everything this grid depends on has completed and is visible.
#ifdef CC_HOST_EMULATION
__device__ __forceinline__ void cc_pdl_entry() {}
#else
__device__ __forceinline__ void cc_pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// 16-byte asynchronous global -> shared copies (L1 bypassed), in commit groups: the staging pipeline of dense-window tiles
#ifdef CC_HOST_EMULATION
__device__ __forceinline__ void cc_cp_async16(float* smem_dst, const float* src) { smem_dst[0] = src[0], smem_dst[1] = src[1], smem_dst[2] = src[2], smem_dst[3] = src[3]; }
__device__ __forceinline__ void cc_cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cc_cp_async_wait() {}
#else
__device__ __forceinline__ void cc_cp_async16(float* smem_dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cc_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

// cached flavours (allocate in L1) for data that is reused across the index space: broadcast operands, the operands of a
// re-rolled reduction whose address does not depend on every output index (matmul / convolution patterns)
#if defined(CC_COHERENT_LOADS) && !defined(CC_HOST_EMULATION)
__device__ __forceinline__ float cc_ldc(const float* p) { return *p; }
#else
__device__ __forceinline__ float cc_ldc(const float* p) { return __ldg(p); }
#endif
// p must be 8-byte aligned. A volatile statement: it keeps its place among the other volatile statements of a kernel (the MMAs of the
// small-N contraction), so a batch of these written ahead of the MMAs that consume them is ISSUED ahead of them — left to itself the compiler
// sinks every load to just before its use and a warp has two loads in flight instead of sixteen.
__device__ __forceinline__ float2 cc_ldc2(const float* p) {
#if defined(CC_HOST_EMULATION)
  return make_float2(p[0], p[1]);
#else
  float2 v;
  asm volatile("ld.global" CC_LD_NC ".v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ float cc_ldc1v(const float* p) {
#if defined(CC_HOST_EMULATION)
  return *p;
#else
  float v;
  asm volatile("ld.global" CC_LD_NC ".f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ void cc_ldc4(const float* p, float (&v)[4]) {
#if defined(CC_COHERENT_LOADS) && !defined(CC_HOST_EMULATION)
  const float4 x = *reinterpret_cast<const float4*>(p);
#else
  const float4 x = __ldg(reinterpret_cast<const float4*>(p));
#endif
  v[0] = x.x;
  v[1] = x.y;
  v[2] = x.z;
  v[3] = x.w;
}

#ifndef CC_HOST_EMULATION
__device__ __forceinline__ void cc_stg4(float* p, const float (&v)[4]) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// x = hi + lo with hi exactly representable in TF32 (top 19 bits) and lo = tf32(x - hi): the operand split of the 3xTF32 contraction
__device__ __forceinline__ void cc_split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  const float r = x - hi;
  unsigned t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(r));
  lo = __uint_as_float(t);
}
#endif

// D (16 x 8, fp32) += A (16 x 8, row) * B (8 x 8, col) on the warp-level tensor-core path. Fragments (g = lane / 4, t = lane % 4):
// a = {A[g][t], A[g + 8][t], A[g][t + 4], A[g + 8][t + 4]}, b = {B[t][g], B[t + 4][g]}, c = {C[g][2t], C[g][2t + 1], C[g + 8][2t], C[g + 8][2t + 1]}
#ifdef CC_HOST_EMULATION
__device__ __forceinline__ void cc_mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) { cc_emu_mma_tf32_16x8x8(c, a, b); }
#else
__device__ __forceinline__ void cc_mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// ---- math -------------------------------------------------------------------------------------------------------------
// CUDA's expf / logf / tanhf are documented at <= 2 / 1 / 2 ulp; kept behind cc_* names so that leaner or tighter
// implementations can be swapped in without touching the generator.

__device__ __forceinline__ float cc_exp(float x) { return expf(x); }
__device__ __forceinline__ float cc_log(float x) { return logf(x); }
__device__ __forceinline__ float cc_tanh(float x) { return tanhf(x); }

// ---- reductions -------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float cc_warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// all 256 threads of the block must call; result valid in thread 0
__device__ __forceinline__ float cc_block_sum_256(float v) {
  __shared__ float cc_red_[8];
  v = cc_warp_sum(v);
  if ((threadIdx.x & 31) == 0) cc_red_[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < 8 ? cc_red_[threadIdx.x] : 0.f;
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
  }
  return v;
}

// ---- collectives folded into a reduction's final stage (sharded tensors, one process per GPU) ---------------------------------------
// The mailboxes of builtin_kernels.h (PeerMailboxes, same layout): every rank owns [2][world][65536] LL slots {float, epoch} in its HBM,
// mapped into every peer. The thread that holds a final value pushes it into all ranks' mailboxes (NVLink stores) and then reads every
// rank's slot out of its own mailbox — the payload carries its own ready flag. A generated reduction takes `mb_` = null on ordinary
// launches; cc_shard_launch_allreduce / _allgather pass the mailboxes and a fresh epoch, and the kernel IS the collective (the sum of
// the contributions is taken in rank order from 0.f, as the stand-alone all-reduce kernel takes it: same bits on every rank and route).
struct cc_peer_mailboxes {
  float* data[8];
  unsigned* flags[8];
  int world;
  int rank;
};
#ifdef CC_HOST_EMULATION
__device__ __forceinline__ void cc_ll_allreduce4(float (&)[4], unsigned long long, const cc_peer_mailboxes*, unsigned) {}
__device__ __forceinline__ void cc_ll_allgather1(float, unsigned long long, unsigned long long, float*, const cc_peer_mailboxes*, unsigned) {}
#else
__device__ __forceinline__ void cc_ll_store2(uint2* p, float v0, float v1, unsigned e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(e), "r"(__float_as_uint(v1)), "r"(e) : "memory");
}
__device__ __forceinline__ void cc_ll_store1(uint2* p, float v, unsigned e) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(e) : "memory");
}
__device__ __forceinline__ void cc_ll_load2(const uint2* p, unsigned e, float& v0, float& v1) {
  unsigned a, b, c, d, spins = 0;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (++spins > (1u << 27)) __trap();  // a peer that never shows up must abort this launch, not hang the GPU
  } while (b != e || d != e);
  v0 = __uint_as_float(a);
  v1 = __uint_as_float(c);
}
__device__ __forceinline__ float cc_ll_load1(const uint2* p, unsigned e) {
  unsigned a, b, spins = 0;
  do {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    if (++spins > (1u << 27)) __trap();
  } while (b != e);
  return __uint_as_float(a);
}
__device__ __forceinline__ uint2* cc_ll_slot(const cc_peer_mailboxes* mb, int owner, int parity, int src_rank, unsigned long long i) {
  return reinterpret_cast<uint2*>(mb->data[owner]) + ((size_t)parity * mb->world + src_rank) * (size_t)65536 + i;
}
// acc[0..4) = sum over ranks of their acc[0..4) for elements [i, i + 4), i % 4 == 0, i + 4 <= 65536
__device__ __forceinline__ void cc_ll_allreduce4(float (&acc)[4], unsigned long long i, const cc_peer_mailboxes* mb, unsigned epoch) {
  const int parity = (int)(epoch & 1u), world = mb->world, rank = mb->rank;
  for (int peer = 0; peer < world; ++peer) {
    uint2* dst = cc_ll_slot(mb, peer, parity, rank, i);
    cc_ll_store2(dst, acc[0], acc[1], epoch);
    cc_ll_store2(dst + 2, acc[2], acc[3], epoch);
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int r = 0; r < world; ++r) {
    const uint2* src = cc_ll_slot(mb, rank, parity, r, i);
    float x0, x1, x2, x3;
    cc_ll_load2(src, epoch, x0, x1);
    cc_ll_load2(src + 2, epoch, x2, x3);
    s0 = __fadd_rn(s0, x0), s1 = __fadd_rn(s1, x1), s2 = __fadd_rn(s2, x2), s3 = __fadd_rn(s3, x3);
  }
  acc[0] = s0, acc[1] = s1, acc[2] = s2, acc[3] = s3;
}
// out[r * n + i] = rank r's v for every rank r (this rank's own included)
__device__ __forceinline__ void cc_ll_allgather1(float v, unsigned long long i, unsigned long long n, float* out, const cc_peer_mailboxes* mb, unsigned epoch) {
  const int parity = (int)(epoch & 1u), world = mb->world, rank = mb->rank;
  for (int peer = 0; peer < world; ++peer) cc_ll_store1(cc_ll_slot(mb, peer, parity, rank, i), v, epoch);
  for (int r = 0; r < world; ++r) out[(size_t)r * n + i] = cc_ll_load1(cc_ll_slot(mb, rank, parity, r, i), epoch);
}
#endif

// ---- monoids (MonoidPrograms, Tensors.scala:308-311: append / zero) ------------------------------------------------------
// ap() never contracts with the producer of its operands (__fadd_rn / __fmul_rn), so a reduction fused with an elementwise
// closure folds exactly the values the unfused "materialise, then reduce" sequence of the reference would fold.

struct cc_plus { static __device__ __forceinline__ float zero() { return 0.f; } static __device__ __forceinline__ float ap(float a, float b) { return __fadd_rn(a, b); } };
struct cc_times { static __device__ __forceinline__ float zero() { return 1.f; } static __device__ __forceinline__ float ap(float a, float b) { return __fmul_rn(a, b); } };
struct cc_min { static __device__ __forceinline__ float zero() { return __int_as_float(0x7f800000); } static __device__ __forceinline__ float ap(float a, float b) { return fminf(a, b); } };
struct cc_max { static __device__ __forceinline__ float zero() { return __int_as_float(0xff800000); } static __device__ __forceinline__ float ap(float a, float b) { return fmaxf(a, b); } };

// min / max as re-rolled chains fold them (`t.split(axis).reduce(Tensor.max)`): NaN is the neutral element of fminf / fmaxf
struct cc_min_nan { static __device__ __forceinline__ float zero() { return __int_as_float(0x7fc00000); } static __device__ __forceinline__ float ap(float a, float b) { return fminf(a, b); } };
struct cc_max_nan { static __device__ __forceinline__ float zero() { return __int_as_float(0x7fc00000); } static __device__ __forceinline__ float ap(float a, float b) { return fmaxf(a, b); } };

template <class M>
__device__ __forceinline__ float cc_warp_fold(float v) {
  v = M::ap(v, __shfl_xor_sync(0xffffffffu, v, 16));
  v = M::ap(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = M::ap(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = M::ap(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = M::ap(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

// every thread of the (multiple-of-32, <= 1024 thread) block must call; result valid in warp 0
template <class M>
__device__ __forceinline__ float cc_block_fold(float v, float* smem /* >= 32 floats */) {
  v = cc_warp_fold<M>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? smem[lane] : M::zero();
    v = cc_warp_fold<M>(v);
  }
  return v;
}

// Second stage of a whole-tensor fold: block `blockIdx.x` publishes its partial; the last block to arrive folds all partials
// in index order (deterministic: no float atomics) and resets the counter for the next launch on this stream.
template <class M>
__device__ __forceinline__ void cc_fold_finish(float block_value /* valid in thread 0 */, float* __restrict__ out, float* __restrict__ partials,
                                               unsigned* __restrict__ counter, float* smem) {
  __shared__ bool cc_is_last_;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_value;
    __threadfence();
    const unsigned done = atomicAdd(counter, 1u);
    cc_is_last_ = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!cc_is_last_) return;
  __threadfence();
  float p = M::zero();
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) p = M::ap(p, __ldcg(partials + i));
  p = cc_block_fold<M>(p, smem);
  if (threadIdx.x == 0) {
    out[0] = p;
    *counter = 0u;
  }
}
