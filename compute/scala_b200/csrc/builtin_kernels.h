// builtin_kernels.h — host launchers of the precompiled (nvcc, sm_100a) kernels: the hand-written programs that in the
// reference are OpenCL C strings inside Tensors.scala (reduction T:313-392, random T:432-443, randomNormal T:398-429),
// plus the tcgen05 contraction the matmul pattern lowers to.
#pragma once
#include <cuda.h>
#include <cuda_runtime_api.h>

#include <cstdint>

namespace cc {

// out[0] = sum(in[0..n)). scratch: >= reduce_sum_scratch_floats() floats; counter: one zero-initialised u32 that the
// kernel resets itself. Deterministic (fixed grid, last-block-done second stage, no float atomics).
uint64_t reduce_sum_scratch_floats();
void launch_reduce_sum(const float* in, uint64_t n, float* out, float* scratch, unsigned* counter, int sm_count,
                       cudaStream_t stream);

void launch_random(float* out, uint64_t n, int32_t seed, cudaStream_t stream);
void launch_random_normal(float* out, uint64_t n, int32_t seed, cudaStream_t stream);

// ---- one-shot all-reduce over NVLink peer memory (one process per GPU, CUDA IPC mailboxes) ---------------------------------
// Every rank owns a mailbox in its own HBM; peers store their contribution straight into it over NVLink / NVSwitch and
// raise a per-block epoch flag; each rank then sums the contributions in rank order (deterministic, identical everywhere).
constexpr int kPeerMaxRanks = 8;
constexpr int kPeerCapFloats = 65536;  // largest vector the mailbox carries (the 16384-float column sums of C3 use a quarter)
constexpr int kPeerChunk = 1024;       // floats per block
constexpr int kPeerMaxBlocks = kPeerCapFloats / kPeerChunk;
struct PeerMailboxes {
  float* data[kPeerMaxRanks];      // data[r]: rank r's mailbox payload: LL slots {float value, u32 epoch} [2][world][kPeerCapFloats]
  unsigned* flags[kPeerMaxRanks];  // flags[r]: rank r's mailbox flags  [2][world][kPeerMaxBlocks]
  int world;
  int rank;
};
size_t peer_mailbox_bytes(int world);
size_t peer_mailbox_flag_offset(int world);  // byte offset of the flags inside one mailbox allocation
// in-place all-reduce(sum) of v[0..n), n <= kPeerCapFloats; `epoch` must increase by one per collective call on every rank
void launch_peer_allreduce(float* v, uint64_t n, const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream);
// one-shot all-gather of <= kPeerCapFloats floats per rank over the same mailboxes: recv[r*n .. (r+1)*n) = rank r's send
void launch_peer_allgather(const float* send, float* recv, uint64_t n_per_rank, const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream);
// cross-rank barrier over the mailbox flags (one tiny kernel): returns on the stream once every rank has reached it
void launch_peer_barrier(const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream);
// Tensor.sum over a sharded tensor as ONE kernel: local two-stage reduction whose last block pushes the partial into every
// peer's mailbox and completes the all-reduce itself
void launch_reduce_sum_allreduce(const float* in, uint64_t n, float* out, float* scratch, unsigned* counter, int sm_count,
                                 const PeerMailboxes& mb, unsigned epoch, cudaStream_t stream);

bool gemm_available();

// ---- contraction ----------------------------------------------------------------------------------------------------
struct GemmWorkspace {
  float* a_hi;   // [M,Kp]  tf32_trunc(A), columns >= K zero
  float* a_lo;   // [M,Kp]  tf32(A - a_hi)
  float* bt_hi;  // [N,Kp]  tf32_trunc(B)^T
  float* bt_lo;  // [N,Kp]  tf32(B - b_hi)^T
  float* k_split_partials = nullptr;  // [gemm_k_splits, M, N] when the launcher may split K over CTA pairs (small products), else null
};
// K rounded up to the pipeline's K step (32 floats): the row length of the four workspace panels
int64_t gemm_padded_k(int64_t k);
// tile configuration the launcher picks for an M x N problem on `sm_count` SMs: 512 = 256 x 256 tiles on CTA pairs
// (tcgen05 cta_group::2), 256 / 128 / 64 = 128 x that many columns on single CTAs
int gemm_pick_config(int64_t m, int64_t n, int sm_count, bool allow_pair);
int gemm_pick_bn(int64_t m, int64_t n, int sm_count);
// The final choice for launch_gemm_3xtf32(a, b, ...): as above, or 1024 = 256 x 256 tiles on CTA pairs with A read as the original fp32
// matrix and split inside the kernel through tensor memory (no A panels: ws.a_hi / ws.a_lo may be null). CC_GEMM_FORCE_CONFIG and
// CC_GEMM_TMEM_A=0 are honoured here.
int gemm_config_for(const float* a, int64_t m, int64_t n, int64_t k, int sm_count, bool gather_epilogue);
// config 1024 only: how many ways K is split so that a product of few tiles still fills the CTA pairs (1 = not split); the caller then
// provides ws.k_split_partials with room for that many M x N partial results, which a second kernel adds in a fixed order
int gemm_k_splits(int64_t m, int64_t n, int64_t k, int sm_count);
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// C[M,N] = A[M,K] * B[K,N], fp32 row-major, 3xTF32 on tcgen05, any M, N, K >= 1 (ragged edges: TMA zero fill in, predicated
// stores out; K is zero-padded inside the workspace). `b_panels_ready`: ws.bt_hi / ws.bt_lo already hold the split of this B
// (the runtime keeps them while B is unchanged), so only A is split. Returns the number of device kernels launched; throws.
int launch_gemm_3xtf32(const float* a, const float* b, float* c, int64_t m, int64_t n, int64_t k,
                       const GemmWorkspace& ws, int sm_count, TensorMapEncodeFn encode, cudaStream_t stream,
                       bool b_panels_ready = false);
// Only the tensor-core pipeline: ws already holds all four K-major hi / lo panels ([M,Kp] and [N,Kp], Kp = gemm_padded_k(k)),
// written by generated gather kernels (general contraction: convolution as an implicit GEMM).
int launch_gemm_3xtf32_panels(float* c, int64_t m, int64_t n, int64_t k, const GemmWorkspace& ws, int sm_count,
                              TensorMapEncodeFn encode, cudaStream_t stream);
// The row-sharded matmul fused with the all-gather of its result: this rank computes C[rank*m_shard .. +m_shard, :] = A_shard * B
// and the epilogue TMA-stores every 32x32 block of it straight into gathered_c[d] (rank d's [world*m_shard, N] result; own HBM for
// d == rank, peer HBM over NVLink otherwise), so the transfer overlaps the MMAs tile by tile. Needs N % 4 == 0. The caller runs a
// cross-rank barrier (launch_peer_barrier) afterwards: when it completes every block of every rank has landed.
int launch_gemm_3xtf32_allgather(const float* a, const float* b, float* const* gathered_c, int world, int rank, int64_t m_shard,
                                 int64_t n, int64_t k, const GemmWorkspace& ws, int sm_count, TensorMapEncodeFn encode,
                                 cudaStream_t stream, bool b_panels_ready = false, float* multicast_c = nullptr);
// (multicast_c: the NVLS multicast mapping of the gathered C, if the symmetric buffer has one — then every block is stored ONCE through it
// and gathered_c is not used)

}  // namespace cc
