// gemm_3xtf32.cu — the contraction the matmul pattern (split / broadcast / sum, benchmarks.scala:174-193,
// README.md:312-343) lowers to:  C[M,N] = A[M,K] * B[K,N], fp32 in / fp32 out, computed on the 5th-generation tensor
// cores as 3xTF32 so fp32 accuracy is kept (north star): with x = hi + lo, hi = the TF32-representable top of x,
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (lo*lo ~ 2^-22 relative is dropped)
// all three products accumulated into the same fp32 TMEM accumulator, small terms first.
//
// Pipeline (sm_100a only — TMA, mbarrier, tcgen05, TMEM):
//   prologue kernels   split A into A_hi / A_lo [M,Kp] and B into B^T_hi / B^T_lo [N,Kp]  (both operands K-major, K zero-padded
//                      to Kp = a multiple of 32 so any K works; ragged M / N edges are TMA out-of-bounds zero fill on the way
//                      in and predicated stores on the way out)
//   gemm kernel        persistent, one CTA per SM, 128 x BN output tile (BN = 256 / 128 / 64 picked per problem so that small
//                      problems still fill the SMs), BLOCK_K = 32 floats (one 128-byte swizzle row)
//     warp 0  TMA producer: 4 tiles per stage (A_hi, A_lo, B_hi, B_lo), SWIZZLE_128B, mbarrier expect_tx
//     warp 1  MMA issuer: one elected thread issues 12 tcgen05.mma.kind::tf32 (M128 N256 K8) per stage,
//             tcgen05.commit releases the smem stage / publishes the accumulator
//     warp 2  TMEM allocator (512 columns = two 128x256 fp32 accumulators, so the epilogue of tile i overlaps tile i+1)
//     warps 4-7  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> 128-bit global stores
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "builtin_kernels.h"
#include "common.h"

namespace cc {
namespace {

constexpr int BM = 128, BK = 32;  // BK floats = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KiB
constexpr int GEMM_THREADS = 256;
constexpr int GROUP_M = 16;  // tile rasterisation: 16 m-tiles share each sweep over n (L2 reuse of the A panels)

template <int BN>
struct Cfg {
  static_assert(BN == 256 || BN == 128 || BN == 64, "BN must be 256, 128 or 64");
  static constexpr int B_TILE_BYTES = BN * BK * 4;                         // 32 / 16 / 8 KiB
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // 96 / 64 / 48 KiB
  static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);       // 192 KiB of operands in flight
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  // all-gather epilogue: 4 epilogue warps x 2 staging buffers x (32 rows x 128 bytes) for the TMA stores to every rank
  static constexpr int GATHER_STAGING_BYTES = 4 * 2 * 4096;
  static constexpr int SMEM_BYTES_GATHER = STAGES * STAGE_BYTES + GATHER_STAGING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators: the epilogue of tile i overlaps the MMAs of tile i+1
  // cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = BN
  static constexpr uint32_t kInstrDesc = (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ | ((uint32_t)(BN >> 3) << 17) |
                                         ((uint32_t)(BM >> 4) << 24);
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > (1u << 26)) __trap();  // a pipeline bug must abort the launch, not hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(x),
               "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major operand, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& m_blk, int& n_blk) {
  const int per_group = GROUP_M * tiles_n;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int rows = min(GROUP_M, tiles_m - first_m);
  const int in_group = tile - group * per_group;
  m_blk = first_m + in_group % rows;
  n_blk = in_group / rows;
}

// ---- the GEMM ------------------------------------------------------------------------------------------------------------

// Destinations of the fused all-gather epilogue: one tensor map per rank, each describing THIS rank's row block [m_shard, N]
// inside that rank's gathered C (so rows past the shard and columns past N are clipped by the TMA unit).
struct GatherMaps {
  CUtensorMap dst[kPeerMaxRanks];
  int world;
  int rank;
};

// Programmatic dependent launch inside one product (split A -> split B -> contraction -> sum of the K splits): a kernel launched with the attribute
// may become resident while its predecessor still runs; it announces its own dependents at once and waits for the predecessor's memory
// before it touches any (CC_GEMM_PDL=0 launches plainly). Saves the launch gaps of small products; nothing at 8192^3.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int BN, bool kGather>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, float* __restrict__ C, int M, int N, int Kp,
                   int tiles_m, int tiles_n, const __grid_constant__ GatherMaps gather, int kb_begin, int accumulate) {
  constexpr int STAGES = Cfg<BN>::STAGES;
  constexpr int STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  constexpr int B_TILE_BYTES = Cfg<BN>::B_TILE_BYTES;
  constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
  constexpr uint32_t kInstrDesc = Cfg<BN>::kInstrDesc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  constexpr int STAGING_BYTES = kGather ? Cfg<BN>::GATHER_STAGING_BYTES : 0;
  const uint32_t staging = smem_base + STAGES * STAGE_BYTES;  // 1024-byte aligned (SWIZZLE_128B boxes)
  const uint32_t bars = staging + STAGING_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  uint8_t* smem_generic = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_generic + STAGES * STAGE_BYTES + STAGING_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = Kp / BK;  // k blocks of THIS launch: [kb_begin, kb_begin + num_kb) of the panels (K chunks, see launch_pipeline)

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a_hi);
    prefetch_tensormap(&tm_a_lo);
    prefetch_tensormap(&tm_b_hi);
    prefetch_tensormap(&tm_b_lo);
    if (kGather)
      for (int d = 0; d < gather.world; ++d) prefetch_tensormap(&gather.dst[d]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 128);  // every epilogue thread arrives
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();  // (set-up done; from here on global memory is read and written)

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t st = smem_base + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(st, &tm_a_hi, full_bar(stage), (kb_begin + kb) * BK, m_blk * BM);
          tma_load_2d(st + A_TILE_BYTES, &tm_a_lo, full_bar(stage), (kb_begin + kb) * BK, m_blk * BM);
          tma_load_2d(st + 2 * A_TILE_BYTES, &tm_b_hi, full_bar(stage), (kb_begin + kb) * BK, n_blk * BN);
          tma_load_2d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &tm_b_lo, full_bar(stage), (kb_begin + kb) * BK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);  // TMA has landed this stage
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + stage * STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(st);
          const uint64_t a_lo = umma_desc_sw128(st + A_TILE_BYTES);
          const uint64_t b_hi = umma_desc_sw128(st + 2 * A_TILE_BYTES);
          const uint64_t b_lo = umma_desc_sw128(st + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // +32 bytes along K inside the 128-byte swizzle row
            umma_tf32(tmem_d, a_lo + adv, b_hi + adv, kInstrDesc, (kb | k) != 0);
            umma_tf32(tmem_d, a_hi + adv, b_lo + adv, kInstrDesc, 1);
            umma_tf32(tmem_d, a_hi + adv, b_hi + adv, kInstrDesc, 1);
          }
          umma_commit(empty_bar(stage));                          // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew + 32)
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + ew * 32 + lane;
      const int col0 = n_blk * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (kGather) {
        // TMEM -> registers -> shared-memory staging -> one TMA store per rank (own HBM and every peer's over NVLink). A store covers 32
        // rows x 64 columns, i.e. 256 contiguous bytes per row: 128-byte row segments (the 32 x 32 boxes of round 1) halve the payload per
        // NVLink packet, and the exchange — 7 peers x this rank's whole block — is what bounds the fused kernel from 4 ranks on. The staging
        // tile is unswizzled (the 128-byte swizzle caps a box at 32 floats per row); its bank-conflicted fills are invisible next to a tile's MMAs.
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          if (col0 + c * 64 >= N) break;
          const uint32_t buf = staging + (uint32_t)(ew * 8192);
          if (lane == 0) bulk_wait_read<0>();  // the stores issued from this buffer one step ago have read it
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + (uint32_t)(c * 64 + h * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t dst = buf + (uint32_t)(lane * 256 + h * 128 + (q << 4));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                           : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // staggered: at any moment every rank targets a different peer (no inbound hot spot), own HBM first
            for (int i = 0; i < gather.world; ++i) {
              int d = gather.rank + i;
              if (d >= gather.world) d -= gather.world;
              tma_store_2d(&gather.dst[d], buf, col0 + c * 64, m_blk * BM + ew * 32);
            }
            bulk_commit();
          }
        }
      } else {
      float* out = C + (size_t)row * (size_t)N + (size_t)col0;
      const bool row_ok = row < M;
      const bool vec_ok = (N & 3) == 0;  // 16-byte aligned row starts
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (col0 + c * 32 >= N) break;  // warp-uniform: the whole 32-column slab is outside the matrix
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (!row_ok) {
          // rows past M (TMA zero fill): nothing to store, but stay converged for the next warp-collective tcgen05.ld
        } else if (vec_ok && col0 + c * 32 + 32 <= N) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            if (accumulate) {  // a later K chunk: C += this chunk's product, added in fp32 with round-to-nearest
              const float4 o = __ldcs(reinterpret_cast<const float4*>(out + c * 32) + q);
              v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
            }
            __stcs(reinterpret_cast<float4*>(out + c * 32) + q, v);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (col0 + c * 32 + q < N) __stcs(out + c * 32 + q, __uint_as_float(r[q]) + (accumulate ? __ldcs(out + c * 32 + q) : 0.f));
        }
        __syncwarp();
      }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(acc));
    }
    if (kGather && lane == 0) {
      bulk_wait_all();  // every store (local and remote) has completed before this CTA retires
      __threadfence_system();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- the same GEMM on CTA pairs (cta_group::2) ----------------------------------------------------------------------------------
//
// ncu on the one-CTA kernel (profiles/r01b_gemm, 4096^3): tensor pipe 80 % active with each CTA pulling 96 KiB of operand panels per
// 1536 MMA cycles (~64 B/clk/SM) through L2 -> shared memory. Two CTAs of one TPC cooperate on a 256 x 256 tile instead: each loads
// its own 128 rows of A_hi / A_lo and only HALF of the B_hi / B_lo tile (128 of the 256 columns); the tensor cores of both SMs read
// both halves (tcgen05.mma.cta_group::2, M = 256). 64 KiB per stage per CTA instead of 96 (a third less L2 / shared-memory traffic
// for the same MMAs) and room for a third stage.
//   CTA 0 (leader): issues every MMA and commit; its `full` barriers collect the TMA bytes of BOTH CTAs (2 x 64 KiB per stage);
//   both CTAs: TMA producer for their own panels, epilogue for their own 128 accumulator rows (own TMEM);
//   commits are multicast to both CTAs (stage free / accumulator ready); epilogue warps of both CTAs arrive on the leader's
//   tmem_empty barriers (remote mbarrier arrive for CTA 1).
constexpr int PAIR_BN = 256;
constexpr int PAIR_STAGES = 3;
constexpr int PAIR_B_HALF_BYTES = (PAIR_BN / 2) * BK * 4;                      // 16 KiB
constexpr int PAIR_STAGE_BYTES = 2 * A_TILE_BYTES + 2 * PAIR_B_HALF_BYTES;     // 64 KiB per CTA
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + 1024 + 256;
constexpr int PAIR_STAGING_BYTES = 4 * 2 * 4096;  // all-gather epilogue: per epilogue warp two 32 x 128-byte staging buffers
constexpr int PAIR_SMEM_BYTES_GATHER = PAIR_SMEM_BYTES + PAIR_STAGING_BYTES;
constexpr uint32_t kPairInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PAIR_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kPairInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <bool kGather>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_3xtf32_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                        const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, float* __restrict__ C, int M, int N,
                        int Kp, int tiles_pm, int tiles_n, const __grid_constant__ GatherMaps gather, int kb_begin, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int STAGING_BYTES = kGather ? PAIR_STAGING_BYTES : 0;
  const uint32_t staging = smem_base + PAIR_STAGES * PAIR_STAGE_BYTES;
  const uint32_t bars = staging + STAGING_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PAIR_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * PAIR_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * PAIR_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * PAIR_STAGES + 4);
  uint8_t* smem_generic = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_generic + PAIR_STAGES * PAIR_STAGE_BYTES + STAGING_BYTES + 8 * (2 * PAIR_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_pm * tiles_n;
  const int num_kb = Kp / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a_hi);
    prefetch_tensormap(&tm_a_lo);
    prefetch_tensormap(&tm_b_hi);
    prefetch_tensormap(&tm_b_lo);
    if (kGather)
      for (int d = 0; d < gather.world; ++d) prefetch_tensormap(&gather.dst[d]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // leader only: its own arrive.expect_tx; the bytes come from both CTAs' TMA loads
      mbar_init(empty_bar(s), 1);  // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);     // one multicast commit
      mbar_init(tmem_empty_bar(a), 256);  // leader only: the 128 epilogue threads of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before anybody signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();  // (set-up done; from here on global memory is read and written)

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A_hi / A_lo, own 128 columns of B_hi / B_lo =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_pm, tiles_n, m_blk, n_blk);
        const int row_a = m_blk * 256 + (int)rank * 128;
        const int row_b = n_blk * PAIR_BN + (int)rank * (PAIR_BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t st = smem_base + stage * PAIR_STAGE_BYTES;
          const uint32_t leader_full = map_to_cta(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * PAIR_STAGE_BYTES);
          tma_load_2d_pair(st, &tm_a_hi, leader_full, (kb_begin + kb) * BK, row_a);
          tma_load_2d_pair(st + A_TILE_BYTES, &tm_a_lo, leader_full, (kb_begin + kb) * BK, row_a);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES, &tm_b_hi, leader_full, (kb_begin + kb) * BK, row_b);
          tma_load_2d_pair(st + 2 * A_TILE_BYTES + PAIR_B_HALF_BYTES, &tm_b_lo, leader_full, (kb_begin + kb) * BK, row_b);
          if (++stage == PAIR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * PAIR_BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);  // both CTAs' panels have landed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + stage * PAIR_STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(st);
          const uint64_t a_lo = umma_desc_sw128(st + A_TILE_BYTES);
          const uint64_t b_hi = umma_desc_sw128(st + 2 * A_TILE_BYTES);
          const uint64_t b_lo = umma_desc_sw128(st + 2 * A_TILE_BYTES + PAIR_B_HALF_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            umma_tf32_pair(tmem_d, a_lo + adv, b_hi + adv, (kb | k) != 0);
            umma_tf32_pair(tmem_d, a_hi + adv, b_lo + adv, 1);
            umma_tf32_pair(tmem_d, a_hi + adv, b_hi + adv, 1);
          }
          umma_commit_pair(empty_bar(stage));                          // frees this stage in BOTH CTAs
          if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar(acc));  // accumulator complete in BOTH CTAs
        }
        __syncwarp();
        if (++stage == PAIR_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): own TMEM (128 rows of the pair's 256) -> registers -> global =====
    const int ew = warp - 4;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_pm, tiles_n, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 256 + (int)rank * 128 + ew * 32 + lane;
      const int col0 = n_blk * PAIR_BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * PAIR_BN);
      if constexpr (kGather) {
#pragma unroll 1
        for (int c = 0; c < PAIR_BN / 64; ++c) {  // 32 rows x 64 columns per store: see the one-CTA kernel
          if (col0 + c * 64 >= N) break;
          const uint32_t buf = staging + (uint32_t)(ew * 8192);
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + (uint32_t)(c * 64 + h * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t dst = buf + (uint32_t)(lane * 256 + h * 128 + (q << 4));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                           : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int i = 0; i < gather.world; ++i) {  // staggered destinations, see the one-CTA kernel
              int d = gather.rank + i;
              if (d >= gather.world) d -= gather.world;
              tma_store_2d(&gather.dst[d], buf, col0 + c * 64, m_blk * 256 + (int)rank * 128 + ew * 32);
            }
            bulk_commit();
          }
        }
      } else {
      float* out = C + (size_t)row * (size_t)N + (size_t)col0;
      const bool row_ok = row < M;
      const bool vec_ok = (N & 3) == 0;
#pragma unroll 1
      for (int c = 0; c < PAIR_BN / 32; ++c) {
        if (col0 + c * 32 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (!row_ok) {
        } else if (vec_ok && col0 + c * 32 + 32 <= N) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            if (accumulate) {  // a later K chunk: C += this chunk's product, added in fp32 with round-to-nearest
              const float4 o = __ldcs(reinterpret_cast<const float4*>(out + c * 32) + q);
              v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
            }
            __stcs(reinterpret_cast<float4*>(out + c * 32) + q, v);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (col0 + c * 32 + q < N) __stcs(out + c * 32 + q, __uint_as_float(r[q]) + (accumulate ? __ldcs(out + c * 32 + q) : 0.f));
        }
        __syncwarp();
      }
      }
      tc_fence_before();
      mbar_arrive_cluster(map_to_cta(tmem_empty_bar(acc), 0));  // the leader's barrier (a remote arrive from CTA 1)
    }
    if (kGather && lane == 0) {
      bulk_wait_all();
      __threadfence_system();
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still read it or signal into it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- prologue: hi / lo split -------------------------------------------------------------------------------------------------

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);  // exactly TF32-representable (the tensor core then has nothing to drop)
  const float r = x - hi;                                  // exact in fp32
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(r));
  lo = __uint_as_float(t);
}

// ---- CTA pairs with A through tensor memory ("config 1024") ------------------------------------------------------------------------
//
// The pair kernel above reads A as hi / lo panels a prologue kernel wrote (A read once, written twice, read again: +0.135 ms and 768 MiB of
// traffic at 8192^3, and a launch that dominates small products). tcgen05.mma can take A from TENSOR memory instead:
//   * A is read as the ORIGINAL fp32 matrix (TMA, SWIZZLE_128B, out-of-bounds rows / k zero filled): no split_a prologue, no A_hi / A_lo
//     panels in HBM, half the L2 -> shared-memory traffic for A. Eight converter warps per CTA (two groups taking alternate k blocks)
//     pull each landed 128 x 32 tile out of shared memory, split it in registers (hi = top 19 bits, lo = tf32(x - hi)) and tcgen05.st the
//     two halves into the stage's TMEM columns: [0, 256) = the accumulator, [256, 512) = four A stages of 64 columns.
//   * B still comes as the K-major hi / lo panels of the (cached) prologue: it is the replicated / weight operand.
//   * barriers per stage: a_full (own CTA: raw A landed) -> converters -> a_ready (leader: both CTAs' A is in TMEM);
//     b_full (leader: both halves of B landed); the leader's MMA warp waits for both, issues 12 MMAs, and one multicast commit (empty)
//     frees the B tiles, the raw A tile and the TMEM A stage in both CTAs.
// Measured (scripts/gpu_gemm_v2.py, profiles/r02_gemm_tmem_a.json): 8192^3 304 vs 300 TFLOP/s (both at the power-limited tensor peak:
// ncu has the tensor pipe 87 % active at 1.72 GHz), 4096^3 220 vs 210, 2048^3 175 vs 153, 1024 x 8192 x 8192 210 vs 193; same bits as the
// panel kernels (same split, same accumulation order). One accumulator instead of two: the epilogue of a tile no longer overlaps the next
// tile's MMAs (~2 % of a K = 8192 tile), which the saved prologue more than pays for.
constexpr int TA_STAGES = 4;
constexpr int TA_A_RAW_BYTES = BM * BK * 4;  // 16 KiB
constexpr int TA_THREADS = 512;  // warp 0 TMA, 1 MMA, 2 TMEM allocator, 3 idle, 4-7 epilogue, 8-15 converters (two groups, alternating stages)
constexpr uint32_t TA_A_COL0 = 256;  // TMEM columns [0, 256) = accumulator(s), [256, 512) = four A stages (hi 32 | lo 32 each)
template <int BN>
struct TaCfg {
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
  static constexpr int ACCS = 256 / BN;  // two 128-column accumulators (epilogue overlaps the next tile) or one of 256
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 4;                      // 8 / 16 KiB
  static constexpr int STAGE_BYTES = TA_A_RAW_BYTES + 2 * B_HALF_BYTES;      // 32 / 48 KiB per CTA
  static constexpr int SMEM_BYTES = TA_STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

__device__ __forceinline__ void umma_tf32_pair_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TA_THREADS, 1)
gemm_3xtf32_tmema_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                         float* __restrict__ C, int M, int N, int Kp, int tiles_pm, int tiles_n, int kb_begin, int accumulate, int splits,
                         long long split_stride) {
  constexpr int ACCS = TaCfg<BN>::ACCS;
  constexpr int B_HALF_BYTES = TaCfg<BN>::B_HALF_BYTES;
  constexpr int STAGE_BYTES = TaCfg<BN>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + TA_STAGES * STAGE_BYTES;
  // barrier layout (8 bytes each): a_full[S], b_full[S], a_ready[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address
  auto a_full_bar = [&](int s) { return bars + 8u * s; };
  auto b_full_bar = [&](int s) { return bars + 8u * (TA_STAGES + s); };
  auto a_ready_bar = [&](int s) { return bars + 8u * (2 * TA_STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (3 * TA_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (4 * TA_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (4 * TA_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (4 * TA_STAGES + 4);
  uint8_t* smem_generic = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_generic + TA_STAGES * STAGE_BYTES + 8 * (4 * TA_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // Work units: (tile, K split). With `splits` > 1 a tile's K range is cut into that many pieces, each computed by another pair into its own
  // partial result (C + split * split_stride) and summed afterwards in a fixed order: small products (1024^3 is 16 tiles for 74 pairs)
  // fill the machine without giving up determinism.
  const int num_tiles = tiles_pm * tiles_n * splits;
  const int total_kb = Kp / BK;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  auto unit_kb = [&](int unit, int& first) {
    const int sp = unit % splits;
    first = kb_begin + sp * kb_per_split;
    return min(kb_per_split, total_kb - sp * kb_per_split);
  };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm_a);
    prefetch_tensormap(&tm_b_hi);
    prefetch_tensormap(&tm_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TA_STAGES; ++s) {
      mbar_init(a_full_bar(s), 1);   // own producer's arrive.expect_tx; bytes from this CTA's A load
      mbar_init(b_full_bar(s), 1);   // leader only: its arrive.expect_tx; bytes from both CTAs' B loads
      mbar_init(a_ready_bar(s), 8);  // leader only: 4 converter warps of each CTA
      mbar_init(empty_bar(s), 1);    // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 256);  // leader only: the 128 epilogue threads of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();  // (barriers, tensor memory and the tensor maps are set up; from here on global memory is read and written)

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of the original A, own half of the B^T hi / lo panels =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk, kb_first;
        tile_coords(tile / splits, tiles_pm, tiles_n, m_blk, n_blk);
        const int num_kb = unit_kb(tile, kb_first);
        const int row_a = m_blk * 256 + (int)rank * 128;
        const int row_b = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t st = smem_base + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(a_full_bar(stage), TA_A_RAW_BYTES);
          tma_load_2d(st, &tm_a, a_full_bar(stage), (kb_first + kb) * BK, row_a);
          const uint32_t leader_b_full = map_to_cta(b_full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(b_full_bar(stage), 2 * 2 * B_HALF_BYTES);
          tma_load_2d_pair(st + TA_A_RAW_BYTES, &tm_b_hi, leader_b_full, (kb_first + kb) * BK, row_b);
          tma_load_2d_pair(st + TA_A_RAW_BYTES + B_HALF_BYTES, &tm_b_lo, leader_b_full, (kb_first + kb) * BK, row_b);
          if (++stage == TA_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only): A from tensor memory, B from shared memory =====
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int acc = ACCS == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = ACCS == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      int kb_first;
      const int num_kb = unit_kb(tile, kb_first);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(b_full_bar(stage), phase);   // both halves of B have landed
        mbar_wait(a_ready_bar(stage), phase);  // both CTAs' converters have written this stage's A into tensor memory
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + stage * STAGE_BYTES;
          const uint64_t b_hi = umma_desc_sw128(st + TA_A_RAW_BYTES);
          const uint64_t b_lo = umma_desc_sw128(st + TA_A_RAW_BYTES + B_HALF_BYTES);
          const uint32_t a_hi = tmem_base + TA_A_COL0 + (uint32_t)(stage * 64);
          const uint32_t a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            umma_tf32_pair_tmem_a(tmem_d, a_lo + (uint32_t)(k * 8), b_hi + adv, TaCfg<BN>::kInstrDesc, (kb | k) != 0);
            umma_tf32_pair_tmem_a(tmem_d, a_hi + (uint32_t)(k * 8), b_lo + adv, TaCfg<BN>::kInstrDesc, 1);
            umma_tf32_pair_tmem_a(tmem_d, a_hi + (uint32_t)(k * 8), b_hi + adv, TaCfg<BN>::kInstrDesc, 1);
          }
          umma_commit_pair(empty_bar(stage));  // frees B, the raw A tile and the TMEM A stage in BOTH CTAs
          if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar(acc));
        }
        __syncwarp();
        if (++stage == TA_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 8) {
    // ===== converters (both CTAs): raw fp32 A tile in shared memory -> hi / lo in registers -> tensor memory. Two groups of four
    // warps take alternate k blocks, so two stages are being converted at any time (one group alone is as slow as the MMAs) =====
    const int group = (warp - 8) >> 2;
    const int cw = (warp - 8) & 3;  // TMEM lanes [32 * cw, 32 * cw + 32) = rows of the A tile
    const int row = cw * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(cw * 32) << 16) + TA_A_COL0;
    long long j = 0;  // k blocks this CTA has seen, over all its tiles
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int kb_first;
      const int num_kb = unit_kb(tile, kb_first);
      for (int kb = 0; kb < num_kb; ++kb, ++j) {
        if ((int)(j & 1) != group) continue;
        const int stage = (int)(j % TA_STAGES);
        const uint32_t phase = (uint32_t)((j / TA_STAGES) & 1);
        mbar_wait(a_full_bar(stage), phase);  // (the producer only refilled this stage after `empty`: its TMEM columns are free too)
        const uint32_t src = smem_base + stage * STAGE_BYTES + (uint32_t)(row * 128);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t x0, x1, x2, x3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(src + (uint32_t)((c ^ (row & 7)) << 4)));
          const uint32_t x[4] = {x0, x1, x2, x3};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float h, l;
            split_tf32(__uint_as_float(x[q]), h, l);
            hi[4 * c + q] = __float_as_uint(h);
            lo[4 * c + q] = __float_as_uint(l);
          }
        }
        tmem_st_32x32(lane_addr + (uint32_t)(stage * 64), hi);
        tmem_st_32x32(lane_addr + (uint32_t)(stage * 64 + 32), lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(a_ready_bar(stage), 0));
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue (both CTAs): own TMEM (128 rows of the pair's 256) -> registers -> global =====
    const int ew = warp - 4;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      int m_blk, n_blk;
      tile_coords(tile / splits, tiles_pm, tiles_n, m_blk, n_blk);
      const int acc = ACCS == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = ACCS == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 256 + (int)rank * 128 + ew * 32 + lane;
      const int col0 = n_blk * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      float* out = C + (size_t)(tile % splits) * (size_t)split_stride + (size_t)row * (size_t)N + (size_t)col0;
      const bool row_ok = row < M;
      const bool vec_ok = (N & 3) == 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (col0 + c * 32 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (!row_ok) {
        } else if (vec_ok && col0 + c * 32 + 32 <= N) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            if (accumulate) {  // a later K chunk: C += this chunk's product, added in fp32 with round-to-nearest
              const float4 o = __ldcs(reinterpret_cast<const float4*>(out + c * 32) + q);
              v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
            }
            __stcs(reinterpret_cast<float4*>(out + c * 32) + q, v);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (col0 + c * 32 + q < N) __stcs(out + c * 32 + q, __uint_as_float(r[q]) + (accumulate ? __ldcs(out + c * 32 + q) : 0.f));
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive_cluster(map_to_cta(tmem_empty_bar(acc), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// A [M,K] row-major -> A_hi / A_lo [M,Kp] (Kp % 32 == 0, columns >= K zero). One thread per 4 output floats.
__global__ void __launch_bounds__(256) split_a_kernel(const float* __restrict__ a, float* __restrict__ hi, float* __restrict__ lo, int M, int K, int Kp) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t vec_per_row = (size_t)Kp / 4;
  const size_t nvec = (size_t)M * vec_per_row;
  const size_t stride = (size_t)gridDim.x * 256;
  const bool aligned = (K & 3) == 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
    const size_t row = i / vec_per_row;
    const int k = (int)(i - row * vec_per_row) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = a + row * (size_t)K + k;
    if (aligned && k + 3 < K) {
      x = __ldcs(reinterpret_cast<const float4*>(src));
    } else {
      if (k < K) x.x = __ldcs(src);
      if (k + 1 < K) x.y = __ldcs(src + 1);
      if (k + 2 < K) x.z = __ldcs(src + 2);
      if (k + 3 < K) x.w = __ldcs(src + 3);
    }
    float4 h, l;
    split_tf32(x.x, h.x, l.x);
    split_tf32(x.y, h.y, l.y);
    split_tf32(x.z, h.z, l.z);
    split_tf32(x.w, h.w, l.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// B [K,N] row-major -> B^T hi / lo [N,Kp] row-major (columns >= K zero), 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256) split_transpose_b_kernel(const float* __restrict__ b, float* __restrict__ bt_hi, float* __restrict__ bt_lo, int K,
                                                                int N, int Kp) {
  __shared__ float th[32][33];
  __shared__ float tl[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int k = k0 + ty + r, n = n0 + tx;
    // (a coherent streaming load, not the ld.global.nc a `const __restrict__` read may become: this grid can be resident — waiting in
    // pdl_wait — while the producer of B is still writing it)
    const float x = (k < K && n < N) ? __ldcs(b + (size_t)k * N + n) : 0.f;
    float h, l;
    split_tf32(x, h, l);
    th[ty + r][tx] = h;
    tl[ty + r][tx] = l;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int n = n0 + ty + r;
    if (n < N) {  // k0 + tx < Kp always: Kp is a multiple of 32
      const size_t o = (size_t)n * Kp + k0 + tx;
      bt_hi[o] = th[tx][ty + r];
      bt_lo[o] = tl[tx][ty + r];
    }
  }
}

void make_map(TensorMapEncodeFn encode, CUtensorMap* map, const float* base, int64_t rows, int64_t k, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(CC_ERR_CUDA, strprintf("cuTensorMapEncodeTiled failed (%d)", (int)r));
}

void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("%s launch failed: %s", what, cudaGetErrorString(e)));
}

}  // namespace

bool gemm_available() { return true; }

int64_t gemm_padded_k(int64_t k) { return (k + BK - 1) / BK * BK; }

// Tile configuration for an M x N problem: 256 x 256 on CTA pairs, or 128 x {256, 128, 64} on single CTAs. Cost model = waves x
// (tile time per SM) with the tensor-pipe efficiencies ncu measured for each variant (profiles/README.md): the wide tiles are the
// most efficient per flop, the narrow ones fill the SMs on small problems.
namespace {
int pick_config_and_cost(int64_t m, int64_t n, int sm_count, bool allow_pair, double* cost_out) {
  struct Cand {
    int code, bm, bn, units_div;
    double eff;
  };
  const Cand cands[4] = {{512, 256, 256, 2, 0.92}, {256, 128, 256, 1, 0.80}, {128, 128, 128, 1, 0.75}, {64, 128, 64, 1, 0.35}};
  int best = 64;
  double best_cost = 1e300;
  for (const Cand& c : cands) {
    if (c.code == 512 && (!allow_pair || m <= BM)) continue;
    const int64_t tiles = ((m + c.bm - 1) / c.bm) * ((n + c.bn - 1) / c.bn);
    const int64_t units = std::max(1, sm_count / c.units_div);
    const int64_t waves = (tiles + units - 1) / units;
    const double cost = (double)waves * (double)c.bn / c.eff;  // per-SM work of one tile is 128 x bn in every variant
    if (cost < best_cost) {
      best_cost = cost;
      best = c.code;
    }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}
// waves x columns per SM / efficiency of the best unsplit configuration (the unit of the picker's cost model)
double gemm_direct_cost(int64_t m, int64_t n, int sm_count, bool allow_pair) {
  double cost;
  pick_config_and_cost(m, n, sm_count, allow_pair, &cost);
  return cost;
}
}  // namespace

int gemm_pick_config(int64_t m, int64_t n, int sm_count, bool allow_pair) { return pick_config_and_cost(m, n, sm_count, allow_pair, nullptr); }

int gemm_pick_bn(int64_t m, int64_t n, int sm_count) {
  const int c = gemm_pick_config(m, n, sm_count, false);
  return c;
}

namespace {
bool tmema_default();
}

// K splits for the tensor-memory-A kernel (256 x 256 pair tiles): the count (<= 8, each split at least 256 deep — below that the pipeline fill
// and the epilogue dominate a unit) the time model below expects to beat every unsplit configuration by 5 %, else 1 = no split.
int gemm_k_splits(int64_t m, int64_t n, int64_t k, int sm_count) {
  if (const char* e = getenv("CC_GEMM_K_SPLITS")) return std::max(1, atoi(e));
  if (k < 512) return 1;
  // Time model in microseconds, fitted to graph replays on a B200 (profiles/r02_gemm_split_k.md): a launch costs its waves x the K range of a
  // unit x the tile's columns per SM / the variant's tensor-pipe efficiency, plus a fixed part (launch, pipeline fill, epilogue); splitting
  // adds the second launch and one pass over the partials (mostly in L2).
  constexpr double kUsPerColumnK = 1.05e-4, kFixedUs = 5.5, kSumLaunchUs = 3.5, kSumBytesPerUs = 4e6;
  const int64_t pairs = std::max(1, sm_count / 2);
  const int64_t tiles = ((m + 255) / 256) * ((n + 255) / 256);
  const double direct = gemm_direct_cost(m, n, sm_count, m > BM) * (double)k * kUsPerColumnK + kFixedUs;
  int best_splits = 1;
  double best = direct * 0.95;  // (a split must win clearly: the model is good to 5 - 10 %)
  for (int s = 2; s <= 8 && k / s >= 256; ++s) {
    const int64_t waves = (tiles * s + pairs - 1) / pairs;
    const int64_t k_unit = ((k + BK - 1) / BK + s - 1) / s * BK;
    const double cost = (double)waves * (double)k_unit * (256.0 / 0.92) * kUsPerColumnK + kFixedUs + kSumLaunchUs +
                        (double)(s + 1) * (double)m * (double)n * 4.0 / kSumBytesPerUs;
    if (cost < best) best = cost, best_splits = s;
  }
  return best_splits;
}

int gemm_config_for(const float* a, int64_t m, int64_t n, int64_t k, int sm_count, bool gather_epilogue) {
  // CC_GEMM_FORCE_CONFIG = 1024 | 512 | 256 | 128 | 64 pins the tile configuration (tests run every variant on the same shapes)
  int config = gemm_pick_config(m, n, sm_count, true);
  // 1024 = A through tensor memory: needs the original A (not pre-gathered panels), a TMA-able row pitch and more than one CTA of rows
  const bool tmema_ok = !gather_epilogue && a && (k & 3) == 0 && ((uintptr_t)a & 15) == 0 && m > BM;
  if (config == 512 && tmema_ok && tmema_default()) config = 1024;
  // a product of few tiles: the tensor-memory-A kernel with K split over the CTA pairs beats a narrow-tile configuration that fills the SMs
  // with inefficient tiles (1024^3: 16 tiles of 256 x 256, 4 splits -> 64 units for 74 pairs)
  if (config != 1024 && tmema_ok && tmema_default() && gemm_k_splits(m, n, k, sm_count) > 1) config = 1024;
  // gather epilogue: every tile's stores leave when the tile is complete, and the receiving side takes ~450 GB/s (8 ranks: 224 MiB per GPU).
  // With fewer than three waves of pair tiles the traffic comes in two bursts and the second one is all tail (8192^3 on 8 ranks: 0.477 ms
  // un-gathered, 0.739 gathered); 128 x 128 tiles are slower per flop but finish continuously: 0.678 ms (profiles/r02_gather_n8_tile_configs.json)
  if (gather_epilogue && config == 512) {
    const int64_t pair_tiles = ((m + 255) / 256) * ((n + 255) / 256), small_tiles = ((m + 127) / 128) * ((n + 127) / 128);
    if (pair_tiles * 2 < (int64_t)3 * sm_count && small_tiles >= (int64_t)3 * sm_count) config = 128;
  }
  if (const char* force = getenv("CC_GEMM_FORCE_CONFIG")) {
    const int f = atoi(force);
    if (f == 256 || f == 128 || f == 64 || (f == 512 && m > BM) || (f == 1024 && tmema_ok)) config = f;
  }
  return config;
}

namespace {
// the tensor-memory-A kernel replaces the plain pair kernel wherever that one is picked, unless CC_GEMM_TMEM_A=0 (A/B timing)
bool tmema_default() {
  static const bool on = [] {
    const char* e = getenv("CC_GEMM_TMEM_A");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

bool gemm_pdl() {
  static const bool on = [] {
    const char* e = getenv("CC_GEMM_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}
// launch with programmatic stream serialisation allowed (the kernel must call pdl_wait() before it touches global memory)
template <class... Params, class... Args>
void launch_dependent_if(bool dependent, void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = dependent && gemm_pdl() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
  if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("cudaLaunchKernelEx: %s", cudaGetErrorString(e)));
}

template <class... Params, class... Args>
void launch_dependent(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  launch_dependent_if(true, kernel, grid, block, smem, stream, std::forward<Args>(args)...);
}

template <int BN, bool kGather>
void launch_main(const GemmWorkspace& ws, float* c, int64_t m, int64_t n, int64_t kp, int sm_count, TensorMapEncodeFn encode, cudaStream_t stream,
                 const GatherMaps& gather, int64_t kb_begin = 0, int64_t kb_count = -1) {
  constexpr int SMEM = kGather ? Cfg<BN>::SMEM_BYTES_GATHER : Cfg<BN>::SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_kernel<BN, kGather>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("cudaFuncSetAttribute(smem=%d): %s", SMEM, cudaGetErrorString(e)));
    attr_set = true;
  }
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  make_map(encode, &ma_hi, ws.a_hi, m, kp, BM);
  make_map(encode, &ma_lo, ws.a_lo, m, kp, BM);
  make_map(encode, &mb_hi, ws.bt_hi, n, kp, BN);
  make_map(encode, &mb_lo, ws.bt_lo, n, kp, BN);
  const int tiles_m = (int)((m + BM - 1) / BM), tiles_n = (int)((n + BN - 1) / BN);
  int grid = tiles_m * tiles_n;
  if (grid > sm_count) grid = sm_count;
  const int64_t kbs = kb_count < 0 ? kp / BK : kb_count;
  // (the gather variant follows a cross-rank flag barrier: it is launched plainly)
  launch_dependent_if(!kGather, gemm_3xtf32_kernel<BN, kGather>, dim3((unsigned)grid), dim3(GEMM_THREADS), SMEM, stream, ma_hi, ma_lo, mb_hi, mb_lo, c, (int)m, (int)n,
                      (int)(kbs * BK), tiles_m, tiles_n, gather, (int)kb_begin, kb_begin > 0 ? 1 : 0);
  check_launch(kGather ? "gemm_3xtf32 (all-gather epilogue)" : "gemm_3xtf32");
}

template <bool kGather>
void launch_pair(const GemmWorkspace& ws, float* c, int64_t m, int64_t n, int64_t kp, int sm_count, TensorMapEncodeFn encode, cudaStream_t stream,
                 const GatherMaps& gather, int64_t kb_begin = 0, int64_t kb_count = -1) {
  constexpr int SMEM = kGather ? PAIR_SMEM_BYTES_GATHER : PAIR_SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_pair_kernel<kGather>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("cudaFuncSetAttribute(pair, smem=%d): %s", SMEM, cudaGetErrorString(e)));
    attr_set = true;
  }
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  make_map(encode, &ma_hi, ws.a_hi, m, kp, BM);
  make_map(encode, &ma_lo, ws.a_lo, m, kp, BM);
  make_map(encode, &mb_hi, ws.bt_hi, n, kp, PAIR_BN / 2);
  make_map(encode, &mb_lo, ws.bt_lo, n, kp, PAIR_BN / 2);
  const int tiles_pm = (int)((m + 255) / 256), tiles_n = (int)((n + PAIR_BN - 1) / PAIR_BN);
  int pairs = tiles_pm * tiles_n;
  if (pairs > sm_count / 2) pairs = sm_count / 2;
  const int64_t kbs = kb_count < 0 ? kp / BK : kb_count;
  launch_dependent_if(!kGather, gemm_3xtf32_pair_kernel<kGather>, dim3((unsigned)(2 * pairs)), dim3(GEMM_THREADS), SMEM, stream, ma_hi, ma_lo, mb_hi, mb_lo, c, (int)m, (int)n,
                      (int)(kbs * BK), tiles_pm, tiles_n, gather, (int)kb_begin, kb_begin > 0 ? 1 : 0);
  check_launch(kGather ? "gemm_3xtf32 (CTA pairs, all-gather epilogue)" : "gemm_3xtf32 (CTA pairs)");
}

// C = sum over s of partial[s] (fixed order: deterministic), 128-bit vectors; n is a multiple of 4 or the tail runs scalar
__global__ void __launch_bounds__(256) sum_k_splits_kernel(const float* __restrict__ partials, float* __restrict__ c, size_t n, int splits, size_t stride) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t nvec = n / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
    float4 acc = __ldcs(reinterpret_cast<const float4*>(partials) + i);
    for (int sp = 1; sp < splits; ++sp) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(partials + (size_t)sp * stride) + i);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    __stcs(reinterpret_cast<float4*>(c) + i, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (n & ~(size_t)3) + threadIdx.x;
    float acc = __ldcs(partials + i);
    for (int sp = 1; sp < splits; ++sp) acc += __ldcs(partials + (size_t)sp * stride + i);
    c[i] = acc;
  }
}

// config 1024: CTA pairs, 256 x 256 tiles (one accumulator), A read as the original fp32 matrix and split inside the kernel (through tensor
// memory). (The kernel also instantiates for 256 x 128 tiles with two accumulators; measured at 8192^3: 211 TFLOP/s, tensor pipe 56 % —
// 128-column MMAs are too short to keep the pipe busy — against 304 for 256 x 256, so only the wide tile is dispatched.)
template <int BN>
int launch_tmema(const float* a, const GemmWorkspace& ws, float* c, int64_t m, int64_t n, int64_t k, int64_t kp, int sm_count, TensorMapEncodeFn encode,
                  cudaStream_t stream, int64_t kb_begin = 0, int64_t kb_count = -1, int splits = 1) {
  constexpr int SMEM = TaCfg<BN>::SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_tmema_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) fail(CC_ERR_CUDA, strprintf("cudaFuncSetAttribute(tmem-A, smem=%d): %s", SMEM, cudaGetErrorString(e)));
    attr_set = true;
  }
  CUtensorMap ma, mb_hi, mb_lo;
  make_map(encode, &ma, a, m, k, BM);  // the ORIGINAL A [M, K]: rows past M and columns past K arrive as zeros (TMA out-of-bounds fill)
  make_map(encode, &mb_hi, ws.bt_hi, n, kp, BN / 2);
  make_map(encode, &mb_lo, ws.bt_lo, n, kp, BN / 2);
  const int tiles_pm = (int)((m + 255) / 256), tiles_n = (int)((n + BN - 1) / BN);
  int pairs = tiles_pm * tiles_n;
  if (pairs > sm_count / 2) pairs = sm_count / 2;
  const int64_t kbs = kb_count < 0 ? kp / BK : kb_count;
  if (splits > 1) {
    // every (tile, K split) unit goes to its own pair and writes its own partial result; the partials are then added in split order by a
    // second kernel. (Adding them in the same launch — the units of a tile meeting on a counter, each summing a slice of rows — measured
    // slower: 1024^3 23.7 us against 21.2 us replayed from a graph, profiles/r02_gemm_split_k.md; the waiting units idle their SMs.)
    CC_REQUIRE(ws.k_split_partials, CC_ERR_ILLEGAL_ARGUMENT, "split-K needs a partials workspace");
    pairs = tiles_pm * tiles_n * splits;
    if (pairs > sm_count / 2) pairs = sm_count / 2;
    const size_t out_floats = (size_t)m * (size_t)n;
    launch_dependent(gemm_3xtf32_tmema_kernel<BN>, dim3((unsigned)(2 * pairs)), dim3(TA_THREADS), SMEM, stream, ma, mb_hi, mb_lo, ws.k_split_partials, (int)m, (int)n,
                     (int)(kbs * BK), tiles_pm, tiles_n, (int)kb_begin, 0, splits, (long long)out_floats);
    check_launch("gemm_3xtf32 (CTA pairs, A through tensor memory, split K)");
    size_t blocks = (out_floats / 4 + 255) / 256;
    if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
    if (blocks == 0) blocks = 1;
    launch_dependent(sum_k_splits_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, (const float*)ws.k_split_partials, c, out_floats, splits, out_floats);
    check_launch("sum_k_splits");
    return 2;
  }
  launch_dependent(gemm_3xtf32_tmema_kernel<BN>, dim3((unsigned)(2 * pairs)), dim3(TA_THREADS), SMEM, stream, ma, mb_hi, mb_lo, c, (int)m, (int)n, (int)(kbs * BK), tiles_pm,
                   tiles_n, (int)kb_begin, kb_begin > 0 ? 1 : 0, 1, 0ll);
  check_launch("gemm_3xtf32 (CTA pairs, A through tensor memory)");
  return 1;
}

template <bool kGather>
int launch_pipeline(const float* a, const float* b, float* c, int64_t m, int64_t n, int64_t k, const GemmWorkspace& ws, int sm_count,
                    TensorMapEncodeFn encode, cudaStream_t stream, bool b_panels_ready, const GatherMaps& gather, bool a_panels_ready = false) {
  CC_REQUIRE(m >= 1 && n >= 1 && k >= 1 && m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31) - BK, CC_ERR_UNSUPPORTED,
             "gemm_3xtf32: bad shape %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
  CC_REQUIRE(encode, CC_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable");
  const int64_t kp = gemm_padded_k(k);
  const int config = a_panels_ready ? gemm_config_for(nullptr, m, n, k, sm_count, kGather) : gemm_config_for(a, m, n, k, sm_count, kGather);
  const bool split_a_needed = !a_panels_ready && config != 1024;
  const size_t nvec = (size_t)(m * kp) / 4;
  size_t blocks = (nvec + 255) / 256;
  if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
  if (split_a_needed) {
    launch_dependent(split_a_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, a, ws.a_hi, ws.a_lo, (int)m, (int)k, (int)kp);
    check_launch("split_a");
  }
  if (!b_panels_ready) {
    launch_dependent(split_transpose_b_kernel, dim3((unsigned)((n + 31) / 32), (unsigned)(kp / 32)), dim3(256), 0, stream, b, ws.bt_hi, ws.bt_lo, (int)k, (int)n, (int)kp);
    check_launch("split_transpose_b");
  }
  // The tensor core accumulates the 3 * K / 8 partial products of an output in fp32 with TRUNCATION, so the error of one launch grows
  // linearly in K (4.9e-6 |A||B| at K = 8192 on normal data; the bar is 1e-5). Beyond kMaxChunkK the product is computed in K chunks whose
  // results are added in the epilogue with round-to-nearest (C += chunk): the error then stays at the one-chunk level for any K, for one
  // extra read of C per chunk. (The all-gather epilogue stores through tensor maps and is not chunked.)
  constexpr int64_t kMaxChunkK = 8192;
  const int64_t kb_total = kp / BK;
  const int64_t kb_chunk = kGather ? kb_total : std::min<int64_t>(kb_total, kMaxChunkK / BK);
  const int k_splits = config == 1024 && ws.k_split_partials ? gemm_k_splits(m, n, k, sm_count) : 1;
  int launches = 0;
  for (int64_t kb0 = 0; kb0 < kb_total; kb0 += kb_chunk, ++launches) {
    const int64_t kbs = std::min<int64_t>(kb_chunk, kb_total - kb0);
    switch (config) {
      case 1024:
        if constexpr (!kGather)
          launches += launch_tmema<256>(a, ws, c, m, n, k, kp, sm_count, encode, stream, kb0, kbs, kb_total <= kb_chunk ? k_splits : 1) - 1;
        break;
      case 512: launch_pair<kGather>(ws, c, m, n, kp, sm_count, encode, stream, gather, kb0, kbs); break;
      case 256: launch_main<256, kGather>(ws, c, m, n, kp, sm_count, encode, stream, gather, kb0, kbs); break;
      case 128: launch_main<128, kGather>(ws, c, m, n, kp, sm_count, encode, stream, gather, kb0, kbs); break;
      default: launch_main<64, kGather>(ws, c, m, n, kp, sm_count, encode, stream, gather, kb0, kbs); break;
    }
  }
  return launches + (split_a_needed ? 1 : 0) + (b_panels_ready ? 0 : 1);
}
}  // namespace

int launch_gemm_3xtf32_panels(float* c, int64_t m, int64_t n, int64_t k, const GemmWorkspace& ws, int sm_count, TensorMapEncodeFn encode,
                              cudaStream_t stream) {
  GatherMaps none{};
  return launch_pipeline<false>(nullptr, nullptr, c, m, n, k, ws, sm_count, encode, stream, true, none, true);
}

int launch_gemm_3xtf32(const float* a, const float* b, float* c, int64_t m, int64_t n, int64_t k, const GemmWorkspace& ws, int sm_count,
                       TensorMapEncodeFn encode, cudaStream_t stream, bool b_panels_ready) {
  GatherMaps none{};
  return launch_pipeline<false>(a, b, c, m, n, k, ws, sm_count, encode, stream, b_panels_ready, none);
}

int launch_gemm_3xtf32_allgather(const float* a, const float* b, float* const* gathered_c, int world, int rank, int64_t m_shard, int64_t n, int64_t k,
                                 const GemmWorkspace& ws, int sm_count, TensorMapEncodeFn encode, cudaStream_t stream, bool b_panels_ready,
                                 float* multicast_c) {
  CC_REQUIRE(world >= 1 && world <= kPeerMaxRanks && rank >= 0 && rank < world, CC_ERR_ILLEGAL_ARGUMENT, "bad world / rank");
  CC_REQUIRE(n % 4 == 0, CC_ERR_UNSUPPORTED, "the all-gather epilogue needs N %% 4 == 0 (16-byte row pitch for the TMA stores)");
  CC_REQUIRE(encode, CC_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable");
  GatherMaps g{};
  // NVLS: ONE destination — the multicast mapping of the gathered C. A store through it is replicated by the NVSwitch into every rank's
  // copy (this rank's own included), so a block leaves the GPU once instead of once per peer.
  g.world = multicast_c ? 1 : world;
  g.rank = multicast_c ? 0 : rank;
  for (int d = 0; d < g.world; ++d) {
    // rank `rank`'s row block inside rank d's gathered C: [m_shard, N] at row offset rank * m_shard
    float* base = (multicast_c ? multicast_c : gathered_c[d]) + (size_t)rank * (size_t)m_shard * (size_t)n;
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m_shard};
    cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    cuuint32_t box[2] = {64, 32};  // 32 rows x 256 contiguous bytes, from an unswizzled staging tile
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&g.dst[d], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(CC_ERR_CUDA, strprintf("cuTensorMapEncodeTiled(gather destination %d) failed (%d)", d, (int)r));
  }
  return launch_pipeline<true>(a, b, nullptr, m_shard, n, k, ws, sm_count, encode, stream, b_panels_ready, g);
}

}  // namespace cc
