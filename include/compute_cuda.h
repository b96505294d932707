/* compute_cuda.h — C ABI of libcompute_cuda.so, the B200 (sm_100a) execution backend that sits where
 * Compute.scala's `OpenCL` runtime trait and `OpenCLKernelBuilder` code generator sit today.
 *
 * Every entry point takes scalars and plain pointers only, so it is callable from the JVM through LWJGL's
 * `SharedLibrary` + `JNI.invoke*` without bespoke JNI glue, and from Python via ctypes.
 * Citations are `file:line` under the reference tree:
 *   O: OpenCL/src/main/scala/com/thoughtworks/compute/OpenCL.scala
 *   T: Tensors/src/main/scala/com/thoughtworks/compute/Tensors.scala
 *   K: OpenCLKernelBuilder/src/main/scala/com/thoughtworks/compute/OpenCLKernelBuilder.scala
 *   R: Trees/src/main/scala/com/thoughtworks/compute/Trees.scala
 *
 * Conventions
 *   - every function returns 0 (CC_OK) or a negative cc_status; the message for the last failure on the calling
 *     thread is returned by cc_last_error() (replaces checkErrorCode -> typed exception, O:251-312; NVRTC build logs
 *     are included like the OpenCL build log, O:905-915);
 *   - handles are opaque 64-bit values; every handle returned to the caller is already retained once and is
 *     released with the matching *_release (deterministic, like clRetain / clRelease, O:636-648, never GC);
 *   - all calls are thread safe and non-blocking unless stated;
 *   - ordering between commands is by event wait lists, not by submission order (T:1363,1374); the library adds
 *     the write-after-read / write-after-write edges needed when pooled memory is reused.
 *   - there is NO CPU fallback: without a CUDA driver and an sm_100 device cc_init fails.
 */
#ifndef COMPUTE_CUDA_H
#define COMPUTE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t cc_buffer; /* DeviceBuffer[Float]            O:636-715 */
typedef uint64_t cc_event;  /* Event                          O:565-612 */
typedef uint64_t cc_kernel; /* Program + Kernel (CompiledKernel, T:1263-1265) */
typedef uint64_t ct_tensor; /* a `Tensor` of the host-side mirror (T:636-1261) */
typedef uint64_t cc_graph;  /* a captured sequence of kernel launches (CUDA graph) */

typedef enum cc_status {
  CC_OK = 0,
  CC_ERR_ILLEGAL_ARGUMENT = -1, /* IllegalArgumentException (T:208-222, 816-855, ...) */
  CC_ERR_NOT_INITIALIZED = -2,
  CC_ERR_NO_DRIVER = -3, /* libcuda.so.1 / device missing: the product path fails loudly */
  CC_ERR_CUDA = -4,      /* CUresult != CUDA_SUCCESS   (O:143-312) */
  CC_ERR_COMPILE = -5,   /* NVRTC failure, log in cc_last_error (O:172-181, 885-915) */
  CC_ERR_BAD_TREE = -6,  /* malformed tree blob */
  CC_ERR_NCCL = -7,
  CC_ERR_UNSUPPORTED = -8,
  CC_ERR_OUT_OF_MEMORY = -9 /* OutOfResources / MemObjectAllocationFailure (O:160-170) */
} cc_status;

/* ---- library / device ------------------------------------------------------------------------------------- */

/* Creates (retains) the primary context of `device_ordinal`, the stream pool and the memory pools.
 * Replaces platform/device discovery + clCreateContext + CommandQueuePool (O:340-374, 423-448, 1376-1393).
 * Idempotent; `device_ordinal < 0` means CUDA_VISIBLE_DEVICES-relative device LOCAL_RANK or 0. */
int cc_init(int device_ordinal);
/* monadicClose: drops the kernel cache, pools, streams, communicator and the context (O:1331-1337, T:1267-1289). */
int cc_shutdown(void);
int cc_is_initialized(void);
/* thread-local message of the last failing call on this thread */
const char* cc_last_error(void);
const char* cc_version(void);

typedef struct cc_device_info_t {
  int32_t ordinal;
  int32_t sm_count;            /* = CommandQueue.deviceId.maxComputeUnits (O:545) */
  int32_t cc_major, cc_minor;  /* 10, 0 on B200 */
  int32_t max_smem_per_block;  /* opt-in bytes */
  int32_t l2_bytes;
  int64_t total_mem;
  int32_t sm_clock_khz, mem_clock_khz;
  char name[64];
} cc_device_info_t;
int cc_device_info(cc_device_info_t* out);
int cc_device_count(int* out);

/* Number of compute streams commands are spread over (replaces numberOfCommandQueuesPerDevice, cpu.scala:115).
 * 1 = one in-order stream (what bench.py uses so CUDA-event timing brackets every kernel). Default 4. */
int cc_set_stream_count(int n);

/* ---- memory ---------------------------------------------------------------------------------------------- */

/* allocateBuffer[Float](n) (O:1399-1411): pooled cuMemAlloc, size classes, contents undefined. */
int cc_buffer_alloc(uint64_t n_floats, cc_buffer* out);
/* allocateBufferFrom(hostBuffer) (O:1415-1431, CL_MEM_COPY_HOST_PTR): async H2D on the copy stream; the host memory
 * must stay valid until `*out_event` completes (pass NULL to make the call blocking). Pinned host memory
 * (cc_host_alloc) copies at full PCIe speed; pageable memory is staged by the driver. */
int cc_buffer_from_host(const float* host, uint64_t n_floats, cc_buffer* out, cc_event* out_event);
/* overwrite an existing buffer from host memory (used by benchmarks to refresh inputs without reallocating) */
int cc_buffer_upload(cc_buffer buf, const float* host, uint64_t n_floats, const cc_event* waits, int n_waits,
                     cc_event* out_event);
/* Adopt device memory owned by someone else (e.g. a torch tensor's data_ptr) — never freed by the library. */
int cc_buffer_wrap(uint64_t device_ptr, uint64_t n_floats, cc_buffer* out);
int cc_buffer_retain(cc_buffer b);  /* DeviceBuffer.retain  O:644 */
int cc_buffer_release(cc_buffer b); /* DeviceBuffer.release O:646 — back to the pool at refcount 0 */
int cc_buffer_device_ptr(cc_buffer b, uint64_t* out_ptr);
int cc_buffer_length(cc_buffer b, uint64_t* out_n_floats);
/* DeviceBuffer.toHostBuffer / enqueueReadBuffer (O:698-715, 1206-1244): async D2H of n floats starting at
 * `offset_floats` after `waits`; `*out_event` completes when `host` is filled (NULL => blocking). */
int cc_buffer_to_host(cc_buffer b, uint64_t offset_floats, float* host, uint64_t n_floats, const cc_event* waits,
                      int n_waits, cc_event* out_event);
/* dst[0..n) = src[0..n) on the device, ordered like every other command (after the writers of src, after the users of dst) */
int cc_buffer_copy(cc_buffer dst, cc_buffer src, uint64_t n_floats, const cc_event* waits, int n_waits, cc_event* out_event);
/* gives every idle pooled device block back to the driver (the pool also does this by itself when an allocation fails) */
int cc_memory_trim(void);
/* pinned host staging memory (replaces LWJGL memAllocFloat, Memory.scala:184-208). Blocks are pooled by size class:
 * cc_host_free returns a block to the pool (cuMemHostAlloc is far too slow for a per-read-back allocation); everything
 * is unpinned and freed by cc_shutdown. */
int cc_host_alloc(uint64_t bytes, void** out);
int cc_host_free(void* p);
/* the device-side address of a cc_host_alloc block (mapped pinned memory): with cc_buffer_wrap it lets a kernel store a
 * small result straight into host memory instead of paying a separate copy command */
int cc_host_device_ptr(void* host, uint64_t* out_device_ptr);

/* ---- events ----------------------------------------------------------------------------------------------- */

int cc_event_retain(cc_event e);  /* O:607 */
int cc_event_release(cc_event e); /* O:609 */
int cc_event_wait(cc_event e);    /* blocking: Event.waitForComplete (O:603-605) */
int cc_event_query(cc_event e, int* out_done); /* waitForStatus probe (O:592-600) */
/* clSetEventCallback replacement (O:1246-1263): cb(user, status) runs on a driver thread after `e` completes */
typedef void (*cc_event_callback)(void* user, int status);
int cc_event_on_complete(cc_event e, cc_event_callback cb, void* user);
/* block until every stream of the pool is idle */
int cc_synchronize(void);

/* ---- expression trees -> kernels ---------------------------------------------------------------------------- */

/* Tree blob (little endian, all fields 32-bit unless noted):
 *   u32 magic 'CCT1' (0x31544343), u32 n_nodes, u32 root, u32 out_rank, i32 out_shape[out_rank], then n_nodes records,
 *   children before parents (post-order), each `u32 kind` + payload:
 *     1 FloatLiteral    f32 value                                             R:373-380
 *     2 ArrayParameter  u64 id, f32 padding, u32 rank, i32 shape[rank], i32 definition_root (-1 = none)   R:755-823
 *                       (`id` = identity of the producing Tensor, T:1259; `definition_root` optionally points at the
 *                        closure of a not-yet-evaluated InlineTensor so patterns can be matched through the barrier)
 *     3 Transform       u32 array, u32 rows, u32 cols, f64 matrix[rows*cols]   R:676-690
 *     4 Extract         u32 array                                             R:660-672
 *     5 Concatenate     u32 n, u32 element[n]                                 R:953-973
 *     6 ConcatenateAt   u32 n, u32 position, u32 element[n]   (root only) the element index becomes output dimension
 *                       `position` instead of the last one: Tensor.join(tensors, dimension), T:560-575, in one kernel
 *     10 Exp 11 Log 12 Abs 13 Tanh 14 Sqrt 15 UnaryMinus     u32 operand       R:384-470,620-658
 *     20 Min 21 Max 22 Plus 23 Minus 24 Times 25 Div 26 Percent   u32 lhs, u32 rhs   R:472-618
 *     30 Reduce         u32 monoid (22 Plus | 20 Min | 21 Max | 24 Times), u32 operand, u32 rank, i32 shape[rank]
 *                       root only, out_shape = []: folds the operand over the index space `shape` (T:303-393, 673-771;
 *                       lets the backend fuse the operand's closure into the reduction instead of materialising it)
 *
 * cc_compile = cache probe by structure (parameters numbered by first visit, literals / shapes / paddings / matrices
 * part of the key — R:70-91,152-177,336-369) and on a miss: pattern matching (axis reduction / contraction),
 * CUDA C++ generation from the sm_100a templates, NVRTC for sm_100a, module load (T:1291-1331, K:135-221). */
int cc_compile(const void* tree_blob, uint64_t n_bytes, cc_kernel* out);
/* Same, and also reports the blob's parameter ids in ordinal order (identity-deduplicated DFS pre-order of the main
 * tree = parameterDescendants, T:230-251, followed by the parameters first met inside definitions) so the caller can
 * map cc_kernel_arg_param ordinals back to its own tensors. `param_ids_out` may be NULL. */
int cc_compile_ex(const void* tree_blob, uint64_t n_bytes, cc_kernel* out, uint64_t* param_ids_out, int capacity,
                  int* n_params_out);
/* kernelCache policy (T:1267-1289): the reference's cache is an overridable Guava CacheBuilder, unbounded by default. A non-zero
 * limit evicts least-recently-used kernels (their modules unload once no caller holds them); clear = clearCache (T:1282-1285). */
/* Opt-in on-disk cubin cache (also: environment variable CC_KERNEL_CACHE_DIR). The reference's kernel cache dies
 * with the process (T:1267-1289) so every run pays the JIT again; with a directory set, the cubin of each generated source is kept
 * under <dir>/<hash of source + compiler identity>.cubin (written atomically, verified against the source on load) and NVRTC is
 * skipped on a hit. NULL or "" turns it off. */
int cc_kernel_disk_cache(const char* directory);
int cc_kernel_cache_limit(uint64_t max_kernels);
int cc_kernel_cache_clear(void);
int cc_kernel_cache_size(uint64_t* out);
/* kernelCache.getIfPresent (T:1293; TensorsSpec.scala:50-52): probe only, never compiles. `*out` = the cached kernel with the
 * blob's structure, retained for the caller, or 0. The library's key includes the output shape (the reference's kernels take it at
 * launch, T:1373); with `any_out_shape` != 0 the shape in the blob header is ignored and any cached output shape matches. */
int cc_kernel_cache_lookup(const void* tree_blob, uint64_t n_bytes, int any_out_shape, cc_kernel* out);
int cc_kernel_retain(cc_kernel k);
int cc_kernel_release(cc_kernel k);

typedef struct cc_kernel_info_t {
  int32_t kind;       /* 0 elementwise, 1 axis reduction, 2 contraction (tcgen05), 3 tiled-transpose elementwise, 4 whole-tensor fold */
  int32_t cache_hit;  /* 1 if this cc_compile call was served from the structural cache */
  int32_t n_args;     /* number of buffers cc_launch expects */
  int32_t n_launches; /* device kernels per cc_launch */
  uint64_t out_floats;
  uint64_t algorithmic_bytes; /* bytes a perfect implementation moves per launch */
  uint64_t flops;
  uint64_t structural_hash;
} cc_kernel_info_t;
int cc_kernel_info(cc_kernel k, cc_kernel_info_t* out);
/* which tree parameter (ordinal in identity-deduplicated DFS pre-order, definitions' parameters numbered after the
 * main tree's — T:230-251) the i-th buffer argument is */
int cc_kernel_arg_param(cc_kernel k, int i, int32_t* out_param_ordinal);
/* generated CUDA C++ (NULL-terminated, owned by the kernel) */
int cc_kernel_source(cc_kernel k, const char** out);
/* Launch geometry of the i-th kernel of a compiled plan (i < cc_kernel_info_t.n_launches) — introspection for tests and tools:
 * the entry point inside the generated module, grid / block / dynamic shared memory, and its argument list (>= 0: plan argument,
 * -1: the output buffer, -2-k: scratch buffer k of `scratch_floats`, -100 / -101: the runtime's fold partials / block counter,
 * -102: the per-stream block counters of a fused axis-reduction second stage). */
typedef struct cc_launch_info_t {
  char entry[64];
  uint32_t grid[3], block[3], smem;
  int32_t n_args;
  int32_t args[32];
  int32_t n_scratch;
  uint64_t scratch_floats[8];
} cc_launch_info_t;
int cc_kernel_launch_info(cc_kernel k, int i, cc_launch_info_t* out);

/* Kernel.enqueue + dispatch (O:788-844, 1298-1329; T:1342-1375): args in cc_kernel_arg_param order. */
int cc_launch(cc_kernel k, const cc_buffer* args, int n_args, cc_buffer out, const cc_event* waits, int n_waits,
              cc_event* out_event);

/* Tensor.sum's reduction programs (T:303-393, 673-771): out[0] = sum(in[0..n)) — deterministic two-stage
 * vector-load / warp-shuffle / shared-memory reduction. */
int cc_reduce_sum(cc_buffer in, uint64_t n_floats, cc_buffer out, const cc_event* waits, int n_waits,
                  cc_event* out_event);
/* Tensor.random / randomNormal kernels (T:398-443, 479-524) */
int cc_random(cc_buffer out, uint64_t n_floats, int32_t seed, cc_event* out_event);
int cc_random_normal(cc_buffer out, uint64_t n_floats, int32_t seed, cc_event* out_event);

/* C[M,N] = A[M,K] * B[K,N] (row-major fp32) with 3xTF32 tcgen05 MMAs; what the contraction pattern lowers to.
 * Exposed for direct measurement; `c` may alias neither input. Any M, N, K >= 1 (ragged edges: TMA zero fill on the way in,
 * predicated stores on the way out; K is zero-padded inside the hi/lo workspace). */
int cc_matmul_3xtf32(cc_buffer a, cc_buffer b, cc_buffer c, int64_t m, int64_t n, int64_t k, const cc_event* waits,
                     int n_waits, cc_event* out_event);

/* The contraction splits B into two K-major TF32 panels (hi / lo) before the tensor-core pipeline runs. The panels of the
 * last few B operands are kept while the B buffer has not been written since (tracked per buffer by the runtime; wrapped
 * memory is never cached), so a replicated / weight operand is split once. 1 = on (default), 0 = off and drop the panels. */
int cc_set_operand_cache(int on);

/* ---- replaying a sequence of evaluations as ONE CUDA graph ------------------------------------------------------- */
/* The reference's API is one slow action = one kernel launch, and small expressions are bound by the host's launch rate (a fused
 * tanh(a*b+c) over 1024^2 floats runs in ~2 us; submitting it costs ~3 us). A loop whose iterations evaluate the same expressions
 * can be captured once and replayed: between cc_graph_begin and cc_graph_end every kernel launch (cc_launch, cc_reduce_sum,
 * cc_random*, cc_matmul_3xtf32; through the ct_* mirror: doBuffer / doCache of any tensor) is RECORDED on one stream instead of
 * executed — buffers are allocated and released as usual, but hold no results yet; copies and collectives are refused
 * (CC_ERR_UNSUPPORTED). cc_graph_launch then runs the whole sequence with one driver call, in capture order; it can be launched any
 * number of times. Buffers the captured commands touched stay alive (and keep their addresses) as long as the graph; a buffer the
 * caller still holds from the capture is refreshed by every replay. cc_graph_begin synchronises the device; captures do not nest and
 * are not concurrent with other threads' commands. */
int cc_graph_begin(void);
int cc_graph_end(cc_graph* out);
int cc_graph_launch(cc_graph g, const cc_event* waits, int n_waits, cc_event* out_event);
int cc_graph_info(cc_graph g, uint64_t* out_commands, uint64_t* out_buffers);
int cc_graph_release(cc_graph g);

/* ---- counters / timing -------------------------------------------------------------------------------------- */

typedef struct cc_stats_t {
  uint64_t compiles, cache_hits, launches, device_kernels, h2d_bytes, d2h_bytes, alloc_calls, pool_hits,
      bytes_in_use, bytes_pooled;
  uint64_t nvrtc_compiles, disk_cache_hits; /* of `compiles` (structures planned): how many ran NVRTC / came from the disk cache */
} cc_stats_t;
int cc_stats(cc_stats_t* out);
int cc_stats_reset(void);
/* Built-in command profiler. The reference has none (queues are created without CL_QUEUE_PROFILING_ENABLE, O:431-436).
 * While enabled every command (kernel launch, copy, collective) is bracketed by timing events on its own stream.
 * cc_profile_report synchronises and writes a JSON array aggregated per kernel structure / copy direction / collective:
 * count, total / avg / min / max device time, the algorithmic bytes and flops the code generator attributes to one launch
 * and the resulting GB/s, TFLOP/s. Call with out = NULL to size (*out_needed), then with a buffer; the records are
 * consumed by the report. */
/* (With CC_NVTX=1 in the environment every command is also an NVTX range of the same name, so that an external profiler can
 * select one expression's kernels: ncu --nvtx --nvtx-include "elementwise #1a2b3c4d/" ...) */
int cc_profile_enable(int on);
int cc_profile_report(char* out, uint64_t capacity, uint64_t* out_needed);
/* device-side stopwatch: joins every pool stream, records a timing event; stop returns elapsed milliseconds */
int cc_timer_start(void);
int cc_timer_stop(float* out_ms);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch --------------------------------------------- */

int cc_comm_unique_id(void* out_128_bytes);                       /* rank 0; ship to peers out of band */
int cc_comm_init(const void* id_128_bytes, int n_ranks, int rank); /* ncclCommInitRank on this process' device */
int cc_comm_destroy(void);
int cc_comm_info(int* out_n_ranks, int* out_rank);
/* counts cc_comm_init calls: a front end that caches per-communicator objects (symmetric arenas) drops them when it changes */
int cc_comm_generation(uint64_t* out);
/* Our own collective over NVLink peer memory (B200 HGX: every peer at full bandwidth through NVSwitch): maps one small
 * CUDA-IPC mailbox per rank into every other rank (handles exchanged once through the communicator). Afterwards
 * cc_allreduce_sum routes vectors of <= 65536 floats through a one-shot kernel (peers store straight into each other's
 * HBM, flags, rank-ordered sum: deterministic, one launch, no NCCL call) and cc_reduce_sum_allreduce fuses Tensor.sum's
 * reduction with the all-reduce of its result in ONE kernel. Returns CC_ERR_UNSUPPORTED if peer access is impossible. */
int cc_comm_enable_peer(void);
int cc_comm_peer_enabled(int* out);
/* route small all-reduces through the peer mailboxes (1, default after cc_comm_enable_peer) or through NCCL (0) — for A/B timing */
int cc_comm_route_peer(int on);
/* out[0] = sum over all ranks of sum(in[0..n)) — collective: every rank of the communicator must call it in the same order */
int cc_reduce_sum_allreduce(cc_buffer in, uint64_t n_floats, cc_buffer out, const cc_event* waits, int n_waits,
                            cc_event* out_event);
int cc_allreduce_sum(cc_buffer buf, uint64_t n_floats, const cc_event* waits, int n_waits, cc_event* out_event);
/* Symmetric memory: every rank allocates `n_floats` and maps every other rank's allocation (CUDA IPC over NVLink). Collective;
 * needs cc_comm_enable_peer. The handle behaves like any cc_buffer; the memory itself lives until cc_comm_destroy. */
int cc_comm_symmetric_alloc(uint64_t n_floats, cc_buffer* out);
/* *out = 1 if the symmetric buffer also has an NVLS multicast mapping (NVSwitch systems with multicast support: the allocation binds every
 * rank's memory to one multicast object, so a store through that mapping is replicated into every rank's copy by the switch and the
 * fused all-gather sends each block once instead of once per peer), 0 if it is mapped peer by peer over CUDA IPC. CC_MULTICAST=0 (on every
 * rank) forces the peer-by-peer mapping. */
int cc_buffer_is_multicast(cc_buffer b, int* out);
/* The row-sharded matmul (A and C row-sharded, B replicated) fused with the all-gather of its result in ONE tensor-core
 * kernel: rank r computes C[r*m_shard .. (r+1)*m_shard, :] = A_shard * B and the epilogue TMA-stores every 32x32 block straight into
 * `gathered` ([n_ranks * m_shard, N], from cc_comm_symmetric_alloc) on EVERY rank — its own HBM and the peers' over NVLink — so the
 * exchange overlaps the MMAs tile by tile instead of following them as an ncclAllGather. Collective (same m_shard on every rank),
 * bracketed by two flag barriers over the peer mailboxes. Needs N % 4 == 0. */
int cc_matmul_3xtf32_allgather(cc_buffer a_shard, cc_buffer b, cc_buffer gathered, int64_t m_shard, int64_t n, int64_t k,
                               const cc_event* waits, int n_waits, cc_event* out_event);
/* recv[rank*n .. (rank+1)*n) = send[0..n) of every rank. With peer mailboxes enabled, blocks of <= 65536 floats per rank go through
 * a one-shot kernel over NVLink peer memory (like cc_allreduce_sum); larger ones through ncclAllGather. */
int cc_allgather(cc_buffer send, cc_buffer recv, uint64_t n_floats_per_rank, const cc_event* waits, int n_waits,
                 cc_event* out_event);
int cc_broadcast(cc_buffer buf, uint64_t n_floats, int root, const cc_event* waits, int n_waits, cc_event* out_event);

/* ---- leading-axis sharding over the GPUs of one box --------------------------------------------------- */
/* Row-major tensors shard into contiguous row blocks, one per rank (process); elementwise graphs and views that do not mix the
 * leading axis need no exchange. These entry points are what a front end needs besides cc_launch to evaluate a kernel on its row
 * block and combine across ranks; the ct_* mirror (ct_shard / ct_gather below) and scala/.../CudaSharding.scala are built on them.
 * All of them except cc_shard_rows are COLLECTIVE: every rank of the communicator calls them in the same order. */

/* the block of `rank`: the first rows % n_ranks ranks own one extra row */
int cc_shard_rows(int64_t rows, int n_ranks, int rank, int64_t* out_first_row, int64_t* out_row_count);
/* *out_all_equal = 1 iff every rank passed the same `value` (exact for the full 64 bits). Gathers need equal blocks on every rank:
 * front ends call this once per block size before the first cc_allgather / cc_shard_launch_allgather of that size, so that uneven
 * shards are an IllegalArgument on EVERY rank instead of a hang in the exchange. */
int cc_shard_agree(uint64_t value, int* out_all_equal);
/* cc_launch, then all-reduce (sum) of the kernel's output across ranks, in place: partial reductions over the sharded axis
 * (column sums `shard.split(0).reduce(_ + _)`, sums of inline expressions). Vectors of <= 65536 floats go through the one-shot
 * kernel over the NVLink peer mailboxes when those are mapped, NCCL otherwise. */
int cc_shard_launch_allreduce(cc_kernel k, const cc_buffer* args, int n_args, cc_buffer out, const cc_event* waits, int n_waits,
                              cc_event* out_event);
/* cc_launch on this rank's row block with the result gathered on every rank: `gathered` holds n_ranks blocks of
 * cc_kernel_info_t.out_floats floats, block r = rank r's output. A contraction plan (kind 2, what the split / broadcast / sum matmul
 * compiles to) with `gathered` from cc_comm_symmetric_alloc, peer mailboxes mapped and N % 4 == 0 runs the exchange INSIDE the
 * tensor-core kernel's epilogue (cc_matmul_3xtf32_allgather); everything else launches into a temporary and all-gathers it.
 * `*out_fused` (may be NULL) reports which route ran. */
int cc_shard_launch_allgather(cc_kernel k, const cc_buffer* args, int n_args, cc_buffer gathered, const cc_event* waits, int n_waits,
                              cc_event* out_event, int* out_fused);

/* ---- host-side mirror of the Tensor API (flat C view of compute::cuda::Tensor, see tensor.h) ------------------- */
/* These build the same lazy graphs as T:395-1442 and evaluate them through the cc_* functions above. */

enum { CT_EXP = 10, CT_LOG = 11, CT_ABS = 12, CT_TANH = 13, CT_SQRT = 14, CT_NEG = 15 };
enum { CT_MIN = 20, CT_MAX = 21, CT_PLUS = 22, CT_MINUS = 23, CT_TIMES = 24, CT_DIV = 25, CT_PERCENT = 26 };

int ct_from_host(const float* data, const int32_t* shape, int rank, float padding, ct_tensor* out); /* Tensor.apply T:445-463 */
int ct_from_buffer(cc_buffer buf, const int32_t* shape, int rank, float padding, ct_tensor* out);
int ct_scalar(float value, float padding, ct_tensor* out);                                         /* T:465-467 */
int ct_fill(float value, const int32_t* shape, int rank, float padding, ct_tensor* out);           /* T:469-477 */
int ct_random(const int32_t* shape, int rank, int32_t seed, float padding, ct_tensor* out);        /* T:479-497 */
int ct_random_normal(const int32_t* shape, int rank, int32_t seed, float padding, ct_tensor* out); /* T:500-524 */
int ct_unary(int op, ct_tensor t, ct_tensor* out);                                                 /* T:526-544, 893-895 */
int ct_binary(int op, ct_tensor lhs, ct_tensor rhs, ct_tensor* out);                               /* T:546-558, 905-945 */
int ct_broadcast(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out);                     /* T:816-855 */
int ct_reshape(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out);                       /* T:879-888 */
int ct_scale(ct_tensor t, const int32_t* shape, int rank, ct_tensor* out);                         /* T:950-965 */
int ct_translate(ct_tensor t, const double* offset, int n_offset, const int32_t* new_shape, int new_rank,
                 ct_tensor* out);                                                                  /* T:970-976 */
int ct_permute(ct_tensor t, const int32_t* dimensions, int n, ct_tensor* out);                     /* T:1008-1025 */
int ct_transpose(ct_tensor t, ct_tensor* out);                                                     /* T:1030 */
/* `out` receives shape[dimension] tensors (T:1035-1074) */
int ct_split(ct_tensor t, int dimension, ct_tensor* out, int capacity, int* out_count);
int ct_join(const ct_tensor* tensors, int n, ct_tensor* out);                                      /* T:577-598 */
int ct_join_dim(const ct_tensor* tensors, int n, int dimension, ct_tensor* out);                   /* T:560-575 */
int ct_sum(ct_tensor t, ct_tensor* out);                                                           /* T:771 */
/* reduce(MonoidPrograms) (T:308-311, 673-766) with monoid = CT_PLUS / CT_MIN / CT_MAX / CT_TIMES. An inline operand's closure
 * is fused into the fold (one pass, nothing materialised); ct_sum(t) == ct_reduce(t, CT_PLUS). */
int ct_reduce(ct_tensor t, int monoid, ct_tensor* out);
int ct_non_inline(ct_tensor t, ct_tensor* out);                                                    /* T:671, 1405-1410 */
int ct_do_cache(ct_tensor t, ct_tensor* out);                                                      /* T:642-666 */
int ct_rank(ct_tensor t, int* out);
int ct_shape(ct_tensor t, int32_t* out, int capacity);
int ct_padding(ct_tensor t, float* out);
/* slow actions (T:1099-1118, 776-811) — evaluate, read back, block */
int ct_flat_array(ct_tensor t, float* host_out, uint64_t capacity_floats);
/* flatBuffer (T:1099-1109): evaluate and read back into callee-allocated PINNED host memory (the reference hands out an
 * LWJGL-malloc'd FloatBuffer valid inside the Do scope, O:691-715); the caller ends the scope with ct_flat_buffer_release.
 * D2H runs at PCIe speed with no staging copy, unlike ct_flat_array into pageable memory. */
int ct_flat_buffer(ct_tensor t, float** out_host, uint64_t* out_n_floats);
int ct_flat_buffer_release(float* host);
int ct_to_string(ct_tensor t, char* out, uint64_t capacity, uint64_t* out_needed);
/* evaluate and keep on the device: doBuffer (T:1401-1403); `*out_event` may be 0 when already complete */
int ct_do_buffer(ct_tensor t, cc_buffer* out, cc_event* out_event);
/* the kernel the tensor's closure compiles to, without running it (for tests of cache / pattern behaviour) */
int ct_compile(ct_tensor t, cc_kernel* out);
/* the tree blob ct_compile hands to cc_compile_ex for this tensor (definitions attached) — introspection: tests pin the blob a JVM
 * front end must write (scala/.../CudaTreeWriter.scala) against it. Call with out = NULL to size (*out_needed). */
int ct_tree_blob(ct_tensor t, void* out, uint64_t capacity, uint64_t* out_needed);
/* Sharded tensors. `ct_shard(local)` declares `local` ([rows on this rank, ...]) to be this rank's row block of a tensor sharded along
 * its leading axis; the property follows the tensor through the lazy graph (ct_distribution: 0 = whole / replicated, 1 = row block,
 * 2 = partial sum):
 *   - elementwise operators, broadcast / permute / translate / split(d > 0) / reshape / join that keep the leading axis in place, and
 *     operations with replicated operands keep a row block a row block — no exchange, each rank runs its own kernel, and the
 *     split / broadcast / sum matmul of a row block of A with a replicated B is still ONE tcgen05 contraction per rank;
 *   - sum / reduce(+) of a row block is the GLOBAL sum: local fold + all-reduce of one float (one fused kernel over NVLink peer memory);
 *   - split(0) of a row block yields the LOCAL rows as partial contributions: folding them with + (`t.split(0).reduce(_ + _)`, the column
 *     sums) gives a partial sum, which is all-reduced when it is evaluated or used by anything that is not + of partial sums;
 *   - a view that would mix the sharded axis (permute moving dimension 0, translate along it, broadcast over it) is CC_ERR_UNSUPPORTED:
 *     gather or replicate first.
 * `ct_gather(t, zero_copy)` is the whole tensor on every rank ([n_ranks * rows, ...]; needs equal blocks): a sharded matmul result is
 * gathered by the contraction's own epilogue. zero_copy != 0 returns a view of the communicator's symmetric arena, valid until the next
 * gather of the same size; 0 copies it out. With no communicator (one GPU) all of this degenerates to the identity. */
int ct_shard(ct_tensor local, ct_tensor* out);
int ct_distribution(ct_tensor t, int* out);
int ct_gather(ct_tensor t, int zero_copy, ct_tensor* out);
int ct_release(ct_tensor t);
int ct_live_tensors(int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* COMPUTE_CUDA_H */
