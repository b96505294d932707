// runtime.cpp — the cc_* half of the C ABI: context, stream pool, events, pooled device memory, pinned staging, the
// structural kernel cache with NVRTC (sm_100a), launches, NCCL. Replaces trait OpenCL (OpenCL.scala:1139-1433) and the
// launcher half of Tensors.scala (:1263-1392); see include/compute_cuda.h for the per-function citations.
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>

#include <nvtx3/nvToolsExt.h>
#include <sys/stat.h>
#include <unistd.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <cstddef>

#include <time.h>

#include <algorithm>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "builtin_kernels.h"
#include "codegen.h"
#include "driver.h"
#include "ir.h"
#include "jit_templates_embed.h"

namespace cc {
namespace {

struct Event {
  CUevent ev = nullptr;
  std::atomic<int> rc{1};
};

// A point on a stream's timeline: "the first `seq` commands submitted to stream `stream`". Hazards between commands
// on different streams are resolved lazily: only when a consumer on another stream shows up is an event recorded on
// the producer's stream (which then covers the producer command and everything before it). Commands on the same
// stream need nothing, so the common single-stream launch path makes no event / wait driver calls at all.
struct Mark {
  int stream = -1;
  uint64_t seq = 0;
};

struct Buffer {
  CUdeviceptr ptr = 0;
  uint64_t n_floats = 0;
  size_t bytes = 0;  // pooled block size (0 for wrapped memory)
  bool owned = true;
  std::atomic<int> rc{1};
  Mark last_write;
  SmallVec<Mark, 4> reads;  // at most one per stream
  std::vector<CUdeviceptr> peers;  // symmetric buffers only: the same allocation on every rank (own pointer at own rank)
  // multicast symmetric buffers (NVLS): `ptr` is this rank's own memory, `mc_ptr` a second mapping through which ONE store lands in the
  // same offset of every rank's memory (replicated by the NVSwitch); created with the virtual-memory-management API, not the pool
  CUdeviceptr mc_ptr = 0;
  CUmemGenericAllocationHandle mc_handle = 0, mem_handle = 0;
  size_t vmm_bytes = 0;
  uint64_t uid = 0;         // never reused: identity for the operand-panel cache
  uint64_t version = 0;     // bumped by every command that writes the buffer
};

struct Block {
  CUdeviceptr ptr;
  size_t bytes;
  SmallVec<Mark, 4> pending;
};

struct Kernel {
  Plan plan;
  std::string full_source;
  std::vector<char> cubin;
  CUmodule mod = nullptr;
  std::vector<CUfunction> fns;
  bool loaded = false;
  bool pdl = false;  // every generated entry point starts with cc_pdl_entry(): launched with programmatic stream serialisation
  std::atomic<int> rc{1};
  uint64_t hash = 0;
  int last_hit = 0;
  uint64_t last_use = 0;  // kernel-cache clock (LRU eviction when a limit is set)
};

struct Nccl {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  ncclComm_t comm = nullptr;
  int n_ranks = 0, rank = 0;
  // NVLink peer mailboxes (CUDA IPC)
  bool peer_enabled = false;
  bool peer_mapped = false;
  CUdeviceptr mailbox = 0;
  CUdeviceptr peer_base[kPeerMaxRanks] = {0};
  PeerMailboxes mb{};
  CUdeviceptr mb_dev = 0;  // device copy of `mb`: what a generated reduction that completes its collective itself is handed (ARG_PEER_MB)
  unsigned epoch = 0;
  std::vector<Buffer*> symmetric;  // cc_comm_symmetric_alloc results, freed when the communicator goes
  void load() {
    if (lib) return;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) fail(CC_ERR_NCCL, strprintf("cannot load libnccl.so.2: %s", dlerror()));
#define CC_NCCL(sym)                                                        \
  sym = (decltype(sym))dlsym(lib, "nccl" #sym);                             \
  if (!sym) fail(CC_ERR_NCCL, "libnccl.so.2 lacks nccl" #sym);
    CC_NCCL(GetUniqueId) CC_NCCL(CommInitRank) CC_NCCL(CommDestroy) CC_NCCL(AllReduce) CC_NCCL(AllGather) CC_NCCL(Broadcast)
        CC_NCCL(GetErrorString)
#undef CC_NCCL
  }
  void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) fail(CC_ERR_NCCL, strprintf("%s failed: %s", what, GetErrorString ? GetErrorString(r) : "?"));
  }
};

struct Runtime {
  std::recursive_mutex mu;
  bool initialized = false;
  int ordinal = 0;
  CUdevice dev = 0;
  CUcontext ctx = nullptr;
  cc_device_info_t info{};
  std::vector<CUstream> streams;       // compute streams [0, stream_count), then h2d, d2h, aux
  std::vector<uint64_t> seq;           // commands submitted per stream
  std::vector<std::vector<uint64_t>> synced;  // synced[s][a]: stream s already waits for the first synced[s][a] commands of a
  std::vector<std::array<uint64_t, 8>> submit_ns;  // host time at which each stream's last 8 commands were submitted (stream affinity)
  size_t next_stream = 0;
  int stream_count = 4;
  int h2d = 0, d2h = 0, aux = 0;       // indices into `streams`
  std::vector<CUevent> event_pool;
  // pinned host blocks (size classes: powers of two from 4 KiB): cuMemHostAlloc costs ~0.3 ms per MiB, a read-back must not pay it
  std::map<size_t, std::vector<void*>> host_pool;     // idle blocks per class
  std::unordered_map<void*, size_t> host_blocks;      // every live block -> its class
  size_t host_bytes_idle = 0;
  std::map<size_t, std::vector<Block>> pool;
  size_t bytes_pooled = 0, bytes_in_use = 0;
  std::unordered_set<Buffer*> buffers;
  std::unordered_set<Event*> events;
  std::unordered_set<Kernel*> kernels;
  std::unordered_map<std::string, Kernel*> cache;
  struct InFlight {  // a structure some thread is compiling right now (outside the lock)
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
  };
  std::unordered_map<std::string, std::shared_ptr<InFlight>> compiling;
  uint64_t cache_clock = 0;
  uint64_t cache_limit = 0;  // 0 = unbounded (the reference's default kernelCacheBuilder, Tensors.scala:1267-1277)
  cc_stats_t stats{};
  // built-in command profiler (cc_profile_*): one timing-event pair per command while enabled
  struct ProfRecord {
    std::string label;
    CUevent start, stop;
    uint64_t bytes, flops;
  };
  bool profiling = false;
  bool nvtx = false;  // CC_NVTX=1: every command is an NVTX range named like the profiler's rows (ncu --nvtx --nvtx-include ...)
  std::vector<ProfRecord> prof_records;
  std::vector<CUevent> prof_event_pool;
  // B^T hi / lo panels of recent contractions, kept while the B buffer is unchanged (same uid and write version)
  struct Panels {
    uint64_t uid, version;
    int64_t k, n;
    Buffer* hi;
    Buffer* lo;
    uint64_t last_use;
  };
  std::vector<Panels> panel_cache;
  bool panel_cache_on = true;
  uint64_t panel_clock = 0, next_uid = 1;
  // builtin scratch
  Buffer* reduce_scratch = nullptr;
  CUdeviceptr reduce_counter = 0;
  CUdeviceptr col_counters = 0;  // kColCounters self-resetting block counters per compute stream (fused axis-reduction second stage)
  CUevent timer0 = nullptr, timer1 = nullptr;
  Nccl nccl;
  // cc_graph_begin .. cc_graph_end: commands are captured on compute stream 0 instead of executed
  struct Graph {
    CUgraph graph = nullptr;
    CUgraphExec exec = nullptr;
    std::vector<Buffer*> reads, writes;  // every buffer the captured commands touch (retained: their addresses are baked into the graph)
    std::vector<Block> blocks;           // device blocks released while capturing: theirs too, so they stay out of the pool
    uint64_t commands = 0;
  };
  Graph* capture = nullptr;
  std::unordered_set<Graph*> graphs;
};

Runtime& rt() {
  static Runtime* r = new Runtime();  // intentionally leaked: survives static destruction order
  return *r;
}

struct Lock {
  std::lock_guard<std::recursive_mutex> g;
  Lock() : g(rt().mu) {}
};

void require_init() {
  CC_REQUIRE(rt().initialized, CC_ERR_NOT_INITIALIZED, "cc_init has not been called (or failed): no CUDA context — there is no CPU fallback");
  CC_CU(cuCtxSetCurrent(rt().ctx));
}

// ---- events ---------------------------------------------------------------------------------------------------------

Event* new_event() {
  Runtime& r = rt();
  Event* e = new Event();
  if (!r.event_pool.empty()) {
    e->ev = r.event_pool.back();
    r.event_pool.pop_back();
  } else {
    CC_CU(cuEventCreate(&e->ev, CU_EVENT_DISABLE_TIMING));
  }
  r.events.insert(e);
  return e;
}
void retain(Event* e) { e->rc.fetch_add(1); }
void release(Event* e) {
  if (e->rc.fetch_sub(1) == 1) {
    Runtime& r = rt();
    r.events.erase(e);
    if (r.initialized)
      r.event_pool.push_back(e->ev);
    delete e;
  }
}
Event* as_event(cc_event h) {
  Event* e = (Event*)(uintptr_t)h;
  CC_REQUIRE(e && rt().events.count(e), CC_ERR_ILLEGAL_ARGUMENT, "invalid event handle");
  return e;
}
Buffer* as_buffer(cc_buffer h) {
  Buffer* b = (Buffer*)(uintptr_t)h;
  CC_REQUIRE(b && rt().buffers.count(b), CC_ERR_ILLEGAL_ARGUMENT, "invalid buffer handle");
  return b;
}
Kernel* as_kernel(cc_kernel h) {
  Kernel* k = (Kernel*)(uintptr_t)h;
  CC_REQUIRE(k && rt().kernels.count(k), CC_ERR_ILLEGAL_ARGUMENT, "invalid kernel handle");
  return k;
}

// ---- streams & hazards ---------------------------------------------------------------------------------------------------

int pick_stream() {
  Runtime& r = rt();
  if (r.capture) return 0;  // a captured sequence lives on one stream
  int s = (int)(r.next_stream % (size_t)r.stream_count);
  r.next_stream++;
  return s;
}

using BufferList = SmallVec<Buffer*, 8>;

// Stream choice with affinity. Round-robin over the compute streams (the reference's "5 queues per device", cpu.scala:115) lets
// independent commands overlap, but a command whose buffers were JUST touched on one compute stream gains nothing from another: the
// hazard tracker would serialise it behind that stream with an event record + wait (two driver calls), and kernels on different
// streams cannot use programmatic dependent launch. The typical case is the steady state of a loop: the output block comes back from
// the pool carrying the mark of the previous iteration's kernel. So: run on the compute stream that needs the fewest cross-stream
// waits for the HOT hazards of these buffers -- marks among the last 8 commands of a compute stream, submitted a few milliseconds ago at most.
// Older marks (inputs uploaded or computed long ago) cost a stream one wait ever and must not pin independent work to one stream;
// marks on the copy streams cannot be avoided by any choice. Ties rotate as before.
// a cheap monotonic tick for "was this submitted a moment ago": the TSC where there is one (1-4 ticks per ns), else nanoseconds
uint64_t now_ns() {
#if defined(__x86_64__)
  return __builtin_ia32_rdtsc();
#else
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
#endif
}
constexpr uint64_t kHotHazardTicks = 6000000ull;  // ~2-6 ms of TSC ticks (or 6 ms of nanoseconds)
int pick_stream_for(const BufferList& reads, const BufferList& writes) {
  Runtime& r = rt();
  const int n = r.stream_count;
  if (n <= 1 || n > 32 || r.capture) return pick_stream();
  int cost[32] = {0};
  uint64_t now = 0;
  auto add = [&](const Mark& m) {
    if (m.stream < 0 || m.stream >= n || r.seq[(size_t)m.stream] - m.seq >= 8) return;
    if (!now) now = now_ns();
    if (now - r.submit_ns[(size_t)m.stream][m.seq & 7] > kHotHazardTicks) return;
    for (int s = 0; s < n; ++s)
      if (m.stream != s && r.synced[(size_t)s][(size_t)m.stream] < m.seq) cost[s]++;
  };
  for (const Buffer* b : reads) add(b->last_write);
  for (const Buffer* b : writes) {
    add(b->last_write);
    for (const Mark& m : b->reads) add(m);
  }
  const int start = pick_stream();
  int best = start;
  for (int k = 1; k < n; ++k) {
    const int s = (start + k) % n;
    if (cost[s] < cost[best]) best = s;
  }
  return best;
}

struct Op {
  int stream;
  BufferList reads, writes;
  // what the profiler files this command under, and the algorithmic work it does (0 = unknown)
  std::string label = "command";
  uint64_t bytes = 0, flops = 0;
  CUevent prof_start = nullptr;
  CUstream cu() const { return rt().streams[(size_t)stream]; }
};

CUevent prof_event() {
  Runtime& r = rt();
  CUevent ev;
  if (!r.prof_event_pool.empty()) {
    ev = r.prof_event_pool.back();
    r.prof_event_pool.pop_back();
  } else {
    CC_CU(cuEventCreate(&ev, CU_EVENT_DEFAULT));
  }
  return ev;
}

void need(int s, const Mark& m) {
  Runtime& r = rt();
  if (m.stream < 0 || m.stream == s || r.synced[(size_t)s][(size_t)m.stream] >= m.seq) return;
  CUevent ev;
  if (!r.event_pool.empty()) {
    ev = r.event_pool.back();
    r.event_pool.pop_back();
  } else {
    CC_CU(cuEventCreate(&ev, CU_EVENT_DISABLE_TIMING));
  }
  CC_CU(cuEventRecord(ev, r.streams[(size_t)m.stream]));
  CC_CU(cuStreamWaitEvent(r.streams[(size_t)s], ev, 0));
  r.synced[(size_t)s][(size_t)m.stream] = r.seq[(size_t)m.stream];
  r.event_pool.push_back(ev);  // the wait has captured this record; the CUevent can be re-recorded
}

void op_begin(Op& op, const cc_event* waits, int n_waits) {
  if (Runtime::Graph* g = rt().capture) {
    // Only kernels on the capture stream can be part of a graph; copies and collectives run on their own streams / through NCCL.
    CC_REQUIRE(op.stream == 0, CC_ERR_UNSUPPORTED, "%s cannot be captured into a graph (only kernel launches can): end the capture first", op.label.c_str());
    CC_REQUIRE(n_waits == 0, CC_ERR_UNSUPPORTED, "captured commands take no wait lists: they run in capture order");
    // (not retained yet: a buffer that dies during the capture hands its block to the capture's pool, where later captured commands may
    // reuse it — the steady state of a loop; release() drops it from these lists. cc_graph_end retains what is still alive.)
    auto note = [](std::vector<Buffer*>& list, Buffer* b) {
      for (Buffer* x : list)
        if (x == b) return;
      list.push_back(b);
    };
    for (Buffer* b : op.reads) note(g->reads, b);
    for (Buffer* b : op.writes) note(g->writes, b);
    return;  // everything before the capture has completed (cc_graph_begin synchronises); inside it, stream order is the order
  }
  for (int i = 0; i < n_waits; ++i)
    if (waits[i]) CC_CU(cuStreamWaitEvent(op.cu(), as_event(waits[i])->ev, 0));
  for (Buffer* b : op.reads) need(op.stream, b->last_write);  // read after write
  for (Buffer* b : op.writes) {
    need(op.stream, b->last_write);                       // write after write
    for (const Mark& m : b->reads) need(op.stream, m);    // write after read (also covers pooled-memory reuse)
  }
  if (rt().nvtx) nvtxRangePushA(op.label.c_str());
  // after the waits: the interval measures the command, not its dependencies. (At most 2^18 unreported records are kept: a caller
  // that never asks for the report must not grow an event list without bound.)
  if (rt().profiling && rt().prof_records.size() < ((size_t)1 << 18)) {
    op.prof_start = prof_event();
    CC_CU(cuEventRecord(op.prof_start, op.cu()));
  }
}

void op_end(Op& op, cc_event* out_event) {
  Runtime& r = rt();
  if (r.capture) {
    // nothing ran: no marks to leave, and no event to hand out (0 = nothing to wait for; the work happens at cc_graph_launch)
    for (Buffer* b : op.writes) b->version++;
    if (out_event) *out_event = 0;
    return;
  }
  if (r.nvtx) nvtxRangePop();
  if (op.prof_start) {
    CUevent stop = prof_event();
    CC_CU(cuEventRecord(stop, op.cu()));
    r.prof_records.push_back(Runtime::ProfRecord{op.label, op.prof_start, stop, op.bytes, op.flops});
    op.prof_start = nullptr;
  }
  const uint64_t q = ++r.seq[(size_t)op.stream];
  if (op.stream < r.stream_count) r.submit_ns[(size_t)op.stream][q & 7] = now_ns();
  for (Buffer* b : op.writes) {
    b->reads.clear();
    b->last_write = Mark{op.stream, q};
    b->version++;
  }
  for (Buffer* b : op.reads) {
    bool also_written = false;
    for (Buffer* w : op.writes) also_written |= (w == b);
    if (also_written) continue;
    bool found = false;
    for (Mark& m : b->reads)
      if (m.stream == op.stream) {
        m.seq = q;
        found = true;
      }
    if (!found) b->reads.push_back(Mark{op.stream, q});
  }
  if (out_event) {
    Event* e = new_event();
    CC_CU(cuEventRecord(e->ev, op.cu()));
    *out_event = (cc_event)(uintptr_t)e;  // the creation reference goes to the caller
  }
}

// A command that failed after part of it was queued (launch i > 0 of a multi-launch plan, an event record): the kernels already on the
// stream still write its outputs and scratch buffers, so those carry the write mark of a command that "ran" before they can go back to the
// pool or be read by anyone else; the NVTX range opened by op_begin is closed.
void op_fail(Op& op) noexcept {
  Runtime& r = rt();
  if (r.capture) return;
  if (r.nvtx) nvtxRangePop();
  if (op.prof_start) {
    r.prof_event_pool.push_back(op.prof_start);
    op.prof_start = nullptr;
  }
  const uint64_t q = ++r.seq[(size_t)op.stream];
  for (Buffer* b : op.writes) {
    b->reads.clear();
    b->last_write = Mark{op.stream, q};
    b->version++;
  }
}

// ---- memory pool ----------------------------------------------------------------------------------------------------------

size_t size_class(size_t bytes) {
  if (bytes < 512) bytes = 512;
  if (bytes <= (1u << 20)) {
    size_t p = 512;
    while (p < bytes) p <<= 1;
    return p;
  }
  const size_t g = 2u << 20;
  return (bytes + g - 1) / g * g;
}

void trim_pool() {
  Runtime& r = rt();
  for (auto& kv : r.pool)
    for (Block& b : kv.second) driver().cuMemFree(b.ptr);
  r.pool.clear();
  r.bytes_pooled = 0;
}

Buffer* alloc_buffer(uint64_t n_floats) {
  Runtime& r = rt();
  size_t bytes = size_class((size_t)n_floats * 4);
  Buffer* b = new Buffer();
  b->n_floats = n_floats;
  b->bytes = bytes;
  b->uid = r.next_uid++;
  r.stats.alloc_calls++;
  auto it = r.pool.find(bytes);
  if (it != r.pool.end() && !it->second.empty()) {
    Block blk = std::move(it->second.back());
    it->second.pop_back();
    b->ptr = blk.ptr;
    b->reads = std::move(blk.pending);
    r.bytes_pooled -= bytes;
    r.stats.pool_hits++;
  } else {
    CUresult res = driver().cuMemAlloc(&b->ptr, bytes);
    if (res == CUDA_ERROR_OUT_OF_MEMORY) {
      trim_pool();
      res = driver().cuMemAlloc(&b->ptr, bytes);
    }
    if (res != CUDA_SUCCESS) {
      delete b;
      check_cu(res, "cuMemAlloc");
    }
  }
  r.bytes_in_use += bytes;
  r.buffers.insert(b);
  return b;
}

void release(Buffer* b);
void drop_panels_of(uint64_t uid) {
  Runtime& r = rt();
  for (size_t i = 0; i < r.panel_cache.size();) {
    if (uid == 0 || r.panel_cache[i].uid == uid) {
      Runtime::Panels p = r.panel_cache[i];
      r.panel_cache.erase(r.panel_cache.begin() + (long)i);
      release(p.hi);
      release(p.lo);
    } else {
      ++i;
    }
  }
}

void release(Buffer* b) {
  if (b->rc.fetch_sub(1) != 1) return;
  Runtime& r = rt();
  r.buffers.erase(b);
  if (r.capture)
    for (std::vector<Buffer*>* list : {&r.capture->reads, &r.capture->writes})
      list->erase(std::remove(list->begin(), list->end(), b), list->end());
  if (b->owned && !r.panel_cache.empty()) drop_panels_of(b->uid);
  if (b->owned && r.initialized) {
    Block blk{b->ptr, b->bytes, std::move(b->reads)};
    if (b->last_write.stream >= 0) blk.pending.push_back(b->last_write);
    // (while capturing the pool is the capture's own: see cc_graph_begin)
    r.pool[b->bytes].push_back(std::move(blk));
    r.bytes_pooled += b->bytes;
    r.bytes_in_use -= b->bytes;
  }
  delete b;
}

// ---- NVRTC ----------------------------------------------------------------------------------------------------------------

// ---- on-disk cubin cache (opt-in) ------------------------------------------------------------------------------------------
// The reference's kernel cache lives and dies with the process (Guava cache, Tensors.scala:1267-1289): every run pays the JIT
// again (30-130 ms per expression structure here). With CC_KERNEL_CACHE_DIR / cc_kernel_disk_cache set, the cubin of each
// generated source is kept under <dir>/<fnv1a-64 of source + compiler identity>.cubin and NVRTC is skipped on a hit.
std::string& disk_cache_dir() {
  static std::string dir = [] {
    const char* e = getenv("CC_KERNEL_CACHE_DIR");
    std::string d = e ? e : "";
    if (!d.empty()) mkdir(d.c_str(), 0777);  // best effort; an unusable directory only means no hits
    return d;
  }();
  return dir;
}
uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}
std::string disk_cache_path(const std::string& dir, const std::string& full_source) {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  const std::string identity = strprintf("nvrtc %d.%d sm_100a fmad lineinfo extra-device-vectorization|%s|", major, minor, cc_version());
  return strprintf("%s/%016llx.cubin", dir.c_str(), (unsigned long long)fnv1a(full_source, fnv1a(identity)));
}
bool disk_cache_load(const std::string& path, const std::string& full_source, std::vector<char>& cubin) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  // layout: u64 source length, u64 source hash (guards against a path collision), cubin bytes
  uint64_t hdr[2] = {0, 0};
  bool ok = fread(hdr, 8, 2, f) == 2 && hdr[0] == full_source.size() && hdr[1] == fnv1a(full_source, 0x9e3779b97f4a7c15ull);
  if (ok) {
    fseek(f, 0, SEEK_END);
    long end = ftell(f);
    ok = end > 16;
    if (ok) {
      cubin.resize((size_t)end - 16);
      fseek(f, 16, SEEK_SET);
      ok = fread(cubin.data(), 1, cubin.size(), f) == cubin.size();
    }
  }
  fclose(f);
  if (!ok) cubin.clear();
  return ok;
}
void disk_cache_store(const std::string& path, const std::string& full_source, const std::vector<char>& cubin) {
  const std::string tmp = path + strprintf(".tmp%d", (int)getpid());  // write + rename: concurrent processes never see half a file
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;  // an unwritable cache directory costs a recompile next time, nothing else
  uint64_t hdr[2] = {full_source.size(), fnv1a(full_source, 0x9e3779b97f4a7c15ull)};
  bool ok = fwrite(hdr, 8, 2, f) == 2 && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

// Programmatic dependent launch (on by default, CC_PDL=0 turns it off): consecutive small kernels on one stream are bound by launch
// latency plus the ramp of one wave of CTAs. With the attribute set, the CTAs of the next kernel are scheduled as soon as every CTA
// of the running one has passed its `griddepcontrol.launch_dependents`, and park at `griddepcontrol.wait` until that grid has
// completed and flushed -- the same ordering as plain stream order, minus the launch gap. Both instructions are the first thing
// every generated entry point executes (cc_pdl_entry), before any global memory access, so a kernel never observes anything its
// predecessor has not finished writing. Measured (profiles/r01_pdl_ab.json): two-launch axis-0 sum of 4096^2 14.3 -> 11.2 us,
// whole-tensor fold of 1024^2 6.1 -> 4.8 us, many-wave kernels unchanged, every result bit-identical.
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CC_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}
std::string with_pdl_entries(const std::string& src) {
  std::string out;
  size_t pos = 0;
  for (;;) {
    const size_t k = src.find("extern \"C\" __global__", pos);
    if (k == std::string::npos) break;
    const size_t brace = src.find("{\n", k);
    if (brace == std::string::npos) break;
    out.append(src, pos, brace + 2 - pos);
    out += "  cc_pdl_entry();\n";
    pos = brace + 2;
  }
  out.append(src, pos, std::string::npos);
  return out;
}

// Runs WITHOUT the runtime lock (cc_compile_ex): touches only `k`; `cache_dir` is the on-disk cache directory as it was under the lock
// ("" = none). Returns true if the cubin came from the on-disk cache.
bool nvrtc_compile(Kernel& k, const std::string& cache_dir) {
  k.pdl = pdl_enabled();
  // CC_COHERENT_LOADS=1 (opt-in, for A/B): argument loads without .nc (jit_templates.cuh). Not needed for correctness: a kernel is only
  // launched with the PDL attribute when nothing it reads was written by the command right before it (cc_launch), so a grid whose
  // lifetime starts early never overlaps a writer of its inputs. (Coherent loads cost the C2 chain 3 %: 7089 -> 6880 GB/s.)
  static const bool coherent = [] {
    const char* e = getenv("CC_COHERENT_LOADS");
    return e && atoi(e) != 0;
  }();
  k.full_source = std::string("// ") + k.plan.note + "\n" + (coherent ? "#define CC_COHERENT_LOADS 1\n" : "") + kJitTemplates + "\n" +
                  (k.pdl ? with_pdl_entries(k.plan.source) : k.plan.source);
  std::string cache_path;
  if (!cache_dir.empty()) {
    cache_path = disk_cache_path(cache_dir, k.full_source);
    if (disk_cache_load(cache_path, k.full_source, k.cubin)) return true;
  }
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, k.full_source.c_str(), "jit_kernel.cu", 0, nullptr, nullptr);
  CC_REQUIRE(r == NVRTC_SUCCESS, CC_ERR_COMPILE, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
  // --minimal (NVRTC >= 12.4) leaves out the texture / runtime-API / lambda support declarations no generated kernel uses: 10-25 %
  // off the JIT time of a small kernel; older compilers reject the option and are asked again without it
  static std::atomic<bool> minimal_ok{true};
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true", "-lineinfo", "--extra-device-vectorization", "--minimal"};
  r = nvrtcCompileProgram(prog, minimal_ok ? 6 : 5, opts);
  if (r == NVRTC_ERROR_INVALID_OPTION && minimal_ok) {
    minimal_ok = false;
    r = nvrtcCompileProgram(prog, 5, opts);
  }
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    fail(CC_ERR_COMPILE, strprintf("NVRTC (sm_100a) failed: %s\n%s\n---- source ----\n%s", nvrtcGetErrorString(r), log.c_str(),
                                   k.full_source.c_str()));
  }
  size_t n = 0;
  nvrtcGetCUBINSize(prog, &n);
  k.cubin.resize(n);
  nvrtcGetCUBIN(prog, k.cubin.data());
  nvrtcDestroyProgram(&prog);
  if (!cache_path.empty()) disk_cache_store(cache_path, k.full_source, k.cubin);
  return false;
}

void ensure_loaded(Kernel& k) {
  if (k.loaded) return;
  if (!k.plan.launches.empty()) {
    // all or nothing: a failure part-way must not leave a module behind or half a function table for the retry to append to
    CUmodule mod = nullptr;
    std::vector<CUfunction> fns;
    CC_CU(cuModuleLoadData(&mod, k.cubin.data()));
    try {
      for (const LaunchSpec& ls : k.plan.launches) {
        CUfunction f;
        CC_CU(cuModuleGetFunction(&f, mod, ls.entry.c_str()));
        fns.push_back(f);
      }
    } catch (...) {
      driver().cuModuleUnload(mod);
      throw;
    }
    k.mod = mod;
    k.fns = std::move(fns);
  }
  k.loaded = true;
}

void release(Kernel* k) {
  if (k->rc.fetch_sub(1) != 1) return;
  Runtime& r = rt();
  r.kernels.erase(k);
  if (k->mod && r.initialized) {
    // launches from this module may still be queued (an evicted kernel, a handle released right after cc_launch): let them run first
    driver().cuCtxSynchronize();
    driver().cuModuleUnload(k->mod);
  }
  delete k;
}

// drop least-recently-used kernels beyond the limit (RemovalListener -> monadicClose, Tensors.scala:1267-1277); handles held by
// callers stay valid until released
void evict_kernels() {
  Runtime& r = rt();
  while (r.cache_limit && r.cache.size() > r.cache_limit) {
    auto victim = r.cache.begin();
    for (auto it = r.cache.begin(); it != r.cache.end(); ++it)
      if (it->second->last_use < victim->second->last_use) victim = it;
    Kernel* k = victim->second;
    r.cache.erase(victim);
    release(k);
  }
}

void join_all(int target) {
  Runtime& r = rt();
  for (int a = 0; a < (int)r.streams.size(); ++a) need(target, Mark{a, r.seq[(size_t)a] + 1});
}

Buffer* reduce_scratch() {
  Runtime& r = rt();
  if (!r.reduce_scratch) {
    r.reduce_scratch = alloc_buffer(reduce_sum_scratch_floats() + 64);
    CC_CU(cuMemAlloc(&r.reduce_counter, 256));
    CC_CU(cuMemsetD32Async(r.reduce_counter, 0, 64, r.streams[0]));
    CC_CU(cuStreamSynchronize(r.streams[0]));
  }
  return r.reduce_scratch;
}

CUdeviceptr col_counters_for(int stream) {
  Runtime& r = rt();
  if (!r.col_counters) {
    const size_t words = (size_t)kColCounters * (size_t)r.stream_count;
    CC_CU(cuMemAlloc(&r.col_counters, words * 4));
    CC_CU(cuMemsetD32Async(r.col_counters, 0, words, r.streams[0]));
    CC_CU(cuStreamSynchronize(r.streams[0]));
  }
  return r.col_counters + (size_t)stream * kColCounters * 4;
}

}  // namespace
}  // namespace cc

using namespace cc;

namespace {
void close_peers(Nccl& n) {
  if (!n.peer_mapped) return;
  driver().cuCtxSynchronize();
  for (Buffer* b : n.symmetric) {
    if (b->mc_ptr) {  // multicast buffer: unmap both views, unbind, drop the handles
      Driver& d = driver();
      d.cuMemUnmap(b->mc_ptr, b->vmm_bytes);
      d.cuMemAddressFree(b->mc_ptr, b->vmm_bytes);
      d.cuMulticastUnbind(b->mc_handle, rt().dev, 0, b->vmm_bytes);
      d.cuMemUnmap(b->ptr, b->vmm_bytes);
      d.cuMemAddressFree(b->ptr, b->vmm_bytes);
      d.cuMemRelease(b->mem_handle);
      d.cuMemRelease(b->mc_handle);
      b->mc_ptr = 0;
      b->ptr = 0;
      release(b);
      continue;
    }
    for (int r = 0; r < (int)b->peers.size(); ++r)
      if (r != n.rank && b->peers[(size_t)r]) driver().cuIpcCloseMemHandle(b->peers[(size_t)r]);
    if (b->ptr) driver().cuMemFree(b->ptr);
    b->ptr = 0;
    b->peers.clear();
    release(b);
  }
  n.symmetric.clear();
  for (int r = 0; r < n.n_ranks; ++r)
    if (r != n.rank && n.peer_base[r]) driver().cuIpcCloseMemHandle(n.peer_base[r]);
  if (n.mailbox) driver().cuMemFree(n.mailbox);
  n.mailbox = 0;
  if (n.mb_dev) driver().cuMemFree(n.mb_dev);
  n.mb_dev = 0;
  n.peer_enabled = false;
  n.peer_mapped = false;
}
}  // namespace

std::atomic<uint64_t> g_comm_generation{0};  // bumped by every cc_comm_init: front ends drop what they cached per communicator

extern "C" {

const char* cc_last_error(void) { return last_error_cstr(); }
const char* cc_version(void) { return "compute_cuda 0.1 (sm_100a)"; }
int cc_is_initialized(void) { return rt().initialized ? 1 : 0; }

namespace {
// CC_SEGV_BACKTRACE=1: native frames on stderr when the process dies of SIGSEGV (the box has no debugger)
void segv_backtrace(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "\n[compute_cuda] SIGSEGV, native frames:\n";
  if (write(2, msg, sizeof msg - 1) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
}  // namespace

int cc_init(int device_ordinal) {
  return guarded([&] {
    Lock lock;
    Runtime& r = rt();
    if (r.initialized) return;
    if (getenv("CC_SEGV_BACKTRACE")) signal(SIGSEGV, segv_backtrace);
    driver().load();
    CC_CU(cuInit(0));
    int count = 0;
    CC_CU(cuDeviceGetCount(&count));
    CC_REQUIRE(count > 0, CC_ERR_NO_DRIVER, "no CUDA device visible — this backend has no CPU fallback");
    if (const char* nv = getenv("CC_NVTX")) r.nvtx = atoi(nv) != 0;
    if (device_ordinal < 0) {
      const char* lr = getenv("LOCAL_RANK");
      device_ordinal = lr ? atoi(lr) % count : 0;
    }
    CC_REQUIRE(device_ordinal < count, CC_ERR_ILLEGAL_ARGUMENT, "device %d out of range (%d visible)", device_ordinal, count);
    r.ordinal = device_ordinal;
    CC_CU(cuDeviceGet(&r.dev, device_ordinal));
    CC_CU(cuDevicePrimaryCtxRetain(&r.ctx, r.dev));
    CC_CU(cuCtxSetCurrent(r.ctx));
    cc_device_info_t& di = r.info;
    memset(&di, 0, sizeof di);
    di.ordinal = device_ordinal;
    auto attr = [&](CUdevice_attribute a) {
      int v = 0;
      CC_CU(cuDeviceGetAttribute(&v, a, r.dev));
      return v;
    };
    di.sm_count = attr(CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT);
    di.cc_major = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR);
    di.cc_minor = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR);
    di.max_smem_per_block = attr(CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN);
    di.l2_bytes = attr(CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE);
    di.sm_clock_khz = attr(CU_DEVICE_ATTRIBUTE_CLOCK_RATE);
    di.mem_clock_khz = attr(CU_DEVICE_ATTRIBUTE_MEMORY_CLOCK_RATE);
    size_t total = 0;
    CC_CU(cuDeviceTotalMem(&total, r.dev));
    di.total_mem = (int64_t)total;
    CC_CU(cuDeviceGetName(di.name, sizeof di.name, r.dev));
    if (di.cc_major != 10) {
      driver().cuDevicePrimaryCtxRelease(r.dev);
      fail(CC_ERR_UNSUPPORTED, strprintf("device %d (%s) is sm_%d%d; this backend only generates sm_100a code", device_ordinal, di.name,
                                         di.cc_major, di.cc_minor));
    }
    // copies always get their own streams so H2D / compute / D2H of independent chunks overlap
    for (int i = 0; i < r.stream_count + 3; ++i) {
      CUstream s;
      CC_CU(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING));
      r.streams.push_back(s);
    }
    r.h2d = r.stream_count;
    r.d2h = r.stream_count + 1;
    r.aux = r.stream_count + 2;
    r.seq.assign(r.streams.size(), 0);
    r.submit_ns.assign(r.streams.size(), std::array<uint64_t, 8>{});
    r.synced.assign(r.streams.size(), std::vector<uint64_t>(r.streams.size(), 0));
    CC_CU(cuEventCreate(&r.timer0, CU_EVENT_DEFAULT));
    CC_CU(cuEventCreate(&r.timer1, CU_EVENT_DEFAULT));
    r.initialized = true;
  });
}

int cc_set_stream_count(int n) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(n >= 1 && n <= 32, CC_ERR_ILLEGAL_ARGUMENT, "stream count must be in [1, 32]");
    CC_REQUIRE(!rt().initialized, CC_ERR_ILLEGAL_ARGUMENT, "cc_set_stream_count must be called before cc_init");
    rt().stream_count = n;
  });
}

int cc_shutdown(void) {
  return guarded([&] {
    Lock lock;
    Runtime& r = rt();
    if (!r.initialized) return;
    CC_CU(cuCtxSetCurrent(r.ctx));
    driver().cuCtxSynchronize();
    close_peers(r.nccl);
    if (r.nccl.comm) {
      r.nccl.CommDestroy(r.nccl.comm);
      r.nccl.comm = nullptr;
    }
    drop_panels_of(0);
    for (auto& kv : r.cache) release(kv.second);
    r.cache.clear();
    if (r.reduce_scratch) {
      release(r.reduce_scratch);
      r.reduce_scratch = nullptr;
      driver().cuMemFree(r.reduce_counter);
      r.reduce_counter = 0;
    }
    // (allocated by the first fused axis-reduction second stage, with or without a whole-tensor fold ever having run; sized by the
    // stream count of THIS initialisation)
    if (r.col_counters) driver().cuMemFree(r.col_counters);
    r.col_counters = 0;
    if (r.capture) {  // a capture left open: abandon it
      CUgraph dangling = nullptr;
      if (driver().cuStreamEndCapture) driver().cuStreamEndCapture(r.streams[0], &dangling);
      if (dangling && driver().cuGraphDestroy) driver().cuGraphDestroy(dangling);
      r.graphs.insert(r.capture);
      r.capture = nullptr;
    }
    for (Runtime::Graph* g : r.graphs) {
      if (g->exec && driver().cuGraphExecDestroy) driver().cuGraphExecDestroy(g->exec);
      if (g->graph && driver().cuGraphDestroy) driver().cuGraphDestroy(g->graph);
      g->exec = nullptr, g->graph = nullptr;
      for (Block& blk : g->blocks) driver().cuMemFree(blk.ptr);
      g->blocks.clear();
    }
    trim_pool();
    for (auto& kv : r.host_blocks) driver().cuMemFreeHost(kv.first);  // pinned host memory goes with the context too
    r.host_blocks.clear();
    r.host_pool.clear();
    r.host_bytes_idle = 0;
    // anything the caller still holds stays valid as a handle but its device memory is gone with the context
    for (Buffer* b : r.buffers)
      if (b->owned && b->ptr) {
        driver().cuMemFree(b->ptr);
        b->ptr = 0;
        b->owned = false;
      }
    for (Kernel* k : r.kernels)
      if (k->mod) {
        driver().cuModuleUnload(k->mod);
        k->mod = nullptr;
        k->loaded = false;
        k->fns.clear();
      }
    for (CUevent e : r.event_pool) driver().cuEventDestroy(e);
    r.event_pool.clear();
    for (auto& rec : r.prof_records) {
      driver().cuEventDestroy(rec.start);
      driver().cuEventDestroy(rec.stop);
    }
    r.prof_records.clear();
    for (CUevent e : r.prof_event_pool) driver().cuEventDestroy(e);
    r.prof_event_pool.clear();
    r.profiling = false;
    for (Event* e : r.events) {
      driver().cuEventDestroy(e->ev);
      e->ev = nullptr;
    }
    for (CUstream s : r.streams) driver().cuStreamDestroy(s);
    r.streams.clear();
    driver().cuEventDestroy(r.timer0);
    driver().cuEventDestroy(r.timer1);
    driver().cuDevicePrimaryCtxRelease(r.dev);
    r.ctx = nullptr;
    r.bytes_in_use = r.bytes_pooled = 0;
    r.initialized = false;
  });
}

int cc_device_count(int* out) {
  return guarded([&] {
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    driver().load();
    CC_CU(cuInit(0));
    CC_CU(cuDeviceGetCount(out));
  });
}

int cc_device_info(cc_device_info_t* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = rt().info;
  });
}

// ---- memory -----------------------------------------------------------------------------------------------------------------

int cc_buffer_alloc(uint64_t n_floats, cc_buffer* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = (cc_buffer)(uintptr_t)alloc_buffer(n_floats);
  });
}

int cc_buffer_upload(cc_buffer buf, const float* host, uint64_t n_floats, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Event* done = nullptr;
    {
      Lock lock;
      require_init();
      Buffer* b = as_buffer(buf);
      CC_REQUIRE(n_floats <= b->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "upload of %llu floats into a buffer of %llu",
                 (unsigned long long)n_floats, (unsigned long long)b->n_floats);
      CC_REQUIRE(host || n_floats == 0, CC_ERR_ILLEGAL_ARGUMENT, "null host pointer");
      Op op{rt().h2d, {}, {b}};
      op.label = "copy host -> device";
      op.bytes = n_floats * 4;
      op_begin(op, waits, n_waits);
      if (n_floats) CC_CU(cuMemcpyHtoDAsync(b->ptr, host, (size_t)n_floats * 4, op.cu()));
      rt().stats.h2d_bytes += n_floats * 4;
      cc_event ev = 0;
      op_end(op, &ev);
      done = (Event*)(uintptr_t)ev;
      if (out_event) {
        *out_event = ev;
        return;
      }
    }
    CC_CU(cuEventSynchronize(done->ev));
    Lock lock;
    release(done);
  });
}

int cc_buffer_from_host(const float* host, uint64_t n_floats, cc_buffer* out, cc_event* out_event) {
  cc_buffer b = 0;
  int st = cc_buffer_alloc(n_floats, &b);
  if (st != CC_OK) return st;
  st = cc_buffer_upload(b, host, n_floats, nullptr, 0, out_event);
  if (st != CC_OK) {
    std::string keep = cc_last_error();
    cc_buffer_release(b);
    set_last_error(keep);
    return st;
  }
  *out = b;
  return CC_OK;
}

int cc_buffer_wrap(uint64_t device_ptr, uint64_t n_floats, cc_buffer* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    CC_REQUIRE(device_ptr % 16 == 0, CC_ERR_ILLEGAL_ARGUMENT, "wrapped device memory must be 16-byte aligned");
    Buffer* b = new Buffer();
    b->ptr = (CUdeviceptr)device_ptr;
    b->n_floats = n_floats;
    b->owned = false;
    rt().buffers.insert(b);
    *out = (cc_buffer)(uintptr_t)b;
  });
}

int cc_buffer_retain(cc_buffer h) {
  return guarded([&] {
    Lock lock;
    as_buffer(h)->rc.fetch_add(1);
  });
}
int cc_buffer_release(cc_buffer h) {
  return guarded([&] {
    Lock lock;
    release(as_buffer(h));
  });
}
int cc_buffer_device_ptr(cc_buffer h, uint64_t* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = (uint64_t)as_buffer(h)->ptr;
  });
}
int cc_buffer_length(cc_buffer h, uint64_t* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = as_buffer(h)->n_floats;
  });
}

int cc_buffer_to_host(cc_buffer h, uint64_t offset, float* host, uint64_t n_floats, const cc_event* waits, int n_waits,
                      cc_event* out_event) {
  return guarded([&] {
    Event* done = nullptr;
    {
      Lock lock;
      require_init();
      Buffer* b = as_buffer(h);
      CC_REQUIRE(offset + n_floats <= b->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "read of [%llu, %llu) from a buffer of %llu floats",
                 (unsigned long long)offset, (unsigned long long)(offset + n_floats), (unsigned long long)b->n_floats);
      CC_REQUIRE(host || n_floats == 0, CC_ERR_ILLEGAL_ARGUMENT, "null host pointer");
      Op op{rt().d2h, {b}, {}};
      op.label = "copy device -> host";
      op.bytes = n_floats * 4;
      op_begin(op, waits, n_waits);
      if (n_floats) CC_CU(cuMemcpyDtoHAsync(host, b->ptr + offset * 4, (size_t)n_floats * 4, op.cu()));
      rt().stats.d2h_bytes += n_floats * 4;
      cc_event ev = 0;
      op_end(op, &ev);
      done = (Event*)(uintptr_t)ev;
      if (out_event) {
        *out_event = ev;
        return;
      }
    }
    CC_CU(cuEventSynchronize(done->ev));
    Lock lock;
    release(done);
  });
}

namespace {
constexpr size_t kHostPoolIdleLimit = (size_t)4 << 30;  // idle pinned bytes kept for reuse
size_t host_class(uint64_t bytes) {
  size_t c = 4096;
  while (c < bytes) c <<= 1;
  return c;
}
}  // namespace

int cc_memory_trim(void) {
  return guarded([&] {
    Lock lock;
    require_init();
    driver().cuCtxSynchronize();  // pooled blocks may still be in use by queued commands
    trim_pool();
  });
}

int cc_host_alloc(uint64_t bytes, void** out) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    Runtime& r = rt();
    const size_t cls = host_class(bytes);
    auto it = r.host_pool.find(cls);
    if (it != r.host_pool.end() && !it->second.empty()) {
      *out = it->second.back();
      it->second.pop_back();
      r.host_bytes_idle -= cls;
      return;
    }
    CC_CU(cuMemHostAlloc(out, cls, CU_MEMHOSTALLOC_PORTABLE | CU_MEMHOSTALLOC_DEVICEMAP));
    r.host_blocks[*out] = cls;
  });
}
int cc_host_device_ptr(void* host, uint64_t* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(host && out, CC_ERR_ILLEGAL_ARGUMENT, "null argument");
    CC_REQUIRE(rt().host_blocks.count(host), CC_ERR_ILLEGAL_ARGUMENT, "not a cc_host_alloc block");
    CUdeviceptr d = 0;
    CC_CU(cuMemHostGetDevicePointer(&d, host, 0));
    *out = (uint64_t)d;
  });
}
int cc_host_free(void* p) {
  return guarded([&] {
    Lock lock;
    require_init();
    if (!p) return;
    Runtime& r = rt();
    auto it = r.host_blocks.find(p);
    CC_REQUIRE(it != r.host_blocks.end(), CC_ERR_ILLEGAL_ARGUMENT, "cc_host_free of memory cc_host_alloc did not return");
    const size_t cls = it->second;
    for (void* q : r.host_pool[cls])
      CC_REQUIRE(q != p, CC_ERR_ILLEGAL_ARGUMENT, "cc_host_free called twice on the same block");
    if (r.host_bytes_idle + cls <= kHostPoolIdleLimit) {
      r.host_pool[cls].push_back(p);
      r.host_bytes_idle += cls;
    } else {
      r.host_blocks.erase(it);
      CC_CU(cuMemFreeHost(p));
    }
  });
}

// ---- events -----------------------------------------------------------------------------------------------------------------

int cc_event_retain(cc_event h) {
  return guarded([&] {
    Lock lock;
    retain(as_event(h));
  });
}
int cc_event_release(cc_event h) {
  return guarded([&] {
    Lock lock;
    release(as_event(h));
  });
}
int cc_event_wait(cc_event h) {
  return guarded([&] {
    CUevent ev;
    {
      Lock lock;
      require_init();
      ev = as_event(h)->ev;
    }
    CC_CU(cuEventSynchronize(ev));
  });
}
int cc_event_query(cc_event h, int* out_done) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(out_done, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    CUresult r = driver().cuEventQuery(as_event(h)->ev);
    if (r == CUDA_SUCCESS)
      *out_done = 1;
    else if (r == CUDA_ERROR_NOT_READY)
      *out_done = 0;
    else
      check_cu(r, "cuEventQuery");
  });
}

namespace {
struct Callback {
  cc_event_callback cb;
  void* user;
};
void CUDA_CB host_trampoline(void* p) {
  Callback* c = (Callback*)p;
  c->cb(c->user, 0);
  delete c;
}
}  // namespace

int cc_event_on_complete(cc_event h, cc_event_callback cb, void* user) {
  return guarded([&] {
    Lock lock;
    require_init();
    CC_REQUIRE(cb, CC_ERR_ILLEGAL_ARGUMENT, "null callback");
    Event* e = as_event(h);
    CUstream aux = rt().streams[(size_t)rt().aux];
    CC_CU(cuStreamWaitEvent(aux, e->ev, 0));
    CC_CU(cuLaunchHostFunc(aux, host_trampoline, new Callback{cb, user}));
    rt().seq[(size_t)rt().aux]++;
  });
}

int cc_synchronize(void) {
  return guarded([&] {
    {
      Lock lock;
      require_init();
    }
    CC_CU(cuCtxSynchronize());
  });
}

// ---- kernels ----------------------------------------------------------------------------------------------------------------

int cc_compile(const void* blob, uint64_t n_bytes, cc_kernel* out) { return cc_compile_ex(blob, n_bytes, out, nullptr, 0, nullptr); }

int cc_compile_ex(const void* blob, uint64_t n_bytes, cc_kernel* out, uint64_t* ids_out, int capacity, int* n_params_out) {
  return guarded([&] {
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    Tree t = parse_tree(blob, n_bytes);
    canonicalize(t);
    if (n_params_out) *n_params_out = (int)t.params.size();
    if (ids_out) {
      CC_REQUIRE(capacity >= (int)t.params.size(), CC_ERR_ILLEGAL_ARGUMENT, "param id capacity %d < %zu", capacity, t.params.size());
      for (size_t i = 0; i < t.params.size(); ++i) ids_out[i] = t.nodes[t.params[i]].param_id;
    }
    // The JIT (planning + NVRTC, 40-100 ms) runs OUTSIDE the runtime lock: other threads keep launching, and different structures
    // compile in parallel (the reference builds programs inside `Do` blocks on its execution context, Tensors.scala:1321-1329).
    // Threads asking for a structure that is being compiled wait for that compilation instead of starting a second one.
    Runtime& r = rt();
    std::shared_ptr<Runtime::InFlight> mine;
    DeviceProps dp;
    std::string cache_dir;
    for (;;) {
      std::shared_ptr<Runtime::InFlight> theirs;
      {
        Lock lock;
        auto it = r.cache.find(t.key);
        if (it != r.cache.end()) {
          Kernel* k = it->second;
          k->rc.fetch_add(1);
          k->last_hit = 1;
          k->last_use = ++r.cache_clock;
          r.stats.cache_hits++;
          *out = (cc_kernel)(uintptr_t)k;
          return;
        }
        auto fl = r.compiling.find(t.key);
        if (fl != r.compiling.end()) {
          theirs = fl->second;
        } else {
          mine = std::make_shared<Runtime::InFlight>();
          r.compiling.emplace(t.key, mine);
          cache_dir = disk_cache_dir();
          dp.contraction = gemm_available() && !plan_knob("CC_DISABLE_CONTRACTION");
          if (r.initialized) {
            dp.sm_count = r.info.sm_count;
            dp.max_smem = r.info.max_smem_per_block;
          }
        }
      }
      if (!theirs) break;
      std::unique_lock<std::mutex> wait(theirs->mu);
      theirs->cv.wait(wait, [&] { return theirs->done; });
      // compiled (next lookup hits) or failed (this thread compiles it itself and reports its own error)
    }
    // whatever happens from here on (planning or NVRTC failing, allocation failure), the marker goes and the waiters wake up
    struct Finisher {
      Runtime& r;
      const std::string key;  // a copy: t.key moves into the cache below
      std::shared_ptr<Runtime::InFlight> mine;
      ~Finisher() {
        Lock lock;
        r.compiling.erase(key);
        {
          std::lock_guard<std::mutex> g(mine->mu);
          mine->done = true;
        }
        mine->cv.notify_all();
      }
    } finisher{r, t.key, mine};
    std::unique_ptr<Kernel> k(new Kernel());
    bool disk_hit = false, compiled = false;
    k->plan = make_plan(t, dp);
    k->hash = t.hash;
    if (!k->plan.launches.empty()) {
      disk_hit = nvrtc_compile(*k, cache_dir);
      compiled = !disk_hit;
    } else {
      k->full_source = std::string("// ") + k->plan.note + "\n" + k->plan.source;
    }
    Lock lock;  // (recursive: the finisher takes it again on the way out, after the kernel is in the cache)
    r.stats.compiles++;
    if (disk_hit) r.stats.disk_cache_hits++;
    if (compiled) r.stats.nvrtc_compiles++;
    r.kernels.insert(k.get());
    r.cache.emplace(std::move(t.key), k.get());
    Kernel* raw = k.release();
    raw->rc.store(2);  // cache + caller
    raw->last_use = ++r.cache_clock;
    evict_kernels();
    *out = (cc_kernel)(uintptr_t)raw;
  });
}

int cc_kernel_disk_cache(const char* directory) {
  return guarded([&] {
    Lock lock;
    disk_cache_dir() = directory ? directory : "";
    while (disk_cache_dir().size() > 1 && disk_cache_dir().back() == '/') disk_cache_dir().pop_back();
    if (!disk_cache_dir().empty()) {
      struct stat st;
      if (stat(disk_cache_dir().c_str(), &st) != 0) mkdir(disk_cache_dir().c_str(), 0777);
      CC_REQUIRE(stat(disk_cache_dir().c_str(), &st) == 0 && S_ISDIR(st.st_mode), CC_ERR_ILLEGAL_ARGUMENT, "kernel cache directory %s does not exist and cannot be created",
                 disk_cache_dir().c_str());
    }
  });
}

int cc_kernel_cache_limit(uint64_t max_kernels) {
  return guarded([&] {
    Lock lock;
    rt().cache_limit = max_kernels;
    if (rt().initialized) CC_CU(cuCtxSetCurrent(rt().ctx));
    evict_kernels();
  });
}

int cc_kernel_cache_clear(void) {
  return guarded([&] {
    Lock lock;
    Runtime& r = rt();
    if (r.initialized) CC_CU(cuCtxSetCurrent(r.ctx));
    for (auto& kv : r.cache) release(kv.second);
    r.cache.clear();
    plan_knobs_refresh();  // an empty cache is the one moment a changed planning switch can take effect consistently
  });
}

int cc_kernel_cache_size(uint64_t* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = rt().cache.size();
  });
}

int cc_kernel_cache_lookup(const void* blob, uint64_t n_bytes, int any_out_shape, cc_kernel* out) {
  return guarded([&] {
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = 0;
    Tree t = parse_tree(blob, n_bytes);
    canonicalize(t);
    Lock lock;
    Runtime& r = rt();
    Kernel* found = nullptr;
    if (!any_out_shape) {
      auto it = r.cache.find(t.key);
      if (it != r.cache.end()) found = it->second;
    } else {
      // the key starts with the output shape (u32 rank, u32 dims[rank]); everything after it is the structure of the term
      auto tail = [](const std::string& key) {
        uint32_t rank = 0;
        if (key.size() >= 4) memcpy(&rank, key.data(), 4);
        const size_t skip = 4 + 4 * (size_t)rank;
        return skip <= key.size() ? std::make_pair(key.data() + skip, key.size() - skip) : std::make_pair(key.data(), (size_t)0);
      };
      const auto want = tail(t.key);
      for (auto& kv : r.cache) {
        const auto have = tail(kv.first);
        if (have.second == want.second && memcmp(have.first, want.first, want.second) == 0) {
          found = kv.second;
          break;
        }
      }
    }
    if (found) {
      found->rc.fetch_add(1);
      *out = (cc_kernel)(uintptr_t)found;
    }
  });
}

int cc_kernel_retain(cc_kernel h) {
  return guarded([&] {
    Lock lock;
    as_kernel(h)->rc.fetch_add(1);
  });
}
int cc_kernel_release(cc_kernel h) {
  return guarded([&] {
    Lock lock;
    release(as_kernel(h));
  });
}
int cc_kernel_info(cc_kernel h, cc_kernel_info_t* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    Kernel* k = as_kernel(h);
    out->kind = k->plan.kind;
    out->cache_hit = k->last_hit;
    out->n_args = (int32_t)k->plan.arg_params.size();
    out->n_launches = k->plan.kind == PLAN_CONTRACTION ? (k->plan.gathered_panels ? 1 + (int32_t)k->plan.launches.size() : 3)  // (2 when B's panels are cached)
                                                        : (int32_t)k->plan.launches.size();
    out->out_floats = k->plan.out_floats;
    out->algorithmic_bytes = k->plan.algorithmic_bytes;
    out->flops = k->plan.flops;
    out->structural_hash = k->hash;
  });
}
int cc_kernel_arg_param(cc_kernel h, int i, int32_t* out) {
  return guarded([&] {
    Lock lock;
    Kernel* k = as_kernel(h);
    CC_REQUIRE(out && i >= 0 && i < (int)k->plan.arg_params.size(), CC_ERR_ILLEGAL_ARGUMENT, "argument index %d out of range", i);
    *out = (int32_t)k->plan.arg_params[i];
  });
}
int cc_kernel_launch_info(cc_kernel h, int i, cc_launch_info_t* out) {
  return guarded([&] {
    Lock lock;
    Kernel* k = as_kernel(h);
    CC_REQUIRE(out && i >= 0 && i < (int)k->plan.launches.size(), CC_ERR_ILLEGAL_ARGUMENT, "launch index %d out of range", i);
    const LaunchSpec& ls = k->plan.launches[(size_t)i];
    memset(out, 0, sizeof *out);
    snprintf(out->entry, sizeof out->entry, "%s", ls.entry.c_str());
    for (int d = 0; d < 3; ++d) out->grid[d] = ls.grid[d], out->block[d] = ls.block[d];
    out->smem = ls.smem;
    CC_REQUIRE(ls.args.size() <= 32, CC_ERR_UNSUPPORTED, "launch has %zu arguments", ls.args.size());
    out->n_args = (int32_t)ls.args.size();
    for (size_t a = 0; a < ls.args.size(); ++a) out->args[a] = ls.args[a];
    out->n_scratch = (int32_t)k->plan.scratch_floats.size();
    for (size_t q = 0; q < k->plan.scratch_floats.size() && q < 8; ++q) out->scratch_floats[q] = k->plan.scratch_floats[q];
  });
}
int cc_kernel_source(cc_kernel h, const char** out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = as_kernel(h)->full_source.c_str();
  });
}

namespace {
// Workspace of one contraction: A panels are per call; the B^T panels are looked up in / added to the panel cache so that a B that
// has not been written since its last contraction (weights, the replicated operand of the row-sharded matmul) is split once.
struct GemmRun {
  Buffer *a_hi = nullptr, *a_lo = nullptr, *bt_hi = nullptr, *bt_lo = nullptr;
  Buffer* partials = nullptr;  // split-K partial results (small products on the tensor-memory-A kernel)
  bool b_ready = false;
  void declare(Op& op) const {
    if (a_hi) op.writes.push_back(a_hi);  // (none when the kernel splits A itself, through tensor memory)
    if (a_lo) op.writes.push_back(a_lo);
    if (partials) op.writes.push_back(partials);
    (b_ready ? op.reads : op.writes).push_back(bt_hi);
    (b_ready ? op.reads : op.writes).push_back(bt_lo);
  }
};

GemmRun gemm_prepare(Buffer* b, int64_t m, int64_t n, int64_t k, const Buffer* a = nullptr, bool gather_epilogue = false) {
  Runtime& r = rt();
  GemmRun g;
  const int64_t kp = gemm_padded_k(k);
  const bool a_panels = gemm_config_for(a ? (const float*)a->ptr : nullptr, m, n, k, r.info.sm_count, gather_epilogue) != 1024;
  if (r.panel_cache_on && b->owned)
    for (Runtime::Panels& p : r.panel_cache)
      if (p.uid == b->uid && p.version == b->version && p.k == k && p.n == n) {
        g.bt_hi = p.hi;
        g.bt_lo = p.lo;
        g.bt_hi->rc.fetch_add(1);
        g.bt_lo->rc.fetch_add(1);
        g.b_ready = true;
        p.last_use = ++r.panel_clock;
        break;
      }
  try {
    if (a_panels) {
      g.a_hi = alloc_buffer((uint64_t)(m * kp));
      g.a_lo = alloc_buffer((uint64_t)(m * kp));
    } else {
      const int splits = gemm_k_splits(m, n, k, r.info.sm_count);
      if (splits > 1) g.partials = alloc_buffer((uint64_t)splits * (uint64_t)m * (uint64_t)n);
    }
    if (!g.b_ready) {
      g.bt_hi = alloc_buffer((uint64_t)(n * kp));
      g.bt_lo = alloc_buffer((uint64_t)(n * kp));
    }
  } catch (...) {
    for (Buffer* x : {g.a_hi, g.a_lo, g.bt_hi, g.bt_lo, g.partials})
      if (x) release(x);
    throw;
  }
  return g;
}

void gemm_finish(GemmRun& g, Buffer* b, int64_t n, int64_t k, bool launched) {
  Runtime& r = rt();
  if (launched && !g.b_ready && r.panel_cache_on && b->owned) {
    constexpr size_t kMaxPanels = 4;
    while (r.panel_cache.size() >= kMaxPanels) {
      size_t lru = 0;
      for (size_t i = 1; i < r.panel_cache.size(); ++i)
        if (r.panel_cache[i].last_use < r.panel_cache[lru].last_use) lru = i;
      Runtime::Panels old = r.panel_cache[lru];
      r.panel_cache.erase(r.panel_cache.begin() + (long)lru);
      release(old.hi);
      release(old.lo);
    }
    g.bt_hi->rc.fetch_add(1);
    g.bt_lo->rc.fetch_add(1);
    r.panel_cache.push_back(Runtime::Panels{b->uid, b->version, k, n, g.bt_hi, g.bt_lo, ++r.panel_clock});
  }
  for (Buffer* x : {g.a_hi, g.a_lo, g.bt_hi, g.bt_lo, g.partials})
    if (x) release(x);
}

void gemm_on_stream(Buffer* a, Buffer* b, Buffer* c, int64_t m, int64_t n, int64_t k, const GemmRun& g, const Op& op) {
  CUstream s = op.cu();
  GemmWorkspace ws{g.a_hi ? (float*)g.a_hi->ptr : nullptr, g.a_lo ? (float*)g.a_lo->ptr : nullptr, (float*)g.bt_hi->ptr, (float*)g.bt_lo->ptr,
                   g.partials ? (float*)g.partials->ptr : nullptr};
  int launched = launch_gemm_3xtf32((const float*)a->ptr, (const float*)b->ptr, (float*)c->ptr, m, n, k, ws, rt().info.sm_count,
                                    (TensorMapEncodeFn)driver().cuTensorMapEncodeTiled, (cudaStream_t)s, g.b_ready);
  rt().stats.device_kernels += (uint64_t)launched;
}
}  // namespace

namespace {
const char* plan_kind_name(int kind) {
  switch (kind) {
    case PLAN_ELEMENTWISE: return "elementwise";
    case PLAN_AXIS_REDUCE: return "axis reduction";
    case PLAN_CONTRACTION: return "contraction 3xTF32";
    case PLAN_TILED_TRANSPOSE: return "tiled transpose";
    case PLAN_FULL_REDUCE: return "whole-tensor fold";
  }
  return "kernel";
}
// "<plan> #<structure hash>: <the generator's one-line description>"
void label_kernel_op(Op& op, const Kernel& k) {
  if (!rt().profiling && !rt().nvtx) return;
  std::string first = k.plan.source.substr(0, k.plan.source.find('\n'));
  if (first.rfind("// ", 0) == 0) first = first.substr(3);
  if (first.size() > 150) first.resize(150);
  op.label = strprintf("%s #%08x: %s", plan_kind_name(k.plan.kind), (unsigned)(k.hash & 0xffffffffu), first.c_str());
  op.bytes = k.plan.algorithmic_bytes;
  op.flops = k.plan.flops;
}
}  // namespace

namespace {
// `collective`: the plan's last launch completes its all-reduce / all-gather over the peer mailboxes itself (Plan::collective; the caller has
// checked that the route is open). Such a launch takes the next epoch and runs on stream 0, where every collective runs: epochs are consumed
// in stream order on every rank.
int launch_kernel(cc_kernel h, const cc_buffer* args, int n_args, cc_buffer out, const cc_event* waits, int n_waits, cc_event* out_event, bool collective) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Kernel* k = as_kernel(h);
    const Plan& p = k->plan;
    CC_REQUIRE(!collective || (p.collective != 0 && r.nccl.comm && r.nccl.peer_enabled && r.nccl.mb_dev), CC_ERR_ILLEGAL_ARGUMENT,
               "this kernel cannot complete a collective itself");
    const unsigned coll_epoch = collective ? ++r.nccl.epoch : 0u;
    CC_REQUIRE(n_args == (int)p.arg_params.size(), CC_ERR_ILLEGAL_ARGUMENT, "kernel expects %zu buffers, got %d", p.arg_params.size(), n_args);
    Buffer* ob = as_buffer(out);
    CC_REQUIRE(ob->n_floats >= p.out_floats, CC_ERR_ILLEGAL_ARGUMENT, "output buffer has %llu floats, kernel writes %llu",
               (unsigned long long)ob->n_floats, (unsigned long long)p.out_floats);
    BufferList in;
    for (int i = 0; i < n_args; ++i) {
      Buffer* b = as_buffer(args[i]);
      CC_REQUIRE(b->n_floats >= p.arg_min_floats[i], CC_ERR_ILLEGAL_ARGUMENT, "argument %d has %llu floats, kernel reads up to %llu", i,
                 (unsigned long long)b->n_floats, (unsigned long long)p.arg_min_floats[i]);
      CC_REQUIRE(b != ob, CC_ERR_ILLEGAL_ARGUMENT, "output buffer aliases argument %d", i);
      in.push_back(b);
    }
    ensure_loaded(*k);
    int launch_stream = 0;  // index of the stream the launches below go to (per-stream counters)
    auto launch_spec = [&](size_t li, const std::vector<Buffer*>& scratch, Buffer* shared_partials, CUstream stream) {
      const LaunchSpec& ls = p.launches[li];
      SmallVec<CUdeviceptr, 24> ptrs;
      SmallVec<void*, 24> argv;
      for (int a : ls.args) {
        if (a >= 0)
          ptrs.push_back(in[a]->ptr);
        else if (a == ARG_OUT)
          ptrs.push_back(ob->ptr);
        else if (a == ARG_REDUCE_PARTIALS)
          ptrs.push_back(shared_partials->ptr);
        else if (a == ARG_REDUCE_COUNTER)
          ptrs.push_back(r.reduce_counter);
        else if (a == ARG_COL_COUNTERS)
          ptrs.push_back(col_counters_for(launch_stream));
        else if (a == ARG_PEER_MB)
          ptrs.push_back(collective ? r.nccl.mb_dev : (CUdeviceptr)0);
        else if (a == ARG_PEER_EPOCH)
          ptrs.push_back((CUdeviceptr)coll_epoch);
        else
          ptrs.push_back(scratch[ARG_SCRATCH0 - a]->ptr);
      }
      for (CUdeviceptr& q : ptrs) argv.push_back(&q);
      // Programmatic dependent launch lets this grid become resident while the previous command on the stream still runs. Its loads are
      // non-coherent (ld.global.nc), which PTX only allows for data that is read-only during the grid's WHOLE lifetime: so the attribute
      // is set only when nothing this launch reads was written by that previous command (later launches of a multi-launch plan read the
      // scratch the launch before them wrote: plain stream order). Loops over long-lived inputs — the launch-bound case PDL is for —
      // keep it; a producer -> consumer chain pays the ~1 us launch gap and stays within the letter of the memory model.
      bool pdl_now = k->pdl && li == 0;
      if (pdl_now) {
        const uint64_t prev = r.seq[(size_t)launch_stream];
        for (const Buffer* b : in)
          if (b->last_write.stream == launch_stream && b->last_write.seq == prev && prev != 0) pdl_now = false;
      }
      if (pdl_now) {
        CUlaunchAttribute attr{};
        attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attr.value.programmaticStreamSerializationAllowed = 1;
        CUlaunchConfig cfg{};
        cfg.gridDimX = ls.grid[0], cfg.gridDimY = ls.grid[1], cfg.gridDimZ = ls.grid[2];
        cfg.blockDimX = ls.block[0], cfg.blockDimY = ls.block[1], cfg.blockDimZ = ls.block[2];
        cfg.sharedMemBytes = ls.smem;
        cfg.hStream = stream;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        CC_CU(cuLaunchKernelEx(&cfg, k->fns[li], argv.data(), nullptr));
      } else {
        CC_CU(cuLaunchKernel(k->fns[li], ls.grid[0], ls.grid[1], ls.grid[2], ls.block[0], ls.block[1], ls.block[2], ls.smem, stream, argv.data(),
                             nullptr));
      }
      r.stats.device_kernels++;
    };
    if (p.kind == PLAN_CONTRACTION && p.gathered_panels) {
      // general contraction: generated kernels gather the operand panels, the tcgen05 pipeline runs on them, then the epilogue
      std::vector<Buffer*> scratch;
      Op op{0, in, {ob}};
      bool begun = false;
      try {
        for (uint64_t n : p.scratch_floats) scratch.push_back(alloc_buffer(n));
        op.stream = pick_stream_for(in, {ob});
        launch_stream = op.stream;
        label_kernel_op(op, *k);
        for (Buffer* s : scratch) op.writes.push_back(s);
        op_begin(op, waits, n_waits);
        begun = true;
        launch_spec(0, scratch, nullptr, op.cu());
        launch_spec(1, scratch, nullptr, op.cu());
        const int64_t Kp = (p.K + 31) / 32 * 32;
        for (int64_t b = 0; b < p.batch; ++b) {  // batch b: rows [b*M, (b+1)*M) of A's panels, [b*N, (b+1)*N) of B's, block b of the result
          GemmWorkspace ws{(float*)scratch[0]->ptr + b * p.M * Kp, (float*)scratch[1]->ptr + b * p.M * Kp, (float*)scratch[2]->ptr + b * p.N * Kp,
                           (float*)scratch[3]->ptr + b * p.N * Kp};
          r.stats.device_kernels += (uint64_t)launch_gemm_3xtf32_panels((float*)ob->ptr + b * p.M * p.N, p.M, p.N, p.K, ws, r.info.sm_count,
                                                                        (TensorMapEncodeFn)driver().cuTensorMapEncodeTiled, (cudaStream_t)op.cu());
        }
        if (p.launches.size() > 2) launch_spec(2, scratch, nullptr, op.cu());
        r.stats.launches++;
        op_end(op, out_event);
      } catch (...) {
        if (begun) op_fail(op);  // whatever was queued still writes `out` and the scratch: they carry its mark into the pool
        for (Buffer* s : scratch) release(s);
        throw;
      }
      for (Buffer* s : scratch) release(s);
      return;
    }
    if (p.kind == PLAN_CONTRACTION) {
      GemmRun g = gemm_prepare(in[1], p.M, p.N, p.K, in[0]);
      bool launched = false, begun = false;
      Op op{pick_stream_for(in, {ob}), in, {ob}};
      try {
        label_kernel_op(op, *k);
        g.declare(op);
        op_begin(op, waits, n_waits);
        begun = true;
        gemm_on_stream(in[0], in[1], ob, p.M, p.N, p.K, g, op);
        launched = true;
        r.stats.launches++;
        op_end(op, out_event);
      } catch (...) {
        if (begun) op_fail(op);
        gemm_finish(g, in[1], p.N, p.K, false);  // (never cache panels of a failed run)
        throw;
      }
      gemm_finish(g, in[1], p.N, p.K, true);
      return;
    }
    std::vector<Buffer*> scratch;
    for (uint64_t n : p.scratch_floats) scratch.push_back(alloc_buffer(n));
    // whole-tensor folds share the runtime's partials buffer and block counter: serialised on stream 0 like cc_reduce_sum
    Buffer* shared_partials = p.kind == PLAN_FULL_REDUCE ? reduce_scratch() : nullptr;
    Op op{(shared_partials || collective) ? 0 : pick_stream_for(in, {ob}), in, {ob}};
    label_kernel_op(op, *k);
    if (collective && (r.profiling || r.nvtx)) op.label += p.collective == 1 ? " + all-reduce over NVLink (same kernel)" : " + all-gather over NVLink (same kernel)";
    for (Buffer* s : scratch) op.writes.push_back(s);
    if (shared_partials) op.writes.push_back(shared_partials);
    launch_stream = op.stream;
    bool begun = false;
    try {
      op_begin(op, waits, n_waits);
      begun = true;
      for (size_t li = 0; li < p.launches.size(); ++li) launch_spec(li, scratch, shared_partials, op.cu());
      r.stats.launches++;
      op_end(op, out_event);
    } catch (...) {
      // earlier launches of the plan may already be queued and still write `out` and the scratch buffers: they go back to the pool (and to
      // the caller) carrying the mark of this command, never unmarked
      if (begun) op_fail(op);
      for (Buffer* s : scratch) release(s);  // a failed launch must not strand the plan's scratch buffers
      throw;
    }
    for (Buffer* s : scratch) release(s);
  });
}
}  // namespace

int cc_launch(cc_kernel h, const cc_buffer* args, int n_args, cc_buffer out, const cc_event* waits, int n_waits, cc_event* out_event) {
  return launch_kernel(h, args, n_args, out, waits, n_waits, out_event, false);
}

int cc_reduce_sum(cc_buffer in, uint64_t n_floats, cc_buffer out, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Buffer* ib = as_buffer(in);
    Buffer* ob = as_buffer(out);
    CC_REQUIRE(n_floats <= ib->n_floats && ob->n_floats >= 1 && ib != ob, CC_ERR_ILLEGAL_ARGUMENT, "bad reduce_sum arguments");
    Buffer* sc = reduce_scratch();
    // the shared scratch + counter serialise full reductions on one stream
    Op op{0, {ib}, {ob, sc}};
    op.label = "sum (builtin reduce_sum)";
    op.bytes = n_floats * 4 + 4;
    op_begin(op, waits, n_waits);
    launch_reduce_sum((const float*)ib->ptr, n_floats, (float*)ob->ptr, (float*)sc->ptr, (unsigned*)r.reduce_counter, r.info.sm_count,
                      (cudaStream_t)op.cu());
    r.stats.launches++;
    r.stats.device_kernels++;
    op_end(op, out_event);
  });
}

int cc_random(cc_buffer out, uint64_t n_floats, int32_t seed, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Buffer* ob = as_buffer(out);
    CC_REQUIRE(n_floats <= ob->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "random: buffer too small");
    Op op{pick_stream(), {}, {ob}};
    op.label = "random (builtin)";
    op.bytes = n_floats * 4;
    op_begin(op, nullptr, 0);
    launch_random((float*)ob->ptr, n_floats, seed, (cudaStream_t)op.cu());
    rt().stats.launches++;
    rt().stats.device_kernels++;
    op_end(op, out_event);
  });
}

int cc_random_normal(cc_buffer out, uint64_t n_floats, int32_t seed, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Buffer* ob = as_buffer(out);
    CC_REQUIRE(n_floats <= ob->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "randomNormal: buffer too small");
    Op op{pick_stream(), {}, {ob}};
    op.label = "randomNormal (builtin)";
    op.bytes = n_floats * 4;
    op_begin(op, nullptr, 0);
    launch_random_normal((float*)ob->ptr, n_floats, seed, (cudaStream_t)op.cu());
    rt().stats.launches++;
    rt().stats.device_kernels++;
    op_end(op, out_event);
  });
}

int cc_matmul_3xtf32(cc_buffer a, cc_buffer b, cc_buffer c, int64_t m, int64_t n, int64_t k, const cc_event* waits, int n_waits,
                     cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Buffer* ab = as_buffer(a);
    Buffer* bb = as_buffer(b);
    Buffer* cb = as_buffer(c);
    CC_REQUIRE(m > 0 && n > 0 && k > 0 && m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31) - 32, CC_ERR_ILLEGAL_ARGUMENT,
               "cc_matmul_3xtf32: bad shape %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
    CC_REQUIRE(ab->n_floats >= (uint64_t)(m * k) && bb->n_floats >= (uint64_t)(k * n) && cb->n_floats >= (uint64_t)(m * n),
               CC_ERR_ILLEGAL_ARGUMENT, "matmul buffers too small");
    CC_REQUIRE(cb != ab && cb != bb, CC_ERR_ILLEGAL_ARGUMENT, "matmul output aliases an input");
    GemmRun g = gemm_prepare(bb, m, n, k, ab);
    bool launched = false;
    try {
      Op op{pick_stream_for({ab, bb}, {cb}), {ab, bb}, {cb}};
      op.label = "contraction 3xTF32 (cc_matmul_3xtf32)";
      op.flops = 2ull * (uint64_t)m * (uint64_t)n * (uint64_t)k;
      op.bytes = 4ull * (uint64_t)(m * k + k * n + m * n);
      g.declare(op);
      op_begin(op, waits, n_waits);
      gemm_on_stream(ab, bb, cb, m, n, k, g, op);
      launched = true;
      rt().stats.launches++;
      op_end(op, out_event);
    } catch (...) {
      gemm_finish(g, bb, n, k, launched);
      throw;
    }
    gemm_finish(g, bb, n, k, true);
  });
}

int cc_set_operand_cache(int on) {
  return guarded([&] {
    Lock lock;
    rt().panel_cache_on = on != 0;
    if (!on && rt().initialized) {
      CC_CU(cuCtxSetCurrent(rt().ctx));
      drop_panels_of(0);
    }
  });
}

// ---- stats / timing -----------------------------------------------------------------------------------------------------------

int cc_stats(cc_stats_t* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = rt().stats;
    out->bytes_in_use = rt().bytes_in_use;
    out->bytes_pooled = rt().bytes_pooled;
  });
}
int cc_stats_reset(void) {
  return guarded([&] {
    Lock lock;
    rt().stats = cc_stats_t{};
  });
}

// ---- built-in command profiler ------------------------------------------------------------------------------------------------
// The reference has no profiling at all (its queues are created without CL_QUEUE_PROFILING_ENABLE, OpenCL.scala:431-436). While
// enabled, every command is bracketed by a pair of timing events on its own stream; the report aggregates device time per
// kernel structure / copy direction / collective with the algorithmic GB/s and TFLOP/s the code generator attributes to it.
int cc_profile_enable(int on) {
  return guarded([&] {
    Lock lock;
    require_init();
    rt().profiling = on != 0;
  });
}

int cc_profile_report(char* out, uint64_t capacity, uint64_t* out_needed) {
  return guarded([&] {
    static thread_local std::string text;  // built on the sizing call, handed out on the copying call
    Lock lock;
    require_init();
    Runtime& r = rt();
    if (!out || capacity == 0 || text.empty()) {
      driver().cuCtxSynchronize();
      struct Agg {
        uint64_t count = 0, bytes = 0, flops = 0;
        double total_ms = 0, min_ms = 1e30, max_ms = 0;
      };
      std::map<std::string, Agg> agg;
      for (auto& rec : r.prof_records) {
        float ms = 0.f;
        if (driver().cuEventElapsedTime(&ms, rec.start, rec.stop) == CUDA_SUCCESS) {
          Agg& a = agg[rec.label];
          a.count++;
          a.total_ms += ms;
          a.min_ms = std::min<double>(a.min_ms, ms);
          a.max_ms = std::max<double>(a.max_ms, ms);
          a.bytes = rec.bytes;
          a.flops = rec.flops;
        }
        r.prof_event_pool.push_back(rec.start);
        r.prof_event_pool.push_back(rec.stop);
      }
      r.prof_records.clear();
      std::vector<std::pair<std::string, Agg>> rows(agg.begin(), agg.end());
      std::sort(rows.begin(), rows.end(), [](auto& x, auto& y) { return x.second.total_ms > y.second.total_ms; });
      text = "[";
      for (size_t i = 0; i < rows.size(); ++i) {
        const Agg& a = rows[i].second;
        std::string name;
        for (char c : rows[i].first) {
          if (c == '"' || c == '\\') name += '\\';
          name += c;
        }
        const double avg_ms = a.total_ms / (double)a.count;
        text += strprintf("%s\n {\"name\": \"%s\", \"count\": %llu, \"total_ms\": %.6f, \"avg_us\": %.3f, \"min_us\": %.3f, \"max_us\": %.3f, "
                          "\"algorithmic_bytes\": %llu, \"flops\": %llu, \"GBs\": %.1f, \"TFLOPs\": %.3f}",
                          i ? "," : "", name.c_str(), (unsigned long long)a.count, a.total_ms, avg_ms * 1e3, a.min_ms * 1e3, a.max_ms * 1e3,
                          (unsigned long long)a.bytes, (unsigned long long)a.flops, avg_ms > 0 ? (double)a.bytes / avg_ms / 1e6 : 0.0,
                          avg_ms > 0 ? (double)a.flops / avg_ms / 1e9 : 0.0);
      }
      text += "\n]";
    }
    if (out_needed) *out_needed = text.size() + 1;
    if (out && capacity > 0) {
      CC_REQUIRE(capacity > text.size(), CC_ERR_ILLEGAL_ARGUMENT, "profile report needs %zu bytes", text.size() + 1);
      memcpy(out, text.c_str(), text.size() + 1);
      text.clear();
    }
  });
}

int cc_timer_start(void) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    join_all(r.aux);  // the stopwatch stream waits for everything submitted so far ...
    CC_CU(cuEventRecord(r.timer0, r.streams[(size_t)r.aux]));
    r.seq[(size_t)r.aux]++;
    for (int s = 0; s < (int)r.streams.size(); ++s)  // ... and nothing submitted from now on starts before the timestamp
      if (s != r.aux) CC_CU(cuStreamWaitEvent(r.streams[(size_t)s], r.timer0, 0));
  });
}
int cc_timer_stop(float* out_ms) {
  return guarded([&] {
    {
      Lock lock;
      require_init();
      Runtime& r = rt();
      CC_REQUIRE(out_ms, CC_ERR_ILLEGAL_ARGUMENT, "null output");
      join_all(r.aux);
      CC_CU(cuEventRecord(r.timer1, r.streams[(size_t)r.aux]));
      r.seq[(size_t)r.aux]++;
    }
    CC_CU(cuEventSynchronize(rt().timer1));
    CC_CU(cuEventElapsedTime(out_ms, rt().timer0, rt().timer1));
  });
}

// ---- NCCL -------------------------------------------------------------------------------------------------------------------------

int cc_comm_unique_id(void* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    Nccl& n = rt().nccl;
    n.load();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    n.check(n.GetUniqueId((ncclUniqueId*)out), "ncclGetUniqueId");
  });
}
int cc_comm_init(const void* id, int n_ranks, int rank) {
  return guarded([&] {
    Lock lock;
    require_init();
    Nccl& n = rt().nccl;
    n.load();
    CC_REQUIRE(id && n_ranks >= 1 && rank >= 0 && rank < n_ranks, CC_ERR_ILLEGAL_ARGUMENT, "bad communicator arguments");
    CC_REQUIRE(!n.comm, CC_ERR_ILLEGAL_ARGUMENT, "communicator already initialised");
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    cudaSetDevice(rt().ordinal);
    n.check(n.CommInitRank(&n.comm, n_ranks, uid, rank), "ncclCommInitRank");
    n.n_ranks = n_ranks;
    n.rank = rank;
    ++g_comm_generation;
  });
}
int cc_comm_generation(uint64_t* out) {
  return guarded([&] {
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = g_comm_generation.load();
  });
}
int cc_comm_enable_peer(void) {
  return guarded([&] {
    Lock lock;
    require_init();
    Nccl& n = rt().nccl;
    CC_REQUIRE(n.comm, CC_ERR_ILLEGAL_ARGUMENT, "cc_comm_enable_peer needs an initialised communicator");
    if (n.peer_mapped) {
      n.peer_enabled = true;
      return;
    }
    CC_REQUIRE(n.n_ranks <= kPeerMaxRanks, CC_ERR_UNSUPPORTED, "peer mailboxes support at most %d ranks", kPeerMaxRanks);
    const size_t bytes = peer_mailbox_bytes(n.n_ranks);
    CC_CU(cuMemAlloc(&n.mailbox, bytes));
    CUstream s0 = rt().streams[0];
    CC_CU(cuMemsetD32Async(n.mailbox, 0, bytes / 4, s0));
    // exchange the IPC handles through the communicator itself (64 bytes per rank)
    CUipcMemHandle mine;
    CC_CU(cuIpcGetMemHandle(&mine, n.mailbox));
    static_assert(sizeof(CUipcMemHandle) == 64, "CUipcMemHandle is 64 bytes");
    CUdeviceptr send = 0, recv = 0;
    CC_CU(cuMemAlloc(&send, 64));
    CC_CU(cuMemAlloc(&recv, 64 * (size_t)n.n_ranks));
    CC_CU(cuMemcpyHtoD(send, &mine, 64));
    n.check(n.AllGather((const void*)send, (void*)recv, 64, ncclChar, n.comm, (cudaStream_t)s0), "ncclAllGather(ipc handles)");
    CC_CU(cuStreamSynchronize(s0));
    std::vector<CUipcMemHandle> all((size_t)n.n_ranks);
    CC_CU(cuMemcpyDtoH(all.data(), recv, 64 * (size_t)n.n_ranks));
    driver().cuMemFree(send);
    driver().cuMemFree(recv);
    const size_t flag_off = peer_mailbox_flag_offset(n.n_ranks);
    n.mb = PeerMailboxes{};
    n.mb.world = n.n_ranks;
    n.mb.rank = n.rank;
    for (int r = 0; r < n.n_ranks; ++r) {
      CUdeviceptr base = n.mailbox;
      if (r != n.rank) {
        CUresult res = driver().cuIpcOpenMemHandle(&base, all[(size_t)r], CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
        if (res != CUDA_SUCCESS) {
          for (int q = 0; q < r; ++q)
            if (q != n.rank && n.peer_base[q]) driver().cuIpcCloseMemHandle(n.peer_base[q]);
          driver().cuMemFree(n.mailbox);
          n.mailbox = 0;
          fail(CC_ERR_UNSUPPORTED, strprintf("cuIpcOpenMemHandle(rank %d) failed (%d): no peer access between these GPUs", r, (int)res));
        }
      }
      n.peer_base[r] = base;
      n.mb.data[r] = (float*)base;
      n.mb.flags[r] = (unsigned*)(base + flag_off);
    }
    // nobody may push into a mailbox before its owner has zeroed it: one barrier through the communicator
    CUdeviceptr token = 0;
    CC_CU(cuMemAlloc(&token, 256));
    CC_CU(cuMemsetD32Async(token, 0, 64, s0));
    n.check(n.AllReduce((const void*)token, (void*)token, 1, ncclFloat, ncclSum, n.comm, (cudaStream_t)s0), "ncclAllReduce(barrier)");
    CC_CU(cuStreamSynchronize(s0));
    driver().cuMemFree(token);
    rt().seq[0]++;
    n.epoch = 0;
    if (!n.mb_dev) CC_CU(cuMemAlloc(&n.mb_dev, sizeof(PeerMailboxes)));
    CC_CU(cuMemcpyHtoD(n.mb_dev, &n.mb, sizeof(PeerMailboxes)));
    n.peer_enabled = true;
    n.peer_mapped = true;
  });
}

int cc_comm_route_peer(int on) {
  return guarded([&] {
    Lock lock;
    Nccl& n = rt().nccl;
    CC_REQUIRE(!on || n.peer_mapped, CC_ERR_ILLEGAL_ARGUMENT, "peer mailboxes are not mapped: call cc_comm_enable_peer first");
    n.peer_enabled = on != 0;
  });
}

int cc_comm_peer_enabled(int* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = rt().nccl.peer_enabled ? 1 : 0;
  });
}

int cc_reduce_sum_allreduce(cc_buffer in, uint64_t n_floats, cc_buffer out, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Nccl& n = r.nccl;
    Buffer* ib = as_buffer(in);
    Buffer* ob = as_buffer(out);
    CC_REQUIRE(n_floats <= ib->n_floats && ob->n_floats >= 1 && ib != ob, CC_ERR_ILLEGAL_ARGUMENT, "bad reduce_sum arguments");
    Buffer* sc = reduce_scratch();
    Op op{0, {ib}, {ob, sc}};
    op.label = "sum + all-reduce over NVLink (one kernel)";
    op.bytes = n_floats * 4 + 4;
    op_begin(op, waits, n_waits);
    if (n.comm && n.peer_enabled && !rt().capture) {  // (an epoch captured into a graph would be replayed: NCCL inside captures)
      // ONE kernel: local reduction + all-reduce of its result over NVLink peer memory
      launch_reduce_sum_allreduce((const float*)ib->ptr, n_floats, (float*)ob->ptr, (float*)sc->ptr, (unsigned*)r.reduce_counter, r.info.sm_count,
                                  n.mb, ++n.epoch, (cudaStream_t)op.cu());
      r.stats.device_kernels++;
    } else {
      launch_reduce_sum((const float*)ib->ptr, n_floats, (float*)ob->ptr, (float*)sc->ptr, (unsigned*)r.reduce_counter, r.info.sm_count,
                        (cudaStream_t)op.cu());
      r.stats.device_kernels++;
      if (n.comm)
        n.check(n.AllReduce((const void*)ob->ptr, (void*)ob->ptr, 1, ncclFloat, ncclSum, n.comm, (cudaStream_t)op.cu()), "ncclAllReduce");
    }
    r.stats.launches++;
    op_end(op, out_event);
  });
}

namespace {
// ---- NVLS multicast symmetric memory -------------------------------------------------------------------------------------------------
// One multicast object over the ranks' devices (rank 0 creates it and hands its POSIX file descriptor to the other processes over an
// abstract unix socket, SCM_RIGHTS), every rank binds `bytes` of its own device memory at offset 0 and maps two views: its own memory, and
// the multicast address through which a store is replicated by the NVSwitch into every rank's memory at the same offset.

int send_fd(int sock, int fd) {
  char data = 'f';
  iovec iov{&data, 1};
  char ctrl[CMSG_SPACE(sizeof(int))] = {0};
  msghdr msg{};
  msg.msg_iov = &iov, msg.msg_iovlen = 1, msg.msg_control = ctrl, msg.msg_controllen = sizeof ctrl;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET, c->cmsg_type = SCM_RIGHTS, c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}
int recv_fd(int sock) {
  char data = 0;
  iovec iov{&data, 1};
  char ctrl[CMSG_SPACE(sizeof(int))] = {0};
  msghdr msg{};
  msg.msg_iov = &iov, msg.msg_iovlen = 1, msg.msg_control = ctrl, msg.msg_controllen = sizeof ctrl;
  if (recvmsg(sock, &msg, 0) != 1) return -1;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd = -1;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}
sockaddr_un abstract_address(uint64_t token, socklen_t* len) {
  sockaddr_un a{};
  a.sun_family = AF_UNIX;
  const int n = snprintf(a.sun_path + 1, sizeof a.sun_path - 1, "compute_cuda_mc_%016llx", (unsigned long long)token);  // sun_path[0] = 0: abstract
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + (size_t)n);
  return a;
}

// a barrier + an 8-byte broadcast through the communicator (stream 0, synchronised): the only coordination the set-up needs
uint64_t comm_broadcast_u64(Nccl& n, uint64_t value) {
  CUstream s0 = rt().streams[0];
  CUdeviceptr d = 0;
  CC_CU(cuMemAlloc(&d, 256));
  CC_CU(cuMemcpyHtoD(d, &value, 8));
  n.check(n.Broadcast((const void*)d, (void*)d, 8, ncclChar, 0, n.comm, (cudaStream_t)s0), "ncclBroadcast(multicast token)");
  CC_CU(cuStreamSynchronize(s0));
  CC_CU(cuMemcpyDtoH(&value, d, 8));
  driver().cuMemFree(d);
  rt().seq[0]++;
  return value;
}
// every rank reports 1 (fine so far) or 0; returns true only if all did — so that a rank that cannot take part makes ALL fall back together
bool comm_all_ok(Nccl& n, bool ok) {
  CUstream s0 = rt().streams[0];
  CUdeviceptr d = 0;
  CC_CU(cuMemAlloc(&d, 256));
  float v = ok ? 0.f : 1.f;
  CC_CU(cuMemcpyHtoD(d, &v, 4));
  n.check(n.AllReduce((const void*)d, (void*)d, 1, ncclFloat, ncclSum, n.comm, (cudaStream_t)s0), "ncclAllReduce(multicast agreement)");
  CC_CU(cuStreamSynchronize(s0));
  CC_CU(cuMemcpyDtoH(&v, d, 4));
  driver().cuMemFree(d);
  rt().seq[0]++;
  return v == 0.f;
}

// Collective. Returns nullptr (on every rank alike) when multicast cannot be used here; the caller falls back to CUDA-IPC peer mappings.
Buffer* multicast_alloc(uint64_t n_floats) {
  Runtime& r = rt();
  Nccl& n = r.nccl;
  Driver& d = driver();
  if (const char* e = getenv("CC_MULTICAST"))
    if (atoi(e) == 0) return nullptr;  // (set identically on every rank, like every CC_* switch)
  bool ok = d.cuMulticastCreate && d.cuMulticastAddDevice && d.cuMulticastBindMem && d.cuMulticastGetGranularity && d.cuMemCreate && d.cuMemMap &&
            d.cuMemAddressReserve && d.cuMemSetAccess && d.cuMemExportToShareableHandle && d.cuMemImportFromShareableHandle;
  int supported = 0;
  if (ok) ok = d.cuDeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, r.dev) == CUDA_SUCCESS && supported;
  CUmulticastObjectProp prop{};
  prop.numDevices = (unsigned)n.n_ranks;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  prop.size = (size_t)n_floats * 4;
  if (ok) ok = d.cuMulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && gran > 0;
  if (!comm_all_ok(n, ok)) return nullptr;
  const size_t bytes = ((size_t)n_floats * 4 + gran - 1) / gran * gran;
  prop.size = bytes;

  // 1. the multicast object: created by rank 0, imported by the others through a file descriptor passed over an abstract unix socket
  CUmemGenericAllocationHandle mc = 0;
  uint64_t token = 0;
  int listener = -1;
  if (n.rank == 0) {
    ok = d.cuMulticastCreate(&mc, &prop) == CUDA_SUCCESS;
    if (ok) {
      timespec ts;
      clock_gettime(CLOCK_REALTIME, &ts);
      token = ((uint64_t)getpid() << 32) ^ (uint64_t)ts.tv_nsec ^ ((uint64_t)ts.tv_sec << 20) ^ (uint64_t)r.next_uid;
      listener = socket(AF_UNIX, SOCK_STREAM, 0);
      socklen_t len = 0;
      sockaddr_un addr = abstract_address(token, &len);
      ok = listener >= 0 && bind(listener, (sockaddr*)&addr, len) == 0 && listen(listener, n.n_ranks) == 0;
    }
    if (!ok) token = 0;
  }
  token = comm_broadcast_u64(n, token);  // 0 = rank 0 could not create / listen: everybody falls back
  if (token == 0) {
    if (listener >= 0) close(listener);
    if (mc) d.cuMemRelease(mc);
    return nullptr;
  }
  if (n.rank == 0) {
    int fd = -1;
    ok = d.cuMemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    for (int peer = 1; peer < n.n_ranks; ++peer) {
      int c = accept(listener, nullptr, nullptr);
      if (c < 0 || !ok || send_fd(c, fd) != 0) ok = false;
      if (c >= 0) close(c);
    }
    if (fd >= 0) close(fd);
    close(listener);
  } else {
    int sock = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len = 0;
    sockaddr_un addr = abstract_address(token, &len);
    ok = sock >= 0 && connect(sock, (sockaddr*)&addr, len) == 0;  // (rank 0 was listening before the broadcast completed)
    int fd = ok ? recv_fd(sock) : -1;
    if (sock >= 0) close(sock);
    ok = fd >= 0 && d.cuMemImportFromShareableHandle(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
    if (fd >= 0) close(fd);
  }
  // 2. every device joins; only then may memory be bound
  if (ok) ok = d.cuMulticastAddDevice(mc, r.dev) == CUDA_SUCCESS;
  if (!comm_all_ok(n, ok)) {
    if (mc) d.cuMemRelease(mc);
    return nullptr;
  }
  // 3. this rank's memory (shareable, as an imported multicast object requires), bound at offset 0, and the two mappings
  CUmemAllocationProp ap{};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = r.ordinal;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mem = 0;
  CUdeviceptr uc = 0, mcva = 0;
  CUmemAccessDesc access{};
  access.location = ap.location;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ok = d.cuMemCreate(&mem, bytes, &ap, 0) == CUDA_SUCCESS;
  if (ok) ok = d.cuMulticastBindMem(mc, 0, mem, 0, bytes, 0) == CUDA_SUCCESS;
  if (ok) ok = d.cuMemAddressReserve(&uc, bytes, gran, 0, 0) == CUDA_SUCCESS && d.cuMemMap(uc, bytes, 0, mem, 0) == CUDA_SUCCESS &&
               d.cuMemSetAccess(uc, bytes, &access, 1) == CUDA_SUCCESS;
  if (ok) ok = d.cuMemAddressReserve(&mcva, bytes, gran, 0, 0) == CUDA_SUCCESS && d.cuMemMap(mcva, bytes, 0, mc, 0) == CUDA_SUCCESS &&
               d.cuMemSetAccess(mcva, bytes, &access, 1) == CUDA_SUCCESS;
  if (!comm_all_ok(n, ok)) {  // (also the barrier: every rank has bound its memory before anybody stores through the multicast address)
    // (best effort clean-up of a half-built mapping; the process keeps working on the IPC route)
    if (mcva) d.cuMemAddressFree(mcva, bytes);
    if (uc) d.cuMemAddressFree(uc, bytes);
    if (mem) d.cuMemRelease(mem);
    if (mc) d.cuMemRelease(mc);
    return nullptr;
  }
  Buffer* b = new Buffer();
  b->ptr = uc;
  b->mc_ptr = mcva;
  b->mc_handle = mc;
  b->mem_handle = mem;
  b->vmm_bytes = bytes;
  b->n_floats = n_floats;
  b->owned = false;  // not pool memory: unmapped and released with the communicator
  b->uid = r.next_uid++;
  return b;
}
}  // namespace

int cc_comm_symmetric_alloc(uint64_t n_floats, cc_buffer* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Nccl& n = r.nccl;
    CC_REQUIRE(out && n_floats > 0, CC_ERR_ILLEGAL_ARGUMENT, "bad symmetric allocation request");
    CC_REQUIRE(n.comm && n.peer_mapped, CC_ERR_ILLEGAL_ARGUMENT, "cc_comm_symmetric_alloc needs cc_comm_enable_peer first");
    // NVLS first: a multicast mapping lets the fused all-gather send every block ONCE (the switch replicates it) instead of once per peer
    if (Buffer* mcb = multicast_alloc(n_floats)) {
      mcb->rc.store(2);  // the caller's handle + the communicator's list
      r.buffers.insert(mcb);
      n.symmetric.push_back(mcb);
      *out = (cc_buffer)(uintptr_t)mcb;
      return;
    }
    CUstream s0 = r.streams[0];
    const size_t bytes = ((size_t)n_floats * 4 + 1023) / 1024 * 1024;
    CUdeviceptr mine = 0;
    CC_CU(cuMemAlloc(&mine, bytes));
    CUipcMemHandle h;
    CC_CU(cuIpcGetMemHandle(&h, mine));
    CUdeviceptr send = 0, recv = 0;
    CC_CU(cuMemAlloc(&send, 64));
    CC_CU(cuMemAlloc(&recv, 64 * (size_t)n.n_ranks));
    CC_CU(cuMemcpyHtoD(send, &h, 64));
    n.check(n.AllGather((const void*)send, (void*)recv, 64, ncclChar, n.comm, (cudaStream_t)s0), "ncclAllGather(ipc handles)");
    CC_CU(cuStreamSynchronize(s0));
    std::vector<CUipcMemHandle> all((size_t)n.n_ranks);
    CC_CU(cuMemcpyDtoH(all.data(), recv, 64 * (size_t)n.n_ranks));
    driver().cuMemFree(send);
    driver().cuMemFree(recv);
    r.seq[0]++;
    Buffer* b = new Buffer();
    b->ptr = mine;
    b->n_floats = n_floats;
    b->owned = false;  // not pool memory: freed with the communicator
    b->uid = r.next_uid++;
    b->peers.assign((size_t)n.n_ranks, 0);
    for (int q = 0; q < n.n_ranks; ++q) {
      CUdeviceptr base = mine;
      if (q != n.rank) {
        CUresult res = driver().cuIpcOpenMemHandle(&base, all[(size_t)q], CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
        if (res != CUDA_SUCCESS) {
          for (int z = 0; z < q; ++z)
            if (z != n.rank && b->peers[(size_t)z]) driver().cuIpcCloseMemHandle(b->peers[(size_t)z]);
          driver().cuMemFree(mine);
          delete b;
          fail(CC_ERR_UNSUPPORTED, strprintf("cuIpcOpenMemHandle(rank %d) failed (%d)", q, (int)res));
        }
      }
      b->peers[(size_t)q] = base;
    }
    b->rc.store(2);  // the caller's handle + the communicator's list
    r.buffers.insert(b);
    n.symmetric.push_back(b);
    *out = (cc_buffer)(uintptr_t)b;
  });
}

int cc_buffer_is_multicast(cc_buffer b, int* out) {
  return guarded([&] {
    Lock lock;
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    *out = as_buffer(b)->mc_ptr ? 1 : 0;
  });
}

int cc_matmul_3xtf32_allgather(cc_buffer a, cc_buffer b, cc_buffer gathered, int64_t m_shard, int64_t n, int64_t k, const cc_event* waits, int n_waits,
                               cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Nccl& nc = r.nccl;
    Buffer* ab = as_buffer(a);
    Buffer* bb = as_buffer(b);
    Buffer* gb = as_buffer(gathered);
    CC_REQUIRE(nc.comm && nc.peer_mapped, CC_ERR_ILLEGAL_ARGUMENT, "cc_matmul_3xtf32_allgather needs cc_comm_enable_peer");
    CC_REQUIRE(!r.capture, CC_ERR_UNSUPPORTED, "the fused all-gather's flag barriers carry an epoch: not inside a graph capture");
    CC_REQUIRE((int)gb->peers.size() == nc.n_ranks || gb->mc_ptr, CC_ERR_ILLEGAL_ARGUMENT, "`gathered` must come from cc_comm_symmetric_alloc");
    CC_REQUIRE(m_shard > 0 && n > 0 && k > 0 && n % 4 == 0, CC_ERR_UNSUPPORTED, "fused all-gather needs N %% 4 == 0 (got %lld x %lld x %lld)",
               (long long)m_shard, (long long)n, (long long)k);
    CC_REQUIRE(ab->n_floats >= (uint64_t)(m_shard * k) && bb->n_floats >= (uint64_t)(k * n) &&
                   gb->n_floats >= (uint64_t)(m_shard * n) * (uint64_t)nc.n_ranks,
               CC_ERR_ILLEGAL_ARGUMENT, "matmul buffers too small");
    GemmRun g = gemm_prepare(bb, m_shard, n, k, ab, /*gather_epilogue=*/true);
    bool launched = false;
    try {
      Op op{0, {ab, bb}, {gb}};  // collectives stay on stream 0, in call order
      op.label = gb->mc_ptr ? "contraction 3xTF32 + all-gather epilogue (NVLS multicast stores)" : "contraction 3xTF32 + all-gather epilogue";
      op.flops = 2ull * (uint64_t)m_shard * (uint64_t)n * (uint64_t)k;
      g.declare(op);
      op_begin(op, waits, n_waits);
      // entry barrier: every rank has finished whatever still read its copy of `gathered` (stream order on each rank) ...
      launch_peer_barrier(nc.mb, ++nc.epoch, (cudaStream_t)op.cu());
      GemmWorkspace ws{g.a_hi ? (float*)g.a_hi->ptr : nullptr, g.a_lo ? (float*)g.a_lo->ptr : nullptr, (float*)g.bt_hi->ptr, (float*)g.bt_lo->ptr};
      float* dst[kPeerMaxRanks] = {nullptr};
      for (int q = 0; q < nc.n_ranks && q < (int)gb->peers.size(); ++q) dst[q] = (float*)gb->peers[(size_t)q];
      int kernels = launch_gemm_3xtf32_allgather((const float*)ab->ptr, (const float*)bb->ptr, dst, nc.n_ranks, nc.rank, m_shard, n, k, ws,
                                                 r.info.sm_count, (TensorMapEncodeFn)driver().cuTensorMapEncodeTiled, (cudaStream_t)op.cu(), g.b_ready,
                                                 (float*)gb->mc_ptr);
      // ... exit barrier: every rank's blocks have landed in every copy
      launch_peer_barrier(nc.mb, ++nc.epoch, (cudaStream_t)op.cu());
      launched = true;
      r.stats.device_kernels += (uint64_t)kernels + 2;
      r.stats.launches++;
      op_end(op, out_event);
    } catch (...) {
      gemm_finish(g, bb, n, k, launched);
      throw;
    }
    gemm_finish(g, bb, n, k, true);
  });
}

int cc_comm_destroy(void) {
  return guarded([&] {
    Lock lock;
    Nccl& n = rt().nccl;
    close_peers(n);
    if (n.comm) {
      driver().cuCtxSynchronize();
      n.check(n.CommDestroy(n.comm), "ncclCommDestroy");
      n.comm = nullptr;
      n.n_ranks = 0;
    }
  });
}
int cc_comm_info(int* out_n, int* out_rank) {
  return guarded([&] {
    Lock lock;
    if (out_n) *out_n = rt().nccl.comm ? rt().nccl.n_ranks : 1;
    if (out_rank) *out_rank = rt().nccl.comm ? rt().nccl.rank : 0;
  });
}

int cc_allreduce_sum(cc_buffer buf, uint64_t n_floats, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Nccl& n = rt().nccl;
    Buffer* b = as_buffer(buf);
    CC_REQUIRE(n_floats <= b->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "allreduce: buffer too small");
    Op op{0, {}, {b}};
    op.label = "all-reduce";
    op.bytes = n_floats * 4;
    op_begin(op, waits, n_waits);
    if (n.comm && n.peer_enabled && !rt().capture && n_floats <= (uint64_t)kPeerCapFloats) {  // (a captured epoch would be replayed)
      launch_peer_allreduce((float*)b->ptr, n_floats, n.mb, ++n.epoch, (cudaStream_t)op.cu());
      rt().stats.device_kernels++;
    } else if (n.comm) {
      n.check(n.AllReduce((const void*)b->ptr, (void*)b->ptr, n_floats, ncclFloat, ncclSum, n.comm, (cudaStream_t)op.cu()), "ncclAllReduce");
    }
    op_end(op, out_event);
  });
}
int cc_allgather(cc_buffer send, cc_buffer recv, uint64_t n_per_rank, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Nccl& n = rt().nccl;
    Buffer* s = as_buffer(send);
    Buffer* d = as_buffer(recv);
    int ranks = n.comm ? n.n_ranks : 1;
    CC_REQUIRE(n_per_rank <= s->n_floats && n_per_rank * ranks <= d->n_floats && s != d, CC_ERR_ILLEGAL_ARGUMENT, "allgather: bad buffers");
    Op op{0, {s}, {d}};
    op.label = "all-gather";
    op.bytes = n_per_rank * 4 * (uint64_t)ranks;
    op_begin(op, waits, n_waits);
    if (n.comm && n.peer_enabled && !rt().capture && n_per_rank <= (uint64_t)kPeerCapFloats) {
      // small blocks (the row sums of a sharded tensor): one kernel over the NVLink mailboxes instead of an NCCL call
      launch_peer_allgather((const float*)s->ptr, (float*)d->ptr, n_per_rank, n.mb, ++n.epoch, (cudaStream_t)op.cu());
      rt().stats.device_kernels++;
    } else if (n.comm)
      n.check(n.AllGather((const void*)s->ptr, (void*)d->ptr, n_per_rank, ncclFloat, n.comm, (cudaStream_t)op.cu()), "ncclAllGather");
    else
      CC_CU(cuMemcpyDtoDAsync(d->ptr, s->ptr, (size_t)n_per_rank * 4, op.cu()));
    op_end(op, out_event);
  });
}
int cc_broadcast(cc_buffer buf, uint64_t n_floats, int root, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Nccl& n = rt().nccl;
    Buffer* b = as_buffer(buf);
    CC_REQUIRE(n_floats <= b->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "broadcast: buffer too small");
    Op op{0, {}, {b}};
    op.label = "broadcast";
    op.bytes = n_floats * 4;
    op_begin(op, waits, n_waits);
    if (n.comm) n.check(n.Broadcast((const void*)b->ptr, (void*)b->ptr, n_floats, ncclFloat, root, n.comm, (cudaStream_t)op.cu()), "ncclBroadcast");
    op_end(op, out_event);
  });
}

// ---- CUDA graphs over a caller-visible sequence of evaluations ------------------------------------------------------------------------

namespace {
Runtime::Graph* as_graph(cc_graph h) {
  Runtime::Graph* g = (Runtime::Graph*)(uintptr_t)h;
  CC_REQUIRE(g && rt().graphs.count(g), CC_ERR_ILLEGAL_ARGUMENT, "invalid graph handle");
  return g;
}
// While a capture is open the memory pool is the capture's own: blocks freed by captured commands may be reused by later captured
// commands (same order at every replay) but must never go back to the general pool while the graph lives — a replay writes them.
std::map<size_t, std::vector<Block>>& parked_pool() {
  static std::map<size_t, std::vector<Block>> p;
  return p;
}
}  // namespace

int cc_graph_begin(void) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    CC_REQUIRE(!r.capture, CC_ERR_ILLEGAL_ARGUMENT, "a graph capture is already open");
    CC_REQUIRE(driver().cuStreamBeginCapture && driver().cuStreamEndCapture && driver().cuGraphInstantiateWithFlags && driver().cuGraphLaunch,
               CC_ERR_UNSUPPORTED, "this CUDA driver has no stream capture");
    // the runtime's lazily created scratch (fold partials, block counters) is set up with a memset + synchronise: not inside a capture
    reduce_scratch();
    col_counters_for(0);
    // everything submitted so far completes first, so the captured commands need no edges to the outside
    CC_CU(cuCtxSynchronize());
    for (size_t s = 0; s < r.synced.size(); ++s)
      for (size_t a = 0; a < r.synced[s].size(); ++a) r.synced[s][a] = r.seq[a];
    parked_pool().swap(r.pool);  // r.pool is now empty: the capture's own pool
    CUresult res = driver().cuStreamBeginCapture(r.streams[0], CU_STREAM_CAPTURE_MODE_RELAXED);
    if (res != CUDA_SUCCESS) {
      parked_pool().swap(r.pool);
      check_cu(res, "cuStreamBeginCapture");
    }
    r.capture = new Runtime::Graph();
    r.capture->commands = r.stats.device_kernels;  // (the difference at cc_graph_end = kernels recorded)
  });
}

int cc_graph_end(cc_graph* out) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    CC_REQUIRE(out, CC_ERR_ILLEGAL_ARGUMENT, "null output");
    CC_REQUIRE(r.capture, CC_ERR_ILLEGAL_ARGUMENT, "no graph capture is open");
    Runtime::Graph* g = r.capture;
    r.capture = nullptr;
    g->commands = r.stats.device_kernels - g->commands;
    r.stats.device_kernels -= g->commands;  // recorded, not run: replays count them
    for (Buffer* b : g->reads) b->rc.fetch_add(1);
    for (Buffer* b : g->writes) b->rc.fetch_add(1);
    // the capture's pool becomes the graph's property; the general pool comes back
    for (auto& kv : r.pool)
      for (Block& blk : kv.second) g->blocks.push_back(std::move(blk));
    r.pool.clear();
    parked_pool().swap(r.pool);
    parked_pool().clear();
    r.graphs.insert(g);
    const bool dbg = getenv("CC_GRAPH_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[graph] end: %llu kernels, %zu reads, %zu writes, %zu blocks\n", (unsigned long long)g->commands, g->reads.size(), g->writes.size(), g->blocks.size());
    CUresult res = driver().cuStreamEndCapture(r.streams[0], &g->graph);
    if (dbg) fprintf(stderr, "[graph] cuStreamEndCapture -> %d graph=%p\n", (int)res, (void*)g->graph);
    if (res == CUDA_SUCCESS && !g->graph) res = CUDA_ERROR_STREAM_CAPTURE_INVALIDATED;  // (a command that failed inside the capture)
    if (res == CUDA_SUCCESS) res = driver().cuGraphInstantiateWithFlags(&g->exec, g->graph, 0);
    if (dbg) fprintf(stderr, "[graph] cuGraphInstantiate -> %d exec=%p\n", (int)res, (void*)g->exec);
    if (res != CUDA_SUCCESS) {
      g->exec = nullptr;
      cc_graph_release((cc_graph)(uintptr_t)g);
      check_cu(res, "cuStreamEndCapture / cuGraphInstantiate");
    }
    *out = (cc_graph)(uintptr_t)g;
  });
}

int cc_graph_launch(cc_graph h, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Runtime& r = rt();
    Runtime::Graph* g = as_graph(h);
    CC_REQUIRE(!r.capture, CC_ERR_UNSUPPORTED, "a graph cannot be launched while another one is being captured");
    CC_REQUIRE(g->exec, CC_ERR_ILLEGAL_ARGUMENT, "the graph was not instantiated");
    BufferList reads, writes;
    for (Buffer* b : g->reads) reads.push_back(b);
    for (Buffer* b : g->writes) writes.push_back(b);
    Op op{0, reads, writes};
    op.label = "graph replay";
    op_begin(op, waits, n_waits);
    CC_CU(cuGraphLaunch(g->exec, op.cu()));
    r.stats.launches++;
    r.stats.device_kernels += g->commands;
    op_end(op, out_event);
  });
}

int cc_graph_info(cc_graph h, uint64_t* out_commands, uint64_t* out_buffers) {
  return guarded([&] {
    Lock lock;
    Runtime::Graph* g = as_graph(h);
    if (out_commands) *out_commands = g->commands;
    if (out_buffers) *out_buffers = g->reads.size() + g->writes.size();
  });
}

int cc_graph_release(cc_graph h) {
  return guarded([&] {
    Lock lock;
    Runtime& r = rt();
    Runtime::Graph* g = as_graph(h);
    r.graphs.erase(g);
    if (r.initialized) {
      CC_CU(cuCtxSetCurrent(r.ctx));
      // replays still in flight read and write the graph's memory: let stream 0 drain before any of it can be handed out again
      if (g->exec) driver().cuStreamSynchronize(r.streams[0]);
      if (g->exec && driver().cuGraphExecDestroy) driver().cuGraphExecDestroy(g->exec);
      if (g->graph && driver().cuGraphDestroy) driver().cuGraphDestroy(g->graph);
      for (Block& blk : g->blocks) {
        blk.pending.clear();
        r.bytes_pooled += blk.bytes;
        r.pool[blk.bytes].push_back(std::move(blk));
      }
    }
    for (Buffer* b : g->reads) release(b);
    for (Buffer* b : g->writes) release(b);
    delete g;
  });
}

int cc_buffer_copy(cc_buffer dst, cc_buffer src, uint64_t n_floats, const cc_event* waits, int n_waits, cc_event* out_event) {
  return guarded([&] {
    Lock lock;
    require_init();
    Buffer* d = as_buffer(dst);
    Buffer* s = as_buffer(src);
    CC_REQUIRE(d != s && n_floats <= d->n_floats && n_floats <= s->n_floats, CC_ERR_ILLEGAL_ARGUMENT, "copy: bad buffers");
    BufferList in;
    in.push_back(s);
    Op op{pick_stream_for(in, {d}), in, {d}};
    op.label = "device-to-device copy";
    op.bytes = n_floats * 8;
    op_begin(op, waits, n_waits);
    CC_CU(cuMemcpyDtoDAsync(d->ptr, s->ptr, (size_t)n_floats * 4, op.cu()));
    op_end(op, out_event);
  });
}

// ---- leading-axis sharding ---------------------------------------------------------------------------------------------------------

int cc_shard_rows(int64_t rows, int n_ranks, int rank, int64_t* out_first, int64_t* out_count) {
  return guarded([&] {
    CC_REQUIRE(rows >= 0 && n_ranks >= 1 && rank >= 0 && rank < n_ranks && out_first && out_count, CC_ERR_ILLEGAL_ARGUMENT, "bad shard request");
    const int64_t base = rows / n_ranks, extra = rows % n_ranks;
    *out_first = rank * base + std::min<int64_t>(rank, extra);
    *out_count = base + (rank < extra ? 1 : 0);
  });
}

int cc_shard_agree(uint64_t value, int* out_all_equal) {
  int st = guarded([&] { CC_REQUIRE(out_all_equal, CC_ERR_ILLEGAL_ARGUMENT, "null output"); });
  if (st != CC_OK) return st;
  int world = 1, rank = 0;
  st = cc_comm_info(&world, &rank);
  if (st != CC_OK) return st;
  if (world == 1) {
    *out_all_equal = 1;
    return CC_OK;
  }
  // four 16-bit digits, each exactly representable in a float: the exchange is a plain all-gather of floats
  float mine[4];
  for (int d = 0; d < 4; ++d) mine[d] = (float)((value >> (16 * d)) & 0xffffu);
  cc_buffer send = 0, recv = 0;
  st = cc_buffer_from_host(mine, 4, &send, nullptr);
  if (st == CC_OK) st = cc_buffer_alloc(4ull * (uint64_t)world, &recv);
  if (st == CC_OK) st = cc_allgather(send, recv, 4, nullptr, 0, nullptr);
  std::vector<float> all(4 * (size_t)world, 0.f);
  if (st == CC_OK) st = cc_buffer_to_host(recv, 0, all.data(), all.size(), nullptr, 0, nullptr);
  if (send) cc_buffer_release(send);
  if (recv) cc_buffer_release(recv);
  if (st != CC_OK) return st;
  int equal = 1;
  for (int r = 0; r < world; ++r)
    for (int d = 0; d < 4; ++d)
      if (all[(size_t)r * 4 + d] != mine[d]) equal = 0;
  *out_all_equal = equal;
  return CC_OK;
}

int cc_shard_launch_allreduce(cc_kernel h, const cc_buffer* args, int n_args, cc_buffer out, const cc_event* waits, int n_waits, cc_event* out_event) {
  uint64_t n = 0;
  bool same_kernel = false;
  int st = guarded([&] {
    Lock lock;
    require_init();
    const Plan& p = as_kernel(h)->plan;
    Nccl& nc = rt().nccl;
    n = p.out_floats;
    // a generated reduction that writes its final values itself completes the all-reduce over the peer mailboxes too: one launch
    same_kernel = p.collective == 1 && nc.comm && nc.n_ranks > 1 && nc.peer_enabled && nc.mb_dev && n <= (uint64_t)kPeerCapFloats && !rt().capture;
  });
  if (st != CC_OK) return st;
  if (same_kernel) return launch_kernel(h, args, n_args, out, waits, n_waits, out_event, true);
  st = cc_launch(h, args, n_args, out, waits, n_waits, nullptr);
  if (st != CC_OK) return st;
  // the hazard tracker orders the combine after the launch (it writes the buffer the launch wrote)
  return cc_allreduce_sum(out, n, nullptr, 0, out_event);
}

int cc_shard_launch_allgather(cc_kernel h, const cc_buffer* args, int n_args, cc_buffer gathered, const cc_event* waits, int n_waits, cc_event* out_event,
                              int* out_fused) {
  bool fuse = false, same_kernel = false;
  uint64_t n = 0;
  int64_t M = 0, N = 0, K = 0;
  int st = guarded([&] {
    Lock lock;
    require_init();
    const Plan& p = as_kernel(h)->plan;
    Nccl& nc = rt().nccl;
    Buffer* gb = as_buffer(gathered);
    const int ranks = nc.comm ? nc.n_ranks : 1;
    n = p.out_floats;
    CC_REQUIRE(gb->n_floats >= n * (uint64_t)ranks, CC_ERR_ILLEGAL_ARGUMENT, "gathered buffer has %llu floats, %d blocks of %llu are needed",
               (unsigned long long)gb->n_floats, ranks, (unsigned long long)n);
    M = p.M, N = p.N, K = p.K;
    fuse = p.kind == PLAN_CONTRACTION && !p.gathered_panels && n_args == 2 && nc.comm && nc.peer_mapped && nc.peer_enabled &&
           ((int)gb->peers.size() == nc.n_ranks || gb->mc_ptr) && N % 4 == 0 && !rt().capture;
    // a row-owner reduction gathers its own outputs over the peer mailboxes (lane 0 of every output): one launch
    same_kernel = p.collective == 2 && nc.comm && nc.n_ranks > 1 && nc.peer_enabled && nc.mb_dev && n <= (uint64_t)kPeerCapFloats && !rt().capture;
  });
  if (st != CC_OK) return st;
  if (out_fused) *out_fused = (fuse || same_kernel) ? 1 : 0;
  if (fuse) return cc_matmul_3xtf32_allgather(args[0], args[1], gathered, M, N, K, waits, n_waits, out_event);
  if (same_kernel) return launch_kernel(h, args, n_args, gathered, waits, n_waits, out_event, true);
  cc_buffer part = 0;
  st = cc_buffer_alloc(n, &part);
  if (st == CC_OK) st = cc_launch(h, args, n_args, part, waits, n_waits, nullptr);
  if (st == CC_OK) st = cc_allgather(part, gathered, n, nullptr, 0, out_event);
  if (part) cc_buffer_release(part);  // (deferred by the runtime until the commands reading it have run)
  return st;
}

}  // extern "C"
