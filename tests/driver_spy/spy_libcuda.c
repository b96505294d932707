/* TEST INFRASTRUCTURE ONLY (see README.md): a libcuda.so.1 that counts calls and executes nothing. */
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_NAMES 128
typedef struct {
  char name[48];
  long count;
  /* fault injection (SPY_FAIL=name:skip:code[:times],...): after `skip` successful calls the next `times` calls return `code` */
  long skip, times;
  int code;
} Entry;
static Entry g_counts[MAX_NAMES];
static int g_n = 0;
static long g_pdl_launches = 0, g_plain_launches = 0;
static uintptr_t g_streams_seen[64];
static int g_nstreams = 0;
static uintptr_t g_bump = 0x7000000000ull;
static uintptr_t g_handle = 0x1000;

static Entry* counter(const char* name) {
  for (int i = 0; i < g_n; ++i)
    if (!strcmp(g_counts[i].name, name)) return &g_counts[i];
  if (g_n == MAX_NAMES) abort();
  Entry* e = &g_counts[g_n++];
  snprintf(e->name, sizeof e->name, "%s", name);
  const char* spec = getenv("SPY_FAIL");
  const size_t len = strlen(name);
  while (spec && *spec) {
    if (!strncmp(spec, name, len) && spec[len] == ':') {
      long skip = 0, times = 1;
      int code = 999;
      const int got = sscanf(spec + len + 1, "%ld:%d:%ld", &skip, &code, &times);
      if (got >= 2) e->skip = skip, e->code = code, e->times = got >= 3 ? times : 1;
      break;
    }
    spec = strchr(spec, ',');
    if (spec) ++spec;
  }
  return e;
}
static int fault(Entry* e) {  /* called after the count was incremented */
  if (e->times > 0 && e->count > e->skip) {
    --e->times;
    return e->code;
  }
  return 0;
}
static void saw_stream(void* s) {
  for (int i = 0; i < g_nstreams; ++i)
    if (g_streams_seen[i] == (uintptr_t)s) return;
  if (g_nstreams < 64) g_streams_seen[g_nstreams++] = (uintptr_t)s;
}

/* ---- entry points that must hand something back ---- */
#define COUNT(name) \
  static Entry* c_ = NULL; \
  if (!c_) c_ = counter(#name); \
  ++c_->count; \
  { const int f_ = fault(c_); if (f_) return (CUresult)f_; }
static CUresult s_cuDeviceGetCount(int* n) { COUNT(cuDeviceGetCount); *n = 1; return 0; }
static CUresult s_cuDeviceGet(CUdevice* d, int o) { COUNT(cuDeviceGet); (void)o; *d = 0; return 0; }
static CUresult s_cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice d) { COUNT(cuDevicePrimaryCtxRetain); (void)d; *c = (CUcontext)0x1234; return 0; }
static CUresult s_cuDeviceGetAttribute(int* v, CUdevice_attribute a, CUdevice d) {
  COUNT(cuDeviceGetAttribute);
  (void)d;
  switch (a) {
    case CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT: *v = 148; break;
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR: *v = 10; break;
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR: *v = 0; break;
    case CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN: *v = 232448; break;
    case CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE: *v = 126 << 20; break;
    case CU_DEVICE_ATTRIBUTE_CLOCK_RATE: *v = 1965000; break;
    case CU_DEVICE_ATTRIBUTE_MEMORY_CLOCK_RATE: *v = 4000000; break;
    default: *v = 1;
  }
  return 0;
}
static CUresult s_cuDeviceTotalMem(size_t* t, CUdevice d) { COUNT(cuDeviceTotalMem); (void)d; *t = (size_t)180 << 30; return 0; }
static CUresult s_cuDeviceGetName(char* n, int len, CUdevice d) { COUNT(cuDeviceGetName); (void)d; snprintf(n, (size_t)len, "driver spy (no device)"); return 0; }
static CUresult s_cuDriverGetVersion(int* v) { COUNT(cuDriverGetVersion); *v = 12090; return 0; }
static CUresult s_cuMemAlloc(CUdeviceptr* p, size_t n) { COUNT(cuMemAlloc); *p = g_bump; g_bump += (n + 511) & ~(uintptr_t)511; return 0; }
static CUresult s_cuMemGetInfo(size_t* f, size_t* t) { COUNT(cuMemGetInfo); *f = (size_t)170 << 30; *t = (size_t)180 << 30; return 0; }
static CUresult s_cuMemHostAlloc(void** p, size_t n, unsigned f) { COUNT(cuMemHostAlloc); (void)f; *p = calloc(1, n ? n : 1); return *p ? 0 : CUDA_ERROR_OUT_OF_MEMORY; }
static CUresult s_cuMemFreeHost(void* p) { COUNT(cuMemFreeHost); free(p); return 0; }
static CUresult s_cuMemHostGetDevicePointer(CUdeviceptr* d, void* p, unsigned f) { COUNT(cuMemHostGetDevicePointer); (void)f; *d = (CUdeviceptr)(uintptr_t)p; return 0; }
static CUresult s_cuStreamCreate(CUstream* s, unsigned f) { COUNT(cuStreamCreate); (void)f; g_handle += 16; *s = (CUstream)g_handle; return 0; }
static CUresult s_cuEventCreate(CUevent* e, unsigned f) { COUNT(cuEventCreate); (void)f; g_handle += 16; *e = (CUevent)g_handle; return 0; }
static CUresult s_cuEventElapsedTime(float* ms, CUevent a, CUevent b) { COUNT(cuEventElapsedTime); (void)a, (void)b; *ms = 1.0f; return 0; }
static CUresult s_cuModuleLoadData(CUmodule* m, const void* img) { COUNT(cuModuleLoadData); (void)img; g_handle += 16; *m = (CUmodule)g_handle; return 0; }
static CUresult s_cuModuleGetFunction(CUfunction* f, CUmodule m, const char* n) { COUNT(cuModuleGetFunction); (void)m, (void)n; g_handle += 16; *f = (CUfunction)g_handle; return 0; }
/* a device-to-host copy delivers a recognisable pattern (42.0f in every float), so that a test can tell that the read-back landed in
 * the caller's memory, whole and at the right place; nothing is computed */
static CUresult s_cuMemcpyDtoHAsync(void* dst, CUdeviceptr src, size_t bytes, CUstream s) {
  COUNT(cuMemcpyDtoHAsync);
  (void)src, (void)s;
  float* f = (float*)dst;
  for (size_t i = 0; i < bytes / 4; ++i) f[i] = 42.0f;
  return 0;
}
static CUresult s_cuLaunchHostFunc(CUstream s, CUhostFn fn, void* u) { COUNT(cuLaunchHostFunc); (void)s; fn(u); return 0; }
static CUresult s_cuGetErrorString(CUresult r, const char** s) { *s = r == CUDA_ERROR_OUT_OF_MEMORY ? "out of memory (injected)" : "injected by the driver spy"; return 0; }
static CUresult s_cuGetErrorName(CUresult r, const char** s) { *s = r == CUDA_ERROR_OUT_OF_MEMORY ? "CUDA_ERROR_OUT_OF_MEMORY" : "CUDA_ERROR_INJECTED"; return 0; }
static CUresult s_cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned smem, CUstream s,
                                 void** params, void** extra) {
  COUNT(cuLaunchKernel);
  (void)f, (void)gx, (void)gy, (void)gz, (void)bx, (void)by, (void)bz, (void)smem, (void)params, (void)extra;
  ++g_plain_launches;
  saw_stream(s);
  return 0;
}
static CUresult s_cuLaunchKernelEx(const CUlaunchConfig* cfg, CUfunction f, void** params, void** extra) {
  COUNT(cuLaunchKernelEx);
  (void)f, (void)params, (void)extra;
  int pdl = 0;
  for (unsigned i = 0; i < cfg->numAttrs; ++i)
    if (cfg->attrs[i].id == CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION && cfg->attrs[i].value.programmaticStreamSerializationAllowed) pdl = 1;
  if (pdl) ++g_pdl_launches; else ++g_plain_launches;
  saw_stream(cfg->hStream);
  return 0;
}

static CUresult s_cuStreamEndCapture(CUstream s, CUgraph* g) { COUNT(cuStreamEndCapture); (void)s; *g = (CUgraph)(g_handle += 16); return 0; }
static CUresult s_cuGraphInstantiate(CUgraphExec* e, CUgraph g, unsigned long long flags) { COUNT(cuGraphInstantiate); (void)g, (void)flags; *e = (CUgraphExec)(g_handle += 16); return 0; }

/* ---- everything else: counted, succeeds, does nothing. One trampoline per name so that the count knows who was called. ---- */
#define GENERIC_LIST(X) \
  X(cuInit) X(cuDevicePrimaryCtxRelease) X(cuCtxSetCurrent) X(cuCtxSynchronize) X(cuMemFree) X(cuMemcpyHtoDAsync) \
  X(cuMemcpyDtoDAsync) X(cuMemsetD32Async) X(cuStreamDestroy) X(cuStreamSynchronize) X(cuStreamWaitEvent) X(cuEventDestroy) X(cuEventRecord) \
  X(cuEventSynchronize) X(cuEventQuery) X(cuModuleUnload) X(cuFuncSetAttribute) X(cuTensorMapEncodeTiled) X(cuIpcGetMemHandle) \
  X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle) X(cuMemcpyHtoD) X(cuMemcpyDtoH) X(cuStreamBeginCapture) X(cuGraphLaunch) X(cuGraphExecDestroy) \
  X(cuGraphDestroy)
#define GENERIC(name) \
  static CUresult g_##name(void) { \
    static Entry* c_ = NULL; \
    if (!c_) c_ = counter(#name); \
    ++c_->count; \
    return (CUresult)fault(c_); \
  }
GENERIC_LIST(GENERIC)

#define SPECIFIC_LIST(X) \
  X(cuDeviceGetCount) X(cuDeviceGet) X(cuDevicePrimaryCtxRetain) X(cuDeviceGetAttribute) X(cuDeviceTotalMem) X(cuDeviceGetName) X(cuDriverGetVersion) \
  X(cuMemAlloc) X(cuMemGetInfo) X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostGetDevicePointer) X(cuStreamCreate) X(cuEventCreate) \
  X(cuEventElapsedTime) X(cuMemcpyDtoHAsync) X(cuModuleLoadData) X(cuModuleGetFunction) X(cuLaunchHostFunc) X(cuGetErrorString) X(cuGetErrorName) X(cuLaunchKernel) \
  X(cuLaunchKernelEx) X(cuStreamEndCapture) X(cuGraphInstantiate)

CUresult cuGetProcAddress_v2(const char* name, void** fn, int version, cuuint64_t flags, CUdriverProcAddressQueryResult* status) {
  (void)version, (void)flags;
  if (status) *status = CU_GET_PROC_ADDRESS_SUCCESS;
  if (!strcmp(name, "cuGraphInstantiateWithFlags")) name = "cuGraphInstantiate";
#define MATCH_S(n) if (!strcmp(name, #n)) { *fn = (void*)s_##n; return 0; }
#define MATCH_G(n) if (!strcmp(name, #n)) { *fn = (void*)g_##n; return 0; }
  SPECIFIC_LIST(MATCH_S)
  GENERIC_LIST(MATCH_G)
  *fn = NULL; /* an entry point the spy does not know: the runtime reports it as missing instead of calling into nothing */
  return CUDA_ERROR_NOT_FOUND;
}

/* ---- read-back for the tests ---- */
int spy_report(char* out, int capacity) {
  int n = snprintf(out, (size_t)capacity, "{\"pdl_launches\": %ld, \"plain_launches\": %ld, \"streams_launched_on\": %d", g_pdl_launches, g_plain_launches, g_nstreams);
  for (int i = 0; i < g_n && n < capacity; ++i) n += snprintf(out + n, (size_t)(capacity - n), ", \"%s\": %ld", g_counts[i].name, g_counts[i].count);
  if (n < capacity) n += snprintf(out + n, (size_t)(capacity - n), "}");
  return n;
}
/* arm a fault from the test itself: after `skip` more successful calls of `name`, the next `times` calls return `code` */
void spy_arm(const char* name, long skip, int code, long times) {
  Entry* e = counter(name);
  e->skip = e->count + skip;
  e->code = code;
  e->times = times;
}
void spy_reset(void) {  /* counts only: an armed fault keeps counting calls from the process start */
  for (int i = 0; i < g_n; ++i) {
    if (g_counts[i].times > 0) g_counts[i].skip -= g_counts[i].count;
    g_counts[i].count = 0;
  }
  g_pdl_launches = g_plain_launches = 0;
  g_nstreams = 0;
}
