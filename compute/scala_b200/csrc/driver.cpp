// driver.cpp — lazy binding of the CUDA driver API.
#include "driver.h"

#include <dlfcn.h>

#include <mutex>

namespace cc {

Driver& driver() {
  static Driver d;
  return d;
}

void Driver::load() {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (loaded) return;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    fail(CC_ERR_NO_DRIVER, strprintf("cannot load the CUDA driver (libcuda.so.1): %s — this backend has no CPU fallback", dlerror()));
  using GetProc = CUresult (*)(const char*, void**, int, cuuint64_t, CUdriverProcAddressQueryResult*);
  GetProc get = (GetProc)dlsym(h, "cuGetProcAddress_v2");
  if (!get) fail(CC_ERR_NO_DRIVER, "libcuda.so.1 has no cuGetProcAddress_v2 (driver older than CUDA 12)");
#define CC_LOAD(name)                                                                                   \
  {                                                                                                     \
    void* fn = nullptr;                                                                                 \
    CUdriverProcAddressQueryResult st;                                                                  \
    CUresult r = get(base_name(#name).c_str(), &fn, CUDA_VERSION, CU_GET_PROC_ADDRESS_DEFAULT, &st);    \
    if (r != CUDA_SUCCESS || !fn) fail(CC_ERR_NO_DRIVER, strprintf("driver entry point %s missing", #name)); \
    name = (decltype(name))fn;                                                                          \
  }
  // #name is stringified after cuda.h's version macros were applied (cuMemAlloc -> cuMemAlloc_v2): strip the suffix
  auto base_name = [](const char* n) {
    std::string s(n);
    size_t p = s.rfind("_v");
    if (p != std::string::npos && p + 2 < s.size() && isdigit((unsigned char)s[p + 2])) s.resize(p);
    return s;
  };
  CC_DRIVER_FUNCTIONS(CC_LOAD)
#undef CC_LOAD
#define CC_LOAD_OPTIONAL(name)                                                                       \
  {                                                                                                  \
    void* fn = nullptr;                                                                              \
    CUdriverProcAddressQueryResult st;                                                               \
    if (get(base_name(#name).c_str(), &fn, CUDA_VERSION, CU_GET_PROC_ADDRESS_DEFAULT, &st) == CUDA_SUCCESS) name = (decltype(name))fn; \
  }
  CC_DRIVER_OPTIONAL_FUNCTIONS(CC_LOAD_OPTIONAL)
#undef CC_LOAD_OPTIONAL
  loaded = true;
}

void check_cu(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* name = nullptr;
  const char* str = nullptr;
  if (driver().loaded) {
    driver().cuGetErrorName(r, &name);
    driver().cuGetErrorString(r, &str);
  }
  fail(r == CUDA_ERROR_OUT_OF_MEMORY ? CC_ERR_OUT_OF_MEMORY : CC_ERR_CUDA,
       strprintf("%s failed: %s (%d): %s", what, name ? name : "CUDA_ERROR", (int)r, str ? str : ""));
}

}  // namespace cc
