// driver.h — the CUDA driver API, resolved at cc_init through dlopen("libcuda.so.1") + cuGetProcAddress so that the
// library itself loads on machines without a driver (symbol / compile-only tests) and fails loudly at cc_init there.
// This is the layer that replaces LWJGL's OpenCL binding (OpenCL.scala:55-142).
#pragma once
#include <cuda.h>

#include "common.h"

namespace cc {

#define CC_DRIVER_FUNCTIONS(X)      \
  X(cuInit)                         \
  X(cuDriverGetVersion)             \
  X(cuDeviceGet)                    \
  X(cuDeviceGetCount)               \
  X(cuDeviceGetName)                \
  X(cuDeviceGetAttribute)           \
  X(cuDeviceTotalMem)               \
  X(cuDevicePrimaryCtxRetain)       \
  X(cuDevicePrimaryCtxRelease)      \
  X(cuCtxSetCurrent)                \
  X(cuCtxSynchronize)               \
  X(cuMemAlloc)                     \
  X(cuMemFree)                      \
  X(cuMemGetInfo)                   \
  X(cuMemcpyHtoDAsync)              \
  X(cuMemcpyDtoHAsync)              \
  X(cuMemcpyDtoDAsync)              \
  X(cuMemsetD32Async)               \
  X(cuMemHostAlloc)                 \
  X(cuMemFreeHost)                  \
  X(cuMemHostGetDevicePointer)      \
  X(cuStreamCreate)                 \
  X(cuStreamDestroy)                \
  X(cuStreamSynchronize)            \
  X(cuStreamWaitEvent)              \
  X(cuEventCreate)                  \
  X(cuEventDestroy)                 \
  X(cuEventRecord)                  \
  X(cuEventSynchronize)             \
  X(cuEventQuery)                   \
  X(cuEventElapsedTime)             \
  X(cuModuleLoadData)               \
  X(cuModuleUnload)                 \
  X(cuModuleGetFunction)            \
  X(cuLaunchKernel)                 \
  X(cuLaunchKernelEx)               \
  X(cuFuncSetAttribute)             \
  X(cuLaunchHostFunc)               \
  X(cuGetErrorString)               \
  X(cuGetErrorName)                 \
  X(cuTensorMapEncodeTiled)         \
  X(cuIpcGetMemHandle)              \
  X(cuIpcOpenMemHandle)             \
  X(cuIpcCloseMemHandle)            \
  X(cuMemcpyHtoD)                   \
  X(cuMemcpyDtoH)

// resolved if the driver has them; the features built on them (cc_graph_*) report CC_ERR_UNSUPPORTED otherwise
#define CC_DRIVER_OPTIONAL_FUNCTIONS(X) \
  X(cuStreamBeginCapture)               \
  X(cuStreamEndCapture)                 \
  X(cuGraphInstantiateWithFlags)        \
  X(cuGraphLaunch)                      \
  X(cuGraphExecDestroy)                 \
  X(cuGraphDestroy)                     \
  X(cuMemCreate)                        \
  X(cuMemRelease)                       \
  X(cuMemAddressReserve)                \
  X(cuMemAddressFree)                   \
  X(cuMemMap)                           \
  X(cuMemUnmap)                         \
  X(cuMemSetAccess)                     \
  X(cuMemGetAllocationGranularity)      \
  X(cuMemExportToShareableHandle)       \
  X(cuMemImportFromShareableHandle)     \
  X(cuMulticastCreate)                  \
  X(cuMulticastAddDevice)               \
  X(cuMulticastBindMem)                 \
  X(cuMulticastUnbind)                  \
  X(cuMulticastGetGranularity)

struct Driver {
#define CC_DECL(name) decltype(&::name) name = nullptr;
  CC_DRIVER_FUNCTIONS(CC_DECL)
  CC_DRIVER_OPTIONAL_FUNCTIONS(CC_DECL)
#undef CC_DECL
  bool loaded = false;
  void load();  // throws CC_ERR_NO_DRIVER
};

Driver& driver();
void check_cu(CUresult r, const char* what);

#define CC_CU(call) ::cc::check_cu(::cc::driver().call, #call)

}  // namespace cc
