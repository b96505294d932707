"""probe: does this box support NVLink multicast objects (cuMulticast*)? and does NCCL use NVLS?"""
import os, sys
from cuda import cuda as cu
print("cuInit", cu.cuInit(0))
err, n = cu.cuDeviceGetCount(); print("devices", n)
for d in range(n):
    err, dev = cu.cuDeviceGet(d)
    for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED"):
        attr = getattr(cu.CUdevice_attribute, name)
        err, v = cu.cuDeviceGetAttribute(attr, dev)
        print(d, name, err, v)
# try creating a multicast object over all devices
err, ctxs = None, []
prop = cu.CUmulticastObjectProp()
prop.numDevices = n
prop.size = 2 << 20
prop.handleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
prop.flags = 0
err, gran = cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED)
print("granularity", err, gran)
if err == cu.CUresult.CUDA_SUCCESS:
    prop.size = max(prop.size, gran)
    err, dev0 = cu.cuDeviceGet(0)
    err, ctx = cu.cuDevicePrimaryCtxRetain(dev0); cu.cuCtxSetCurrent(ctx)
    err, mc = cu.cuMulticastCreate(prop)
    print("cuMulticastCreate", err)
    if err == cu.CUresult.CUDA_SUCCESS:
        err, fd = cu.cuMemExportToShareableHandle(mc, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)
        print("export fd", err, fd)
        for d in range(n):
            err, dev = cu.cuDeviceGet(d)
            print("add device", d, cu.cuMulticastAddDevice(mc, dev))
