"""Does this box expose NVSwitch multicast (NVLS) to user processes?  (cuda-python driver API probe)"""
from cuda import cuda
def ck(r):
    if r[0] != cuda.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
ck(cuda.cuInit(0))
n = ck(cuda.cuDeviceGetCount())
print("devices", n)
for d in range(n):
    dev = ck(cuda.cuDeviceGet(d))
    mc = ck(cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev))
    fab = ck(cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev))
    fd = ck(cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev))
    print("dev", d, "multicast", mc, "fabric", fab, "posix_fd", fd)
dev = ck(cuda.cuDeviceGet(0))
ctx = ck(cuda.cuDevicePrimaryCtxRetain(dev))
ck(cuda.cuCtxSetCurrent(ctx))
prop = cuda.CUmulticastObjectProp()
prop.numDevices = n
prop.size = 64 << 20
prop.handleTypes = cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
prop.flags = 0
try:
    gran = ck(cuda.cuMulticastGetGranularity(prop, cuda.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED))
    print("granularity", gran)
    h = ck(cuda.cuMulticastCreate(prop))
    print("cuMulticastCreate ok", h)
    fdh = ck(cuda.cuMemExportToShareableHandle(h, cuda.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
    print("export fd ok", fdh)
except Exception as e:
    print("multicast create failed:", e)
