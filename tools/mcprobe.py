from cuda import cuda
import os
err, = cuda.cuInit(0)
err, n = cuda.cuDeviceGetCount()
print("devices", n)
for d in range(n):
    err, dev = cuda.cuDeviceGet(d)
    err, mc = cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)
    err, vmm = cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev)
    err, fd = cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev)
    err, fab = cuda.cuDeviceGetAttribute(cuda.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev)
    print(d, "multicast", mc, "vmm", vmm, "posix_fd", fd, "fabric", fab)
print(open("/proc/sys/kernel/yama/ptrace_scope").read() if os.path.exists("/proc/sys/kernel/yama/ptrace_scope") else "no yama")
os.system("uname -r; nvidia-smi topo -m | head -12; nvidia-smi -q | grep -i -A3 fabric | head -12")
