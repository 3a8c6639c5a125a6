// dist.cu -- column-partitioned multi-GPU path, one process per GPU (BASELINE.json north_star (3)).
//
// W's N columns are split across the ranks (output columns are independent: column n needs col_start_*[n..n+1], its
// index ranges, b[n] and all of X -- tcsc.c:148-163); X is broadcast, each rank computes its M x (N/P) slab, and the
// slabs are all-gathered so every rank ends with the full Y.  Realisations of the exchange (tsgemm_b200.h has the
// caller-side contract of each mode):
//   mode 0  NCCL:  ncclBroadcast(X); the kernel writes a contiguous slab; ncclAllGather of the slabs; a re-layout
//                  kernel interleaves them into row-major Y.
//   mode 1  fused per-lane peer stores: Y lives in a symmetric cudaMalloc buffer whose CUDA-IPC mappings of all peers
//                  are known to the kernel; the GEMM epilogue stores each finished row segment into the local Y and
//                  straight into every peer's Y over NVLink.  A 4-byte ncclAllReduce after the kernel is the
//                  cross-rank completion barrier.
//   mode 2  copy engines: the kernel bumps progress counters; copy streams wait on them (cuStreamWaitValue32) and push
//                  finished row blocks with strided 2-D DMA copies while the kernel still runs.
//   mode 3/4 fused through the TMA engine: every finished 128-row tile is staged in shared memory and written to the
//                  local Y and every peer's Y with bulk async stores (mode 4: the tile has shared memory of its own).
//   mode 5  fused through the NVSwitch: Y (and X) live in symmetric buffers from the virtual-memory API with a multicast
//                  mapping on top; root writes X once with multimem.st and every finished tile is written once to the
//                  multicast Y, the switch replicating it into every rank (default when the fabric supports it).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded, else the system one), so the
// library has no link-time NCCL dependency and loads on machines without it.
#include <cuda.h>
#include <dlfcn.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "tsg_internal.h"

namespace tsg {

// ---- minimal NCCL binding (stable C ABI of NCCL 2.x) -----------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclSum = 0 };

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};
static Nccl g_nccl;

static int load_nccl() {
    if (g_nccl.lib) return TSG_OK;
    const char *cands[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void *h = nullptr;
    for (const char *c : cands)
        if ((h = dlopen(c, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return set_error(TSG_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define TSG_SYM(field, name)                                                            \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                          \
    if (!g_nccl.field) return set_error(TSG_ENCCL, "libnccl lacks %s", name);
    TSG_SYM(GetUniqueId, "ncclGetUniqueId")
    TSG_SYM(CommInitRank, "ncclCommInitRank")
    TSG_SYM(CommDestroy, "ncclCommDestroy")
    TSG_SYM(Broadcast, "ncclBroadcast")
    TSG_SYM(AllGather, "ncclAllGather")
    TSG_SYM(AllReduce, "ncclAllReduce")
    TSG_SYM(GetErrorString, "ncclGetErrorString")
#undef TSG_SYM
    g_nccl.lib = h;
    return TSG_OK;
}

#define TSG_NCCL(call)                                                                                              \
    do {                                                                                                            \
        ncclResult_t r__ = (call);                                                                                  \
        if (r__ != 0) return set_error(TSG_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__));              \
    } while (0)

// slabs [P][M][wmax] -> Y[M][N]; rank p owns columns [c0.v[p], c0.v[p+1]) (tsg_dist_partition, passed in so that the
// kernel never has to invert the dealing rule: ranks with zero columns and ragged last ranks need no special case)
struct ColStarts { int v[TSG_MAX_PEERS + 1]; };
__global__ void k_relayout_slabs(const float *__restrict__ G, float *__restrict__ Y, int M, int N, int world, int wmax, ColStarts c0) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)M * N) return;
    const int m = (int)(e / N), n = (int)(e % N);
    int p = 0;
    while (p + 1 < world && n >= c0.v[p + 1]) ++p;
    Y[e] = G[((size_t)p * M + m) * wmax + (n - c0.v[p])];
}

// root's X -> the multicast mapping of the symmetric X buffer: one read of X, one write that the NVSwitch replicates into every rank
// (U independent 16-byte loads in flight per thread before their stores)
template <int U>
__global__ void __launch_bounds__(512) k_mc_broadcast(const float4 *__restrict__ src, float *__restrict__ mc, long long nvec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * (i + u * stride)), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
    }
    for (; i < nvec; i += stride) {
        const float4 v = __ldg(src + i);
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

__global__ void __launch_bounds__(512) k_mc_broadcast_words(const float *__restrict__ src, float *__restrict__ mc, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(mc + i), "f"(__ldg(src + i)) : "memory");
}

}  // namespace tsg

using namespace tsg;

typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);

// ---- driver entry points of the virtual-memory / multicast API, bound at run time (no link-time libcuda dependency) -----
namespace {
struct DriverApi {
    bool tried = false, ok = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
};
DriverApi g_drv;

bool load_driver_api() {
    if (g_drv.tried) return g_drv.ok;
    g_drv.tried = true;
    bool ok = true;
    auto get = [&](const char *name, void *slot) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { cudaGetLastError(); ok = false; return; }
        *reinterpret_cast<void **>(slot) = fn;
    };
    get("cuMemCreate", &g_drv.MemCreate);
    get("cuMemRelease", &g_drv.MemRelease);
    get("cuMemAddressReserve", &g_drv.MemAddressReserve);
    get("cuMemAddressFree", &g_drv.MemAddressFree);
    get("cuMemMap", &g_drv.MemMap);
    get("cuMemUnmap", &g_drv.MemUnmap);
    get("cuMemSetAccess", &g_drv.MemSetAccess);
    get("cuMemExportToShareableHandle", &g_drv.MemExportToShareableHandle);
    get("cuMemImportFromShareableHandle", &g_drv.MemImportFromShareableHandle);
    get("cuMemGetAllocationGranularity", &g_drv.MemGetAllocationGranularity);
    get("cuMulticastCreate", &g_drv.MulticastCreate);
    get("cuMulticastAddDevice", &g_drv.MulticastAddDevice);
    get("cuMulticastBindMem", &g_drv.MulticastBindMem);
    get("cuMulticastUnbind", &g_drv.MulticastUnbind);
    get("cuMulticastGetGranularity", &g_drv.MulticastGetGranularity);
    get("cuDeviceGet", &g_drv.DeviceGet);
    get("cuDeviceGetAttribute", &g_drv.DeviceGetAttribute);
    g_drv.ok = ok;
    return ok;
}
}  // namespace

// D->flag: ints 0-3 are the barrier / consensus words, bytes [64, 64 + 16 * TSG_MAX_PEERS) the exchange area of vmm_alloc
static const size_t kFlagBytes = 64 + 16 * TSG_MAX_PEERS;

struct tsg_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // mode 2: copy-engine all-gather overlapped with the kernel
    unsigned int *done = nullptr;        // [max row tiles] progress counters written by the kernel
    int done_cap = 0;
    cudaStream_t copy_stream[TSG_MAX_PEERS] = {nullptr};
    cudaEvent_t ev_start = nullptr, ev_copy[TSG_MAX_PEERS] = {nullptr};
    StreamWaitValue32Fn wait_value = nullptr;
    // fused mode: symmetric Y buffer + peer mappings
    float *y_local = nullptr;
    size_t y_bytes = 0;
    float *y_peer[TSG_MAX_PEERS] = {nullptr};
    // symmetric Y from the virtual-memory API with an NVSwitch multicast mapping on top (mode 5): a store to y_mc lands in
    // the Y of EVERY rank, so a finished tile leaves its GPU once instead of world-1 times
    struct SymBuf {  // one symmetric buffer: this rank's physical memory, its multicast mapping, the peers' unicast mappings
        bool vmm = false;
        size_t vmm_size = 0;
        CUmemGenericAllocationHandle mem = 0, mc = 0, peer_mem[TSG_MAX_PEERS] = {0};
        CUdeviceptr va_local = 0, va_mc = 0, va_peer[TSG_MAX_PEERS] = {0};
    };
    SymBuf ysym;
    float *y_mc = nullptr;
    // X broadcast through the switch (mode 5): root writes X once into the multicast mapping of a symmetric X buffer
    SymBuf xsym;
    size_t x_sym_bytes = 0;
    float *x_sym = nullptr, *x_sym_mc = nullptr;
    bool x_sym_failed = false;
    cudaStream_t side = nullptr;  // copy-engine work that runs beside the GEMM kernel
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
    int *flag = nullptr;
    // host-buffer entry point: device copy of X assembled from the ranks' row blocks
    float *x_dev = nullptr;
    size_t x_bytes = 0;
};

extern "C" {

void tsg_dist_partition(int N, int rank, int world, int *col0, int *ncols) {
    // deal columns in granules of 32 (kernel tiles and 16-byte alignment); the last rank takes the ragged remainder
    const int units = N / 32;
    const int base_u = units / world, rem = units % world;
    const int base = base_u * 32, big = base + 32;
    int c0 = (rank < rem) ? rank * big : rem * big + (rank - rem) * base;
    int nc = (rank < rem) ? big : base;
    if (rank == world - 1) nc = N - c0;
    *col0 = c0;
    *ncols = nc;
}

int tsg_dist_unique_id(unsigned char id128[128]) {
    TSG_TRY(load_nccl());
    ncclUniqueId id;
    TSG_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return TSG_OK;
}

int tsg_dist_create(const unsigned char id128[128], int rank, int world, tsg_dist **out) {
    *out = nullptr;
    TSG_TRY(ensure_device());
    TSG_TRY(load_nccl());
    if (world < 1 || world > TSG_MAX_PEERS || rank < 0 || rank >= world) return set_error(TSG_EINVAL, "tsg_dist_create: bad rank/world");
    tsg_dist *D = new (std::nothrow) tsg_dist();
    if (!D) return set_error(TSG_ENOMEM, "out of host memory");
    D->rank = rank;
    D->world = world;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&D->comm, world, id, rank);
    if (r != 0) {
        delete D;
        return set_error(TSG_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    if (cudaMalloc(&D->flag, kFlagBytes) != cudaSuccess) {
        g_nccl.CommDestroy(D->comm);
        delete D;
        return set_error(TSG_ENOMEM, "cudaMalloc failed");
    }
    cudaMemset(D->flag, 0, kFlagBytes);
    *out = D;
    return TSG_OK;
}

// host-side consensus: true iff `ok` is true on every rank (collective; synchronises the current stream)
static bool all_ranks_ok(tsg_dist *D, bool ok) {
    if (D->world == 1) return ok;
    int h[2] = {ok ? 0 : 1, 0};
    cudaStream_t st = stream();
    if (cudaMemcpyAsync(D->flag + 2, &h[0], 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
    if (g_nccl.AllReduce(D->flag + 2, D->flag + 3, 1, ncclInt32, ncclSum, D->comm, st) != 0) return false;
    if (cudaMemcpyAsync(&h[1], D->flag + 3, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    return h[1] == 0;
}

static void vmm_release(tsg_dist::SymBuf &S) {
    if (!g_drv.ok) return;
    for (int p = 0; p < TSG_MAX_PEERS; ++p) {
        if (S.va_peer[p]) { g_drv.MemUnmap(S.va_peer[p], S.vmm_size); g_drv.MemAddressFree(S.va_peer[p], S.vmm_size); }
        if (S.peer_mem[p]) g_drv.MemRelease(S.peer_mem[p]);
        S.va_peer[p] = 0;
        S.peer_mem[p] = 0;
    }
    if (S.va_mc) { g_drv.MemUnmap(S.va_mc, S.vmm_size); g_drv.MemAddressFree(S.va_mc, S.vmm_size); }
    if (S.mc && S.mem) {
        int dev = 0;
        cudaGetDevice(&dev);
        CUdevice cudev;
        if (g_drv.DeviceGet(&cudev, dev) == CUDA_SUCCESS) g_drv.MulticastUnbind(S.mc, cudev, 0, S.vmm_size);
    }
    if (S.va_local) { g_drv.MemUnmap(S.va_local, S.vmm_size); g_drv.MemAddressFree(S.va_local, S.vmm_size); }
    if (S.mc) g_drv.MemRelease(S.mc);
    if (S.mem) g_drv.MemRelease(S.mem);
    S.va_mc = S.va_local = 0;
    S.mc = S.mem = 0;
    S.vmm = false;
    S.vmm_size = 0;
}

static void dist_unmap(tsg_dist *D) {
    if (D->ysym.vmm) {
        vmm_release(D->ysym);
        D->y_mc = nullptr;
        memset(D->y_peer, 0, sizeof D->y_peer);
        D->y_local = nullptr;
        D->y_bytes = 0;
        return;
    }
    for (int p = 0; p < D->world; ++p)
        if (p != D->rank && D->y_peer[p]) cudaIpcCloseMemHandle(D->y_peer[p]);
    memset(D->y_peer, 0, sizeof D->y_peer);
    if (D->y_local) cudaFree(D->y_local);
    D->y_local = nullptr;
    D->y_bytes = 0;
}

#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif
// duplicate file descriptor `fd` of process `pid` into this process (Linux >= 5.6; same user / same PID namespace)
static int steal_fd(int pid, int fd) {
    if (pid == (int)getpid()) return dup(fd);
    const int pidfd = (int)syscall(SYS_pidfd_open, pid, 0);
    if (pidfd < 0) return -1;
    const int got = (int)syscall(SYS_pidfd_getfd, pidfd, fd, 0);
    close(pidfd);
    return got;
}

// Symmetric Y through the virtual-memory API: physical memory per rank (cuMemCreate), mapped locally, bound to ONE
// multicast object shared by all ranks (created by rank 0, passed around as a POSIX file descriptor that the other
// processes duplicate with pidfd_getfd), and every peer's memory mapped for unicast access (modes 1-3 keep working).
// Collective.  Every step ends in a consensus, so either all ranks succeed or all fall back to cudaMalloc + CUDA IPC.
static bool vmm_alloc(tsg_dist *D, tsg_dist::SymBuf &S, size_t bytes, bool want_peers) {
    if (getenv("TSG_DIST_NO_MULTICAST")) return false;  // same on every rank (environment of the launcher)
    bool ok = load_driver_api();
    int dev = 0;
    CUdevice cudev = 0;
    if (ok) ok = cudaGetDevice(&dev) == cudaSuccess && g_drv.DeviceGet(&cudev, dev) == CUDA_SUCCESS;
    if (ok) {
        int mc_ok = 0, fd_ok = 0;
        g_drv.DeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
        g_drv.DeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev);
        ok = mc_ok && fd_ok;
    }
    if (!all_ranks_ok(D, ok)) return false;
    cudaStream_t st = stream();
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof ap);
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = dev;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    mp.numDevices = (unsigned)D->world;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g_alloc = 0, g_mc = 0;
    mp.size = bytes;
    ok = g_drv.MemGetAllocationGranularity(&g_alloc, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
         g_drv.MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS;
    size_t gran = g_alloc > g_mc ? g_alloc : g_mc;
    if (gran == 0) gran = (size_t)2 << 20;
    const size_t size = (bytes + gran - 1) / gran * gran;
    mp.size = size;
    S.vmm_size = size;
    S.vmm = true;  // from here on dist_unmap releases whatever exists
    int my_mem_fd = -1, mc_fd = -1;
    // (1) physical memory + local mapping
    if (ok) ok = g_drv.MemCreate(&S.mem, size, &ap, 0) == CUDA_SUCCESS;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (ok) ok = g_drv.MemAddressReserve(&S.va_local, size, gran, 0, 0) == CUDA_SUCCESS && g_drv.MemMap(S.va_local, size, 0, S.mem, 0) == CUDA_SUCCESS &&
                 g_drv.MemSetAccess(S.va_local, size, &acc, 1) == CUDA_SUCCESS;
    if (ok) ok = g_drv.MemExportToShareableHandle(&my_mem_fd, S.mem, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    // (2) the multicast object: rank 0 creates and exports it
    if (ok && D->rank == 0)
        ok = g_drv.MulticastCreate(&S.mc, &mp) == CUDA_SUCCESS &&
             g_drv.MemExportToShareableHandle(&mc_fd, S.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    // (3) everybody learns everybody's (pid, memory fd, multicast fd)
    struct Card { int pid, mem_fd, mc_fd, ok; };
    Card mine = {(int)getpid(), my_mem_fd, mc_fd, ok ? 1 : 0}, all[TSG_MAX_PEERS];
    memset(all, 0, sizeof all);
    {
        static_assert(sizeof(Card) == 16, "exchange area of D->flag is sized for 16-byte cards");
        Card *all_d = reinterpret_cast<Card *>(reinterpret_cast<char *>(D->flag) + 64);  // no allocation here: every rank must reach the collective
        bool x = cudaMemcpyAsync(all_d + D->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, st) == cudaSuccess;
        x = (g_nccl.AllGather(all_d + D->rank, all_d, sizeof mine, /*ncclChar*/ 0, D->comm, st) == 0) && x;
        if (x) x = cudaMemcpyAsync(all, all_d, sizeof(Card) * D->world, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
        ok = ok && x;
        for (int p = 0; p < D->world; ++p) ok = ok && all[p].ok;
    }
    // (4) import the multicast object, join it, bind my memory, map it
    if (ok && D->rank != 0) {
        const int fd = steal_fd(all[0].pid, all[0].mc_fd);
        ok = fd >= 0 && g_drv.MemImportFromShareableHandle(&S.mc, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
        if (fd >= 0) close(fd);
    }
    if (ok) ok = g_drv.MulticastAddDevice(S.mc, cudev) == CUDA_SUCCESS;
    ok = all_ranks_ok(D, ok);  // every device has joined before anybody binds
    if (ok) ok = g_drv.MulticastBindMem(S.mc, 0, S.mem, 0, size, 0) == CUDA_SUCCESS;
    if (ok) ok = g_drv.MemAddressReserve(&S.va_mc, size, gran, 0, 0) == CUDA_SUCCESS && g_drv.MemMap(S.va_mc, size, 0, S.mc, 0) == CUDA_SUCCESS &&
                 g_drv.MemSetAccess(S.va_mc, size, &acc, 1) == CUDA_SUCCESS;
    // (5) unicast mappings of every peer's memory
    for (int p = 0; ok && want_peers && p < D->world; ++p) {
        if (p == D->rank) continue;
        const int fd = steal_fd(all[p].pid, all[p].mem_fd);
        ok = fd >= 0 && g_drv.MemImportFromShareableHandle(&S.peer_mem[p], (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
        if (fd >= 0) close(fd);
        if (ok) ok = g_drv.MemAddressReserve(&S.va_peer[p], size, gran, 0, 0) == CUDA_SUCCESS && g_drv.MemMap(S.va_peer[p], size, 0, S.peer_mem[p], 0) == CUDA_SUCCESS &&
                     g_drv.MemSetAccess(S.va_peer[p], size, &acc, 1) == CUDA_SUCCESS;
    }
    ok = all_ranks_ok(D, ok);  // also: nobody closes its descriptors before everybody has duplicated them
    if (my_mem_fd >= 0) close(my_mem_fd);
    if (mc_fd >= 0) close(mc_fd);
    if (!ok) {
        vmm_release(S);
        return false;
    }
    return true;
}

static bool vmm_alloc_y(tsg_dist *D, size_t bytes) {
    if (!vmm_alloc(D, D->ysym, bytes, true)) return false;
    D->y_local = reinterpret_cast<float *>(D->ysym.va_local);
    D->y_mc = reinterpret_cast<float *>(D->ysym.va_mc);
    D->y_bytes = bytes;
    for (int p = 0; p < D->world; ++p) D->y_peer[p] = (p == D->rank) ? D->y_local : reinterpret_cast<float *>(D->ysym.va_peer[p]);
    return true;
}

void tsg_dist_destroy(tsg_dist *D) {
    if (!D) return;
    cudaDeviceSynchronize();
    dist_unmap(D);
    if (D->xsym.vmm) vmm_release(D->xsym);
    if (D->side) cudaStreamDestroy(D->side);
    if (D->ev_side[0]) cudaEventDestroy(D->ev_side[0]);
    if (D->ev_side[1]) cudaEventDestroy(D->ev_side[1]);
    if (D->flag) cudaFree(D->flag);
    if (D->done) cudaFree(D->done);
    if (D->x_dev) cudaFree(D->x_dev);
    for (int p = 0; p < TSG_MAX_PEERS; ++p) {
        if (D->copy_stream[p]) cudaStreamDestroy(D->copy_stream[p]);
        if (D->ev_copy[p]) cudaEventDestroy(D->ev_copy[p]);
    }
    if (D->ev_start) cudaEventDestroy(D->ev_start);
    if (D->comm) g_nccl.CommDestroy(D->comm);
    delete D;
}

// Symmetric Y buffer for the fused mode: every rank allocates `bytes` with cudaMalloc and maps all peers' buffers
// through CUDA IPC (handles exchanged with ncclAllGather).  Collective.  Returns this rank's buffer.
int tsg_dist_alloc_y(tsg_dist *D, size_t bytes, float **y_local) {
    *y_local = nullptr;
    if (!D) return set_error(TSG_EINVAL, "null handle");
    cudaStream_t st = stream();
    TSG_CUDA(cudaStreamSynchronize(st));
    dist_unmap(D);
    if (D->world > 1 && vmm_alloc_y(D, bytes)) {  // multicast-capable symmetric buffer (all ranks agree on the outcome)
        *y_local = D->y_local;
        return TSG_OK;
    }
    TSG_CUDA(cudaMalloc(&D->y_local, bytes));
    D->y_bytes = bytes;
    D->y_peer[D->rank] = D->y_local;
    if (D->world > 1) {
        cudaIpcMemHandle_t mine;
        TSG_CUDA(cudaIpcGetMemHandle(&mine, D->y_local));
        cudaIpcMemHandle_t *all_d = nullptr;
        TSG_CUDA(cudaMalloc(&all_d, sizeof(cudaIpcMemHandle_t) * D->world));
        TSG_CUDA(cudaMemcpy(all_d + D->rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
        TSG_NCCL(g_nccl.AllGather(all_d + D->rank, all_d, sizeof mine, /*ncclChar*/ 0, D->comm, st));
        TSG_CUDA(cudaStreamSynchronize(st));
        cudaIpcMemHandle_t all_h[TSG_MAX_PEERS];
        TSG_CUDA(cudaMemcpy(all_h, all_d, sizeof(cudaIpcMemHandle_t) * D->world, cudaMemcpyDeviceToHost));
        cudaFree(all_d);
        for (int p = 0; p < D->world; ++p) {
            if (p == D->rank) continue;
            void *ptr = nullptr;
            TSG_CUDA(cudaIpcOpenMemHandle(&ptr, all_h[p], cudaIpcMemLazyEnablePeerAccess));
            D->y_peer[p] = static_cast<float *>(ptr);
        }
    }
    *y_local = D->y_local;
    return TSG_OK;
}

// diagnostic: the peer mappings of the symmetric Y (entry `rank` is the local buffer)
int tsg_dist_peer_ptrs(tsg_dist *D, void *out[TSG_MAX_PEERS]) {
    if (!D) return set_error(TSG_EINVAL, "null handle");
    for (int p = 0; p < TSG_MAX_PEERS; ++p) out[p] = D->y_peer[p];
    return TSG_OK;
}

int tsg_dist_has_multicast(tsg_dist *D) { return (D && D->y_mc) ? 1 : 0; }

int tsg_dist_barrier(tsg_dist *D) {
    if (!D) return set_error(TSG_EINVAL, "null handle");
    if (D->world > 1) TSG_NCCL(g_nccl.AllReduce(D->flag, D->flag + 1, 1, ncclInt32, ncclSum, D->comm, stream()));
    return TSG_OK;
}

int tsg_dist_gemm(tsg_dist *D, tsg_tcsc *W_local, float *X, int root, const float *B, float a, int use_prelu, int order, float *Y, int M,
                  int N, int K, int mode) {
    if (!D || !W_local) return set_error(TSG_EINVAL, "tsg_dist_gemm: null handle");
    TSG_TRY(ensure_device());
    cudaStream_t st = stream();
    int col0, ncols;
    tsg_dist_partition(N, D->rank, D->world, &col0, &ncols);
    if (W_local->cols != ncols || W_local->rows != K)
        return set_error(TSG_EINVAL, "tsg_dist_gemm: rank %d owns columns [%d,%d) but W_local is %d x %d", D->rank, col0, col0 + ncols,
                         W_local->rows, W_local->cols);
    // TSG_DIST_TRACE=1 (diagnostic): per-call device times of the step's phases on this rank, printed to stderr
    static const bool trace_phases = getenv("TSG_DIST_TRACE") != nullptr;
    struct PhaseTrace {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        cudaStream_t st;
        int rank, mode;
        bool on;
        PhaseTrace(bool on_, cudaStream_t s, int r, int m) : st(s), rank(r), mode(m), on(on_) {
            if (on) for (auto &x : e) cudaEventCreate(&x);
        }
        void mark(int i) { if (on) cudaEventRecord(e[i], st); }
        ~PhaseTrace() {
            if (!on) return;
            cudaEventRecord(e[2], st);
            cudaEventSynchronize(e[2]);
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, e[0], e[1]);
            cudaEventElapsedTime(&b, e[1], e[2]);
            fprintf(stderr, "[tsg_dist phases rank %d mode %d] X broadcast %.3f ms, barrier + GEMM + exchange + barrier %.3f ms\n", rank, mode, a, b);
            for (auto &x : e) cudaEventDestroy(x);
        }
    } trace(trace_phases && D->world > 1, st, D->rank, mode);
    trace.mark(0);
    // (1) X broadcast.  With a multicast-capable fabric (mode 5) root writes X ONCE into the multicast mapping of a symmetric X
    //     buffer and every rank multiplies from its own copy; the copy back into the caller's X ("broadcast in place") runs on
    //     a copy engine beside the GEMM kernel.  Otherwise ncclBroadcast.
    const float *Xuse = X;
    bool copy_back = false;
    if (root >= 0 && D->world > 1) {
        const size_t xbytes = (size_t)M * K * 4;
        // one GPU writes through the multicast mapping at about 500 GB/s, NCCL's pipelined broadcast reaches 700 GB/s but needs
        // 0.35 ms for 64 MB across 8 ranks (tools/bcast_probe.py, dist_breakdown.py): the switch wins on latency, NCCL on size
        const char *max_mb = getenv("TSG_MC_BCAST_MAX_MB");  // read per call: the launcher's environment, the same on every rank
        const size_t mc_bcast_max = (size_t)(max_mb ? atoll(max_mb) : 256) << 20;
        // (the decision may only depend on what every rank knows: sizes and the environment, not this rank's pointer)
        bool mc_bcast = (mode == 5) && D->y_mc && !D->x_sym_failed && !(xbytes & 15) && xbytes <= mc_bcast_max && !getenv("TSG_DIST_NCCL_BCAST");
        if (mc_bcast && D->x_sym_bytes < xbytes) {  // (re)allocate: collective, every rank takes this branch in the same call
            TSG_CUDA(cudaStreamSynchronize(st));
            if (D->xsym.vmm) vmm_release(D->xsym);
            D->x_sym = D->x_sym_mc = nullptr;
            D->x_sym_bytes = 0;
            if (vmm_alloc(D, D->xsym, xbytes, false)) {
                D->x_sym = reinterpret_cast<float *>(D->xsym.va_local);
                D->x_sym_mc = reinterpret_cast<float *>(D->xsym.va_mc);
                D->x_sym_bytes = xbytes;
            } else {
                D->x_sym_failed = true;  // consensus inside vmm_alloc: the same on every rank
                mc_bcast = false;
            }
        }
        if (mc_bcast) {
            if (D->rank == root) {
                static const int unroll = getenv("TSG_MC_BCAST_UNROLL") ? atoi(getenv("TSG_MC_BCAST_UNROLL")) : 1;  // diagnostic
                static const int ctas_per_sm = getenv("TSG_MC_BCAST_CTAS") ? atoi(getenv("TSG_MC_BCAST_CTAS")) : 2;
                const int grid = num_sms() * (ctas_per_sm > 0 ? ctas_per_sm : 2);
                if (reinterpret_cast<uintptr_t>(X) & 15)  // a caller's X that is only float-aligned: word stores
                    k_mc_broadcast_words<<<grid, 512, 0, st>>>(X, D->x_sym_mc, (long long)(xbytes / 4));
                else if (unroll >= 4)
                    k_mc_broadcast<4><<<grid, 512, 0, st>>>(reinterpret_cast<const float4 *>(X), D->x_sym_mc, (long long)(xbytes / 16));
                else
                    k_mc_broadcast<1><<<grid, 512, 0, st>>>(reinterpret_cast<const float4 *>(X), D->x_sym_mc, (long long)(xbytes / 16));
                TSG_KERNEL_CHECK("k_mc_broadcast");
            }
            // ordering: the barrier that opens every exchange mode below follows root's kernel in root's stream, so no rank
            // starts its GEMM before root's stores have been performed
            Xuse = D->x_sym;
            copy_back = (D->rank != root);
        } else {
            TSG_NCCL(g_nccl.Broadcast(X, X, (size_t)M * K, ncclFloat32, root, D->comm, st));
        }
    }
    trace.mark(1);
    if (D->world == 1) return tsg_tcsc_gemm(W_local, X, B, a, use_prelu, order, Y, M, N, K, N);

    if (mode == 1) {
        // (2+3 fused) kernel stores into the local Y and every peer's Y
        if (Y != D->y_local) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 1): Y must be the buffer returned by tsg_dist_alloc_y");
        if ((size_t)M * N * 4 > D->y_bytes) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 1): Y buffer too small");
        float *peers[TSG_MAX_PEERS];
        int np = 0;
        for (int p = 1; p < D->world; ++p) {  // start with the next rank so that the ranks do not all hit the same peer first
            const int q = (D->rank + p) % D->world;
            peers[np++] = D->y_peer[q] + col0;
        }
        // every rank must have finished READING its previous Y before a peer overwrites it
        TSG_TRY(tsg_dist_barrier(D));
        // a failure on this rank must still reach the closing barrier, or the other ranks hang in theirs
        const int rc = (ncols > 0) ? tcsc_gemm_peers(W_local, X, B + col0, a, use_prelu, order, Y + col0, M, ncols, K, N, np, peers, nullptr, nullptr) : TSG_OK;
        const int rc2 = tsg_dist_barrier(D);  // all peers' stores have landed when every rank's kernel has retired
        return rc ? rc : rc2;
    }

    if (mode == 5) {
        // (2+3 fused, multicast) every finished 128-row tile is staged in shared memory and written ONCE, with multimem.st to
        // the NVSwitch multicast mapping of the symmetric Y: the switch replicates it into the Y of every rank (this one
        // included), so a slab leaves its GPU once instead of world-1 times and the epilogue never waits for a drain
        if (Y != D->y_local) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 5): Y must be the buffer returned by tsg_dist_alloc_y");
        if (!D->y_mc) return set_error(TSG_EUNSUPPORTED, "tsg_dist_gemm(mode 5): no multicast mapping (tsg_dist_has_multicast() == 0); use mode 3");
        if ((size_t)M * N * 4 > D->y_bytes) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 5): Y buffer too small");
        if (M < TSG_SKINNY_M) return set_error(TSG_EUNSUPPORTED, "tsg_dist_gemm(mode 5) needs M >= %d", TSG_SKINNY_M);
        float *peers[1] = {D->y_mc + col0};
        TSG_TRY(tsg_dist_barrier(D));  // every rank has finished READING its previous Y (and root's X has arrived)
        // a failure between the two barriers must still reach the closing one, or the other ranks hang in theirs
        auto body = [&]() -> int {
            if (copy_back) {  // the caller's X on the non-root ranks: copy engine, concurrent with the GEMM kernel
                if (!D->side) {
                    TSG_CUDA(cudaStreamCreateWithFlags(&D->side, cudaStreamNonBlocking));
                    TSG_CUDA(cudaEventCreateWithFlags(&D->ev_side[0], cudaEventDisableTiming));
                    TSG_CUDA(cudaEventCreateWithFlags(&D->ev_side[1], cudaEventDisableTiming));
                }
                TSG_CUDA(cudaEventRecord(D->ev_side[0], st));
                TSG_CUDA(cudaStreamWaitEvent(D->side, D->ev_side[0], 0));
                TSG_CUDA(cudaMemcpyAsync(X, D->x_sym, (size_t)M * K * 4, cudaMemcpyDeviceToDevice, D->side));
                TSG_CUDA(cudaEventRecord(D->ev_side[1], D->side));
            }
            // more than 4 ranks: half-width units -- the exchange (7/8 of Y inbound per rank) is as long as the GEMM itself, and
            // it only hides behind it when finished tiles leave in a steady trickle (measured at 8 ranks: 1.40 -> 1.32 ms per
            // step; at 4 ranks 1.209 -> 1.203, within noise)
            set_plan_sub_all(D->world > 4 ? 2 : 0);
            const int rc = (ncols > 0) ? tcsc_gemm_peers(W_local, Xuse, B + col0, a, use_prelu, order, Y + col0, M, ncols, K, N, 1, peers, nullptr, nullptr, 3) : TSG_OK;
            set_plan_sub_all(0);
            if (copy_back) TSG_CUDA(cudaStreamWaitEvent(st, D->ev_side[1], 0));
            return rc;
        };
        const int rc = body();
        const int rc2 = tsg_dist_barrier(D);
        return rc ? rc : rc2;
    }

    if (mode == 3 || mode == 4) {
        // (2+3 fused, TMA) the kernel stages every finished 128-row tile in shared memory and the TMA engine writes each
        // row segment (up to 1 KB) to the local Y and to every peer's Y with bulk async stores: NVLink-sized packets
        // (687 GB/s measured, tools/peer_store_bw.py, vs 145 GB/s for mode 1's per-lane stores), issued while the SMs
        // are already gathering the next tile.  Mode 4 (experimental) gives the staged tile shared memory of its own
        // (shorter K chunks) so that the producer never waits for the stores of the previous unit.
        if (Y != D->y_local) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 3/4): Y must be the buffer returned by tsg_dist_alloc_y");
        if ((size_t)M * N * 4 > D->y_bytes) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 3/4): Y buffer too small");
        if (M < TSG_SKINNY_M) return set_error(TSG_EUNSUPPORTED, "tsg_dist_gemm(mode 3/4) needs M >= %d", TSG_SKINNY_M);
        float *peers[TSG_MAX_PEERS];
        int np = 0;
        for (int p = 1; p < D->world; ++p) peers[np++] = D->y_peer[(D->rank + p) % D->world] + col0;
        TSG_TRY(tsg_dist_barrier(D));  // every rank has finished READING its previous Y
        const int rc = (ncols > 0) ? tcsc_gemm_peers(W_local, X, B + col0, a, use_prelu, order, Y + col0, M, ncols, K, N, np, peers, nullptr, nullptr, mode == 4 ? 2 : 1) : TSG_OK;
        const int rc2 = tsg_dist_barrier(D);
        return rc ? rc : rc2;
    }

    if (mode == 2) {
        // (2) one persistent kernel writes the local slab and bumps a progress counter per 128-row tile;
        // (3) per peer, a copy stream waits (stream memory op, no SM involved) until a block of row tiles is complete
        //     and pushes it with a strided 2-D DMA copy over NVLink -- the all-gather runs on the copy engines while
        //     the gather-add is still in flight.
        if (Y != D->y_local) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 2): Y must be the buffer returned by tsg_dist_alloc_y");
        if ((size_t)M * N * 4 > D->y_bytes) return set_error(TSG_EINVAL, "tsg_dist_gemm(mode 2): Y buffer too small");
        if (M < TSG_SKINNY_M) return set_error(TSG_EUNSUPPORTED, "tsg_dist_gemm(mode 2) needs M >= %d", TSG_SKINNY_M);
        const int mtiles = (M + 127) / 128;
        if (!D->wait_value) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
                return set_error(TSG_EUNSUPPORTED, "cuStreamWaitValue32 is not available");
            D->wait_value = reinterpret_cast<StreamWaitValue32Fn>(fn);
            TSG_CUDA(cudaEventCreateWithFlags(&D->ev_start, cudaEventDisableTiming));
            for (int p = 0; p < D->world; ++p) {
                TSG_CUDA(cudaStreamCreateWithFlags(&D->copy_stream[p], cudaStreamNonBlocking));
                TSG_CUDA(cudaEventCreateWithFlags(&D->ev_copy[p], cudaEventDisableTiming));
            }
        }
        if (!D->done) {
            TSG_CUDA(cudaMalloc(&D->done, sizeof(unsigned int) * 8));
            D->done_cap = 8;
        }
        TSG_CUDA(cudaMemsetAsync(D->done, 0, sizeof(unsigned int) * 8, st));
        TSG_TRY(tsg_dist_barrier(D));  // peers have finished reading their previous Y
        // a failure between the two barriers must still reach the closing one, or the other ranks hang in theirs
        auto pushes = [&]() -> int {
            TSG_CUDA(cudaEventRecord(D->ev_start, st));
            Progress prog;
            const bool trace = getenv("TSG_DIST_TRACE") != nullptr;  // diagnostic: when do the pushes run relative to the kernel?
            cudaEvent_t tr_k0 = nullptr, tr_k1 = nullptr, tr_g[8] = {nullptr};
            if (trace) {
                cudaEventCreate(&tr_k0); cudaEventCreate(&tr_k1);
                for (int g = 0; g < 8; ++g) cudaEventCreate(&tr_g[g]);
                cudaEventRecord(tr_k0, st);
            }
            if (ncols > 0)
                TSG_TRY(tcsc_gemm_peers(W_local, X, B + col0, a, use_prelu, order, Y + col0, M, ncols, K, N, 0, nullptr, D->done, &prog));
            if (trace) cudaEventRecord(tr_k1, st);
            (void)mtiles;
            // A FEW copy streams (peers dealt round-robin): one stream tops out near 350 GB/s when all ranks push at once,
            // while one stream per peer makes the blocking stream-wait memory ops alias onto the same hardware queues and
            // serialise behind one another (both measured at 8 GPUs); TSG_DIST_COPY_STREAMS overrides the default of 2.
            if (ncols > 0) {
                int ncs = 2;
                if (const char *e = getenv("TSG_DIST_COPY_STREAMS")) ncs = atoi(e);
                if (ncs < 1) ncs = 1;
                if (ncs > D->world - 1) ncs = D->world - 1;
                for (int c = 0; c < ncs; ++c) {
                    cudaStream_t cs = D->copy_stream[c];
                    TSG_CUDA(cudaStreamWaitEvent(cs, D->ev_start, 0));
                    for (int g = 0; g < prog.ngroups; ++g) {
                        int rc = D->wait_value(cs, (unsigned long long)(uintptr_t)(D->done + g), prog.target[g], /*CU_STREAM_WAIT_VALUE_GEQ*/ 0);
                        if (rc != 0) return set_error(TSG_ECUDA, "cuStreamWaitValue32 failed (%d)", rc);
                        const int r0 = prog.gbound[g] * 128, r1 = (prog.gbound[g + 1] * 128 < M) ? prog.gbound[g + 1] * 128 : M;
                        for (int p = 1 + c; p < D->world; p += ncs) {
                            const int q = (D->rank + p) % D->world;  // start with the next rank: the ranks do not all hit one peer at once
                            TSG_CUDA(cudaMemcpy2DAsync(D->y_peer[q] + (size_t)r0 * N + col0, (size_t)N * 4, Y + (size_t)r0 * N + col0,
                                                       (size_t)N * 4, (size_t)ncols * 4, (size_t)(r1 - r0), cudaMemcpyDeviceToDevice, cs));
                        }
                        if (trace && c == 0) cudaEventRecord(tr_g[g], cs);
                    }
                    TSG_CUDA(cudaEventRecord(D->ev_copy[c], cs));
                    TSG_CUDA(cudaStreamWaitEvent(st, D->ev_copy[c], 0));
                }
            }
            if (trace) {
                cudaStreamSynchronize(st);
                float tk = 0.f;
                cudaEventElapsedTime(&tk, tr_k0, tr_k1);
                fprintf(stderr, "[tsg_dist trace rank %d] transpose+kernel %.3f ms; pushes of group done at:", D->rank, tk);
                for (int g = 0; g < prog.ngroups; ++g) {
                    float tg = 0.f;
                    cudaEventElapsedTime(&tg, tr_k0, tr_g[g]);
                    fprintf(stderr, " g%d(rows %d..%d)=%.3f", g, prog.gbound[g] * 128, prog.gbound[g + 1] * 128, tg);
                }
                fprintf(stderr, " ms\n");
                cudaEventDestroy(tr_k0); cudaEventDestroy(tr_k1);
                for (int g = 0; g < 8; ++g) cudaEventDestroy(tr_g[g]);
            }
            return TSG_OK;
        };
        const int rc = pushes();
        const int rc2 = tsg_dist_barrier(D);  // every rank's pushes have completed => every Y is whole
        return rc ? rc : rc2;
    }

    // mode 0: (2) contiguous slab, (3) ncclAllGather + re-layout
    int c0_0, w0;
    tsg_dist_partition(N, 0, D->world, &c0_0, &w0);
    int wl, cl;
    tsg_dist_partition(N, D->world - 1, D->world, &cl, &wl);
    const int wmax = (w0 > wl) ? w0 : wl;
    float *G = nullptr;
    TSG_TRY(dev_alloc_t(&G, (size_t)D->world * M * wmax));
    float *slab = G + (size_t)D->rank * M * wmax;
    // a failed local GEMM still takes part in the collective (the other ranks are already in it); its error wins
    const int rc_gemm = (ncols > 0) ? tsg_tcsc_gemm(W_local, X, B + col0, a, use_prelu, order, slab, M, ncols, K, wmax) : TSG_OK;
    TSG_NCCL(g_nccl.AllGather(slab, G, (size_t)M * wmax, ncclFloat32, D->comm, st));
    ColStarts c0;
    for (int p = 0; p <= TSG_MAX_PEERS; ++p) c0.v[p] = N;
    for (int p = 0; p < D->world; ++p) {
        int cp, np_;
        tsg_dist_partition(N, p, D->world, &cp, &np_);
        c0.v[p] = cp;
    }
    const long long total = (long long)M * N;
    k_relayout_slabs<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(G, Y, M, N, D->world, wmax, c0);
    TSG_KERNEL_CHECK("k_relayout_slabs");
    const int rc_free = dev_free(G);
    return rc_gemm ? rc_gemm : rc_free;
}

// ---- host-buffer entry point: every rank moves 1/world of the bytes over its own PCIe link ---------------------------------
int tsg_host_register(void *p, size_t bytes) {
    TSG_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return TSG_OK;
}
int tsg_host_unregister(void *p) {
    TSG_CUDA(cudaHostUnregister(p));
    return TSG_OK;
}

int tsg_dist_gemm_host(tsg_dist *D, tsg_tcsc *W_local, const float *X_host, const float *B_dev, float a, int use_prelu, int order,
                       float *Y_host, int M, int N, int K, int mode) {
    if (!D || !W_local || !X_host || !Y_host) return set_error(TSG_EINVAL, "tsg_dist_gemm_host: null argument");
    if (M <= 0 || N <= 0 || K <= 0) return TSG_OK;
    TSG_TRY(ensure_device());
    if (!D->y_local || (size_t)M * N * 4 > D->y_bytes)
        return set_error(TSG_EINVAL, "tsg_dist_gemm_host: call tsg_dist_alloc_y with at least M*N*4 bytes first");
    cudaStream_t st = stream();
    const int rows_per = (M + D->world - 1) / D->world;
    const int r0 = (D->rank * rows_per < M) ? D->rank * rows_per : M, r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
    const size_t need = (size_t)rows_per * D->world * K * 4;
    if (D->x_bytes < need) {
        TSG_CUDA(cudaStreamSynchronize(st));
        if (D->x_dev) cudaFree(D->x_dev);
        D->x_dev = nullptr;
        D->x_bytes = 0;
        TSG_CUDA(cudaMalloc(&D->x_dev, need));
        D->x_bytes = need;
    }
    // (1) my row block of X over my PCIe link, then the blocks are all-gathered over NVLink (in place: NCCL's in-place
    //     all-gather wants sendbuff == recvbuff + rank * count)
    float *mine = D->x_dev + (size_t)D->rank * rows_per * K;
    if (r1 > r0) TSG_CUDA(cudaMemcpyAsync(mine, X_host + (size_t)r0 * K, (size_t)(r1 - r0) * K * 4, cudaMemcpyHostToDevice, st));
    if (D->world > 1) TSG_NCCL(g_nccl.AllGather(mine, D->x_dev, (size_t)rows_per * K, ncclFloat32, D->comm, st));
    // (2) partitioned GEMM + exchange; every rank ends with the full Y in its symmetric buffer
    const int rc = tsg_dist_gemm(D, W_local, D->x_dev, -1, B_dev, a, use_prelu, order, D->y_local, M, N, K, mode);
    // (3) my row block of the full Y back over my PCIe link
    if (rc == TSG_OK && r1 > r0)
        TSG_CUDA(cudaMemcpyAsync(Y_host + (size_t)r0 * N, D->y_local + (size_t)r0 * N, (size_t)(r1 - r0) * N * 4, cudaMemcpyDeviceToHost, st));
    TSG_CUDA(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"

// =====================================================================================================================
// diagnostic: how fast can SMs write a row-major tile into a PEER's memory over NVLink?
//   mode 0  the access pattern of mode 1's epilogue: every lane stores 64 contiguous bytes of a different row (4 x st.v4)
//   mode 1  whole 1 KB row segments staged in shared memory and written with bulk async stores (TMA engine, UBLKCP)
// Used by tools/peer_store_bw.py; informs the design of a fused all-gather epilogue.
// =====================================================================================================================
namespace tsg {

__global__ void __launch_bounds__(512) k_peer_store_scattered(float *__restrict__ dst, long long ld, int rows, int cols, int iters) {
    // CTA b owns row tile (b % mt), column tile (b / mt); 16 warps x 16 columns, every lane four rows of 64 contiguous bytes (the pattern of mode 1's epilogue)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mtiles = rows / 128, ntiles = cols / 256;
    for (int it = 0; it < iters; ++it)
        for (int u = blockIdx.x; u < mtiles * ntiles; u += gridDim.x) {
            const int mt = u / ntiles, nt = u % ntiles;
            const float4 v = make_float4((float)u, (float)lane, (float)warp, (float)it);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float *row = dst + (size_t)(mt * 128 + lane * 4 + r) * ld + nt * 256 + warp * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(row + j) = v;
            }
        }
}

__global__ void __launch_bounds__(512) k_peer_store_bulk(float *__restrict__ dst, long long ld, int rows, int cols, int iters) {
    extern __shared__ __align__(128) float tile[];  // [128][256] floats = 128 KB
    const int mtiles = rows / 128, ntiles = cols / 256;
    for (int i = threadIdx.x; i < 128 * 256; i += blockDim.x) tile[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int it = 0; it < iters; ++it)
        for (int u = blockIdx.x; u < mtiles * ntiles; u += gridDim.x) {
            const int mt = u / ntiles, nt = u % ntiles;
            if (threadIdx.x < 128) {  // one bulk store of a 1 KB row segment per thread
                const int r = threadIdx.x;
                float *row = dst + (size_t)(mt * 128 + r) * ld + nt * 256;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row),
                             "r"((uint32_t)__cvta_generic_to_shared(tile + r * 256)), "r"(1024)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        }
    if (threadIdx.x < 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace tsg

extern "C" int tsg_dbg_peer_store(float *dst, long long ld, int rows, int cols, int iters, int mode) {
    TSG_TRY(ensure_device());
    if (rows % 128 || cols % 256) return set_error(TSG_EINVAL, "rows %% 128 and cols %% 256 must be 0");
    const int grid = num_sms();
    if (mode == 0) {
        k_peer_store_scattered<<<grid, 512, 0, stream()>>>(dst, ld, rows, cols, iters);
    } else {
        static std::atomic<unsigned long long> attr_done{0};
        TSG_TRY(once_per_device(attr_done, [] {
            TSG_CUDA(cudaFuncSetAttribute(k_peer_store_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 256 * 4));
            return (int)TSG_OK;
        }));
        k_peer_store_bulk<<<grid, 512, 128 * 256 * 4, stream()>>>(dst, ld, rows, cols, iters);
    }
    TSG_KERNEL_CHECK("k_peer_store");
    return TSG_OK;
}
