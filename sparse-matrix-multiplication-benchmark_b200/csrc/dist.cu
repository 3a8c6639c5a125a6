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
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded, else the system one), so the
// library has no link-time NCCL dependency and loads on machines without it.
#include <dlfcn.h>

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

}  // namespace tsg

using namespace tsg;

typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);

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
    int *flag = nullptr;
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
    if (cudaMalloc(&D->flag, 16) != cudaSuccess) {
        g_nccl.CommDestroy(D->comm);
        delete D;
        return set_error(TSG_ENOMEM, "cudaMalloc failed");
    }
    cudaMemset(D->flag, 0, 16);
    *out = D;
    return TSG_OK;
}

static void dist_unmap(tsg_dist *D) {
    for (int p = 0; p < D->world; ++p)
        if (p != D->rank && D->y_peer[p]) cudaIpcCloseMemHandle(D->y_peer[p]);
    memset(D->y_peer, 0, sizeof D->y_peer);
    if (D->y_local) cudaFree(D->y_local);
    D->y_local = nullptr;
    D->y_bytes = 0;
}

void tsg_dist_destroy(tsg_dist *D) {
    if (!D) return;
    cudaDeviceSynchronize();
    dist_unmap(D);
    if (D->flag) cudaFree(D->flag);
    if (D->done) cudaFree(D->done);
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
    // (1) X broadcast
    if (root >= 0 && D->world > 1) TSG_NCCL(g_nccl.Broadcast(X, X, (size_t)M * K, ncclFloat32, root, D->comm, st));
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
