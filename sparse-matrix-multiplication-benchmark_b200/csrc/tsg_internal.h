// tsg_internal.h -- declarations shared by the CUDA translation units of libtsgemm_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>

#include "tsgemm_b200.h"

#define TSG_MAX_PEERS 8

namespace tsg {

// ---- error channel (thread-local message, returned by tsg_last_error / sparse_last_error) ------------------------
int set_error(int code, const char *fmt, ...);
void clear_error();

#define TSG_CUDA(call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return ::tsg::set_error(TSG_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                                    __FILE__, __LINE__);                                                   \
    } while (0)

#define TSG_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != TSG_OK) return rc__; \
    } while (0)

#define TSG_KERNEL_CHECK(name)                                                                              \
    do {                                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                               \
        if (e__ != cudaSuccess)                                                                             \
            return ::tsg::set_error(TSG_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));   \
        ::tsg::count_launch();                                                                              \
    } while (0)

// ---- runtime state ------------------------------------------------------------------------------------------------
cudaStream_t stream();
void count_launch();
int num_sms();
int ensure_device();  // TSG_OK iff a usable sm_100 device is current (cached)

int dev_alloc(void **out, size_t bytes);
int dev_free(void *p);
// workspace that persists across calls on this thread (slot 0: XT tiles, 1: skinny X pack, 2: conversion scratch,
// 3: gather-stream build scratch, 4: scan block sums); stream-safe: a
// user on another stream is ordered behind the previous one with an event
struct Workspace {
    void *p = nullptr;
    size_t cap = 0;
    cudaStream_t last = nullptr;
    cudaEvent_t ev = nullptr;
    bool used = false;
};
int ws_acquire(int slot, size_t bytes, void **out);
int ws_release(int slot);
// scope guard: the slot is released (event recorded for the next cross-stream user) on every exit path
struct WsHold {
    int slot;
    bool held = false;
    explicit WsHold(int s) : slot(s) {}
    WsHold(const WsHold &) = delete;
    WsHold &operator=(const WsHold &) = delete;
    int acquire(size_t bytes, void **out) {
        const int rc = ws_acquire(slot, bytes, out);
        held = (rc == TSG_OK);
        return rc;
    }
    int release() {
        if (!held) return TSG_OK;
        held = false;
        return ws_release(slot);
    }
    ~WsHold() { if (held) ws_release(slot); }
};
// function attributes (dynamic shared memory opt-in) are per DEVICE: run `set` once per device and kernel family
int current_device();
template <typename F>
inline int once_per_device(std::atomic<unsigned long long> &done, F set) {
    const unsigned long long bit = 1ull << (current_device() & 63);
    if (done.load(std::memory_order_acquire) & bit) return TSG_OK;
    const int rc = set();  // setting an attribute twice from two threads is harmless
    if (rc == TSG_OK) done.fetch_or(bit, std::memory_order_release);
    return rc;
}
template <typename T>
inline int dev_alloc_t(T **out, size_t count) {
    return dev_alloc(reinterpret_cast<void **>(out), count * sizeof(T));
}

// ---- private K-tiled gather stream of a TCSC matrix (ktformat.cu) --------------------------------------------------
// Plane p = sign * nchunk + chunk holds, for every column, the local row numbers (k - chunk*kc, one byte each, four
// per 32-bit word, 0xFF padding) of that column's +1 (sign 0) / -1 (sign 1) entries that fall into the chunk.
struct KStream {
    int kc = 0;           // rows of X per chunk (<= 255)
    int nchunk = 0;       // ceil(K / kc)
    int ncols_pad = 0;    // N rounded up to a multiple of 256
    int ngroup = 0;       // ncols_pad / 8 (offset granule: 8 columns)
    uint8_t *cnt = nullptr;    // [2*nchunk][ncols_pad]   words per column
    uint32_t *woff = nullptr;  // [2*nchunk][ngroup + 4]  word offset (into body) of each 8-column group, 4-word aligned
    uint32_t *body = nullptr;  // packed words
    long long body_words = 0;  // capacity
    int max_tile_words = 0;    // max over planes and 256-column tiles of the tile's body words
    int smem_reserved = 0;     // shared memory left out of the ring budget when kc was chosen (0 = the whole 227 KB)
    int nstage = 2;            // ring depth kc was chosen for
    bool built = false;
};

// ---- private chunk-major stream of a BCSR matrix (gemm_bcsr_ring.cu) -------------------------------------------------
// An entry is one block ROW (1 x c values at one k), so an r x c block contributes r consecutive entries.  Run (tile,
// chunk) = the entries of one 256-output-column tile whose k falls into chunk `chunk` (kcb block-rows = kcb*r rows of X),
// ordered by (block-column, k) and padded to a multiple of 16 entries, so that a run's local k (`hdr`, one byte each)
// and its values (`val`, c floats each) are each ONE contiguous, 16-byte aligned span.
struct BStream {
    int kcb = 0;      // block-rows per chunk (kcb * r <= 224 rows of X, kcb <= 255)
    int nchunk = 0;   // ceil(br / kcb)
    int ntile = 0;    // ceil(bc / tbc)
    int tbc = 0;      // block-columns per tile = 256 / c
    uint8_t *cnt = nullptr;      // [ntile*nchunk][256]  entries per block-column of the tile
    uint32_t *wstart = nullptr;  // [ntile*nchunk][16]   first entry (relative to the run) of each compute warp
    uint32_t *eoff = nullptr;    // [ntile*nchunk + 1]   first entry of each run
    uint8_t *hdr = nullptr;      // [entries]            row of X inside the chunk (k - chunk*kcb*r)
    float *val = nullptr;        // [entries][c]
    long long entries = 0;
    int max_run = 0;             // longest run (entries, padded)
    bool built = false, unsupported = false;
};

}  // namespace tsg

// opaque handle of tsgemm_b200.h
struct tsg_tcsc {
    int rows = 0, cols = 0, n_pos = 0, n_neg = 0;
    int *csp = nullptr, *csn = nullptr, *rip = nullptr, *rin = nullptr;  // device
    tsg::KStream ks;       // exact orders: one stream slice per pipeline stage
    tsg::KStream ks_fast;  // TSG_ORDER_FAST: the +1 and the -1 slice of a chunk share a stage (shorter chunks)
    float2 *w2 = nullptr;  // TSG_ORDER_FAST, dense regime: W as {w, w} pairs, [tile of 128 columns][k][128] (gemm_dense_fast.cu); built lazily
    std::mutex mu;  // the private stream is built lazily inside the first GEMM: concurrent GEMMs on one handle serialise here
};

struct tsg_bcsr {
    int r = 0, c = 0, br = 0, bc = 0, k = 0;
    int *row_start = nullptr, *col_idx = nullptr;  // device, reference layout (block-row major)
    float *values = nullptr;
    // private column-major mirror (ascending block-row inside each block-column) used by the kernel
    int *cptr = nullptr;   // [bc+1]
    int *crow = nullptr;   // [k] block-row of each block in column-major order
    int *cblk = nullptr;   // [k] index of the block in `values`
    bool col_built = false;
    float *cval = nullptr;  // [k][r*c] block values in column-major block order (decode kernel, decode_bcsr.cu); built lazily
    tsg::BStream bs;
    std::recursive_mutex mu;  // guards the lazy builds (column index, BStream; the second calls the first)
};

namespace tsg {
// exclusive prefix sum over n uint32 (out may not alias in); grand total written to *total_dev (convert.cu)
int scan_exclusive_u32(const uint32_t *in, uint32_t *out, long long n, uint32_t *total_dev);
// classify a pointer: 1 device (or managed), 0 host
int is_device_pointer(const void *p);
int build_kstream(tsg_tcsc *W, int smem_reserved = 0, int areas = 1, int nstage = 2);
int bcsr_build_cols(tsg_bcsr *W);
// ring kernel for BCSR (gemm_bcsr_ring.cu): *handled = 0 when the matrix does not fit its limits (caller falls back
// to the plain kernel); XT = K-major 128-row tiles of X
int bcsr_gemm_ring(tsg_bcsr *W, const float *XT, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy,
                   int *handled);
// progress groups of the tiled kernel (dist.cu mode 2): row tiles [gbound[g], gbound[g+1]) are complete when done[g] == target[g]
struct Progress {
    int ngroups = 0;
    int gbound[9] = {0};
    unsigned int target[8] = {0};
};
// tsg_tcsc_gemm whose epilogue also stores the result into npeer remote copies of Y (fused all-gather, dist.cu);
// fused_tma: 0 = per-lane stores, 1 = TMA bulk stores from a tile that overlays the stage ring, 2 = TMA bulk stores from
// a tile of its own (shorter chunks, but the stores of one unit overlap the gathers of the next), 3 = the staged tile is
// written once with multimem.st to peerY[0] = the NVSwitch multicast mapping of Y (which includes this rank's Y)
int tcsc_gemm_peers(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N, int K,
                    long long ldy, int npeer, float *const *peerY, unsigned int *done, Progress *prog, int fused_tma = 0);
// BCSR decode shape (decode_bcsr.cu): *handled = 0 outside its limits
int bcsr_decode(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled);
// decode shape (M < TSG_SKINNY_M): rows of X in shared memory, warp per column (decode_tcsc.cu)
int tcsc_decode(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled);
// planning override for the next tiled launches of this thread: cut every 256-column tile into `sub` units (0 = automatic)
void set_plan_sub_all(int sub);
// X (M x K row-major) -> XT[ceil(M/128)][K][128] (zero padded rows)
// mtiles_min > ceil(M/128): additional all-zero tiles are written behind the real ones
int transpose_x_tiles(const float *X, float *XT, int M, int K, int mtiles_min = 0);
// dense regime of the opt-in fast order (gemm_dense_fast.cu); *handled = 0 when the matrix is too sparse for it
int tcsc_gemm_dense_fast(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled);
// tsg_profile_enable: CUDA events around the dominant kernel of a call (begin / end) on the launching stream
void profile_mark(bool begin);
}  // namespace tsg
