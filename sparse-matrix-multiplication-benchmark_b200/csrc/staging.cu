// staging.cu -- host<->device staging for the reference-named entry points (callers hand us host buffers).
#include <cstdlib>
#include <cstring>

#include "tsg_host_shim.h"
#include "tsg_internal.h"

using namespace tsg;

// ---- small and medium host-pointer calls --------------------------------------------------------------------------------
// The reference's own shapes are tiny (main.cpp:258-264: M=1, K=512, N=2048 ...), so a host-pointer call is dominated by
// driver round trips, not by bytes.  Calls whose operands fit kArenaBytes go through a per-thread ARENA that lives across
// calls: X and B are packed into one pinned buffer and reach the device with ONE async copy, the kernel writes Y
// straight into mapped pinned memory (posted PCIe writes, visible after the stream synchronises), so a call costs one
// copy, the kernel launches and one synchronisation -- no pool allocation, no second or third copy.  Larger calls use
// pool allocations and direct copies from/to the caller's buffers.
namespace {

constexpr size_t kArenaBytes = (size_t)1 << 20;      // X + B staged through the arena up to this size
constexpr size_t kArenaOutBytes = (size_t)256 << 10;  // Y written through mapped pinned memory up to this size

struct Arena {
    int device = -1;
    char *pin = nullptr;      // pinned + mapped host memory: [0, kArenaBytes) inputs, [kArenaBytes, +kArenaOutBytes) output
    char *pin_dev = nullptr;  // device alias of `pin`
    char *dev = nullptr;      // device copy of the inputs
};
thread_local Arena g_arena;

int arena_get(Arena **out) {
    int dev = 0;
    TSG_CUDA(cudaGetDevice(&dev));
    Arena &a = g_arena;
    if (a.device != dev) {
        if (a.pin) cudaFreeHost(a.pin);
        if (a.dev) cudaFree(a.dev);
        a = Arena();
        TSG_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&a.pin), kArenaBytes + kArenaOutBytes, cudaHostAllocMapped));
        TSG_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&a.pin_dev), a.pin, 0));
        TSG_CUDA(cudaMalloc(reinterpret_cast<void **>(&a.dev), kArenaBytes));
        a.device = dev;
    }
    *out = &a;
    return TSG_OK;
}

inline size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }

template <typename Gemm>
int staged_gemm(const float *X, const float *B, float *Y, int M, int N, int K, Gemm gemm) {
    TSG_TRY(ensure_device());
    const size_t xb = (size_t)M * K * 4, bb = (size_t)N * 4, yb = (size_t)M * N * 4;
    const bool x_dev = is_device_pointer(X), b_dev = is_device_pointer(B), y_dev = is_device_pointer(Y);
    cudaStream_t st = stream();
    if (x_dev && b_dev && y_dev) return gemm(X, B, Y);  // nothing to stage: asynchronous, like every device-pointer call
    const size_t in_bytes = (x_dev ? 0 : up256(xb)) + (b_dev ? 0 : up256(bb));
    if (in_bytes <= kArenaBytes && (y_dev || yb <= kArenaOutBytes)) {
        Arena *ar = nullptr;
        TSG_TRY(arena_get(&ar));
        // the previous call on this thread synchronised before it returned, so the arena is idle
        size_t off = 0;
        const float *dX = X, *dB = B;
        if (!x_dev) { memcpy(ar->pin + off, X, xb); dX = reinterpret_cast<const float *>(ar->dev + off); off += up256(xb); }
        if (!b_dev) { memcpy(ar->pin + off, B, bb); dB = reinterpret_cast<const float *>(ar->dev + off); off += up256(bb); }
        if (off) TSG_CUDA(cudaMemcpyAsync(ar->dev, ar->pin, off, cudaMemcpyHostToDevice, st));
        float *dY = y_dev ? Y : reinterpret_cast<float *>(ar->pin_dev + kArenaBytes);
        int rc = gemm(dX, dB, dY);
        // host inputs were staged from the caller's memory by memcpy, so X/B may be reused as soon as we return; the
        // arena itself must be idle before the next call, and a host Y must be complete
        cudaError_t e = cudaStreamSynchronize(st);
        if (rc) return rc;
        if (e != cudaSuccess) return set_error(TSG_ECUDA, "staged GEMM failed: %s", cudaGetErrorString(e));
        if (!y_dev) memcpy(Y, ar->pin + kArenaBytes, yb);
        return TSG_OK;
    }
    void *dX = nullptr, *dB = nullptr, *dY = nullptr;
    int ox = 0, ob = 0, oy = 0, rc;
    if ((rc = tsg_shim_stage_in(X, xb, &dX, &ox))) return rc;
    if ((rc = tsg_shim_stage_in(B, bb, &dB, &ob))) { tsg_shim_release(dX, ox); return rc; }
    if ((rc = tsg_shim_stage_out_begin(Y, yb, &dY, &oy))) { tsg_shim_release(dX, ox); tsg_shim_release(dB, ob); return rc; }
    rc = gemm(static_cast<const float *>(dX), static_cast<const float *>(dB), static_cast<float *>(dY));
    if (rc == TSG_OK) rc = tsg_shim_stage_out_end(Y, yb, dY, oy);  // synchronises when Y is a host buffer
    else tsg_shim_release(dY, oy);
    tsg_shim_release(dX, ox);
    tsg_shim_release(dB, ob);
    // a host X or B is read by an asynchronous copy (truly asynchronous when it is pinned): with a device Y nothing has
    // waited for it yet, and the reference's contract is that inputs may be reused once the call returns
    if (rc == TSG_OK && (ox || ob) && !oy) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return set_error(TSG_ECUDA, "staged GEMM failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

constexpr int kMaxSlabs = 72;

// row slabs of the pipelined host-pointer GEMM (see tsg_shim_tcsc_gemm_hostpipe); uniform_rows > 0 forces uniform slabs
int host_slab_schedule(int M, int uniform_rows, int *slab_rows) {
    int nslab = 0;
    if (M < TSG_SKINNY_M || M <= 256) {
        slab_rows[nslab++] = M;
    } else if (uniform_rows > 0) {
        int n = (M + uniform_rows - 1) / uniform_rows;
        if (n > 32) n = 32;
        const int rows = (((M + n - 1) / n) + 127) / 128 * 128;
        for (int m0 = 0; m0 < M; m0 += rows) slab_rows[nslab++] = (M - m0 < rows) ? (M - m0) : rows;
    } else {
        // head: 128, 128, 256, 512 while it stays below half of M; the tail mirrors it; the middle is cut into <= 1024-row slabs
        int head[4], nh = 0, sum = 0;
        for (int sz = 128, i = 0; i < 4; ++i) {
            if (2 * (sum + sz) > M) break;
            head[nh++] = sz;
            sum += sz;
            if (i >= 1) sz *= 2;
        }
        const int mid = M - 2 * sum;  // >= 0; a ragged M leaves its odd rows here
        int cap = 1024;
        while ((mid + cap - 1) / cap + 2 * nh > kMaxSlabs) cap *= 2;
        for (int i = 0; i < nh; ++i) slab_rows[nslab++] = head[i];
        const int nmid = (mid + cap - 1) / cap;
        for (int i = 0, left = mid; i < nmid; ++i) {
            int rows = ((left / (nmid - i)) + 127) / 128 * 128;  // equal parts, rounded up to row tiles
            if (rows > left) rows = left;
            if (rows > 0) slab_rows[nslab++] = rows;
            left -= rows;
        }
        for (int i = nh - 1; i >= 0; --i) slab_rows[nslab++] = head[i];
        // the ragged rest of M (fewer than 128 rows) rides with its left neighbour instead of being a slab of its own
        for (int i = 1; i < nslab; ++i) {
            if (slab_rows[i] < 128) {
                slab_rows[i - 1] += slab_rows[i];
                for (int k = i; k + 1 < nslab; ++k) slab_rows[k] = slab_rows[k + 1];
                --nslab;
                --i;
            }
        }
    }
    return nslab;
}

}  // namespace

extern "C" {

int tsg_shim_is_device(const void *p) { return is_device_pointer(p); }

// diagnostic (tests/test_plan.py): the slab schedule of the pipelined host-pointer GEMM; out must hold 72 ints
int tsg_dbg_host_slabs(int M, int uniform_rows, int *out) { return host_slab_schedule(M, uniform_rows, out); }

int tsg_shim_stage_in(const void *p, size_t bytes, void **dev, int *owned) {
    TSG_TRY(ensure_device());
    if (is_device_pointer(p)) {
        *dev = const_cast<void *>(p);
        *owned = 0;
        return TSG_OK;
    }
    TSG_TRY(dev_alloc(dev, bytes));
    *owned = 1;
    if (bytes) TSG_CUDA(cudaMemcpyAsync(*dev, p, bytes, cudaMemcpyHostToDevice, stream()));
    return TSG_OK;
}

int tsg_shim_stage_out_begin(void *p, size_t bytes, void **dev, int *owned) {
    TSG_TRY(ensure_device());
    if (is_device_pointer(p)) {
        *dev = p;
        *owned = 0;
        return TSG_OK;
    }
    *owned = 1;
    return dev_alloc(dev, bytes);
}

int tsg_shim_stage_out_end(void *p, size_t bytes, void *dev, int owned) {
    if (!owned) return TSG_OK;
    if (bytes) TSG_CUDA(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, stream()));
    TSG_CUDA(cudaStreamSynchronize(stream()));
    return dev_free(dev);
}

int tsg_shim_release(void *dev, int owned) { return owned ? dev_free(dev) : TSG_OK; }

int tsg_shim_tcsc_gemm_staged(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N,
                              int K) {
    return staged_gemm(X, B, Y, M, N, K, [&](const float *dX, const float *dB, float *dY) {
        return tsg_tcsc_gemm(W, dX, dB, a, use_prelu, order, dY, M, N, K, N);
    });
}

int tsg_shim_bcsr_gemm_staged(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K) {
    return staged_gemm(X, B, Y, M, N, K, [&](const float *dX, const float *dB, float *dY) {
        return tsg_bcsr_gemm(W, dX, dB, a, use_prelu, dY, M, N, K, N);
    });
}

// ---- pipelined host-pointer GEMM ----------------------------------------------------------------------------------------
// X and Y live in host memory (pinned or pageable).  Rows are independent, so the call is cut into row slabs and the
// three legs -- H2D of slab i+1, kernel on slab i, D2H of slab i-1 -- run concurrently on three streams; PCIe is full
// duplex, so the wall time approaches max(H2D, D2H) + one slab instead of H2D + kernel + D2H.
int tsg_shim_tcsc_gemm_hostpipe(tsg_tcsc *W, const float *X, const float *B_any, float a, int use_prelu, int order, float *Y, int M,
                                int N, int K) {
    TSG_TRY(ensure_device());
    cudaStream_t user = stream();
    // streams and events live across calls, per thread AND per device (a thread that changes device gets a new set)
    struct Pipe {
        int device = -1;
        cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
        cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_k[3] = {nullptr, nullptr, nullptr}, ev_out[3] = {nullptr, nullptr, nullptr};
    };
    static thread_local Pipe pipe;
    const int dev_now = current_device();
    if (pipe.device != dev_now) {
        pipe = Pipe();  // the old device's handles are left to its context
        TSG_CUDA(cudaStreamCreateWithFlags(&pipe.s_in, cudaStreamNonBlocking));
        TSG_CUDA(cudaStreamCreateWithFlags(&pipe.s_k, cudaStreamNonBlocking));
        TSG_CUDA(cudaStreamCreateWithFlags(&pipe.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) {
            TSG_CUDA(cudaEventCreateWithFlags(&pipe.ev_in[i], cudaEventDisableTiming));
            TSG_CUDA(cudaEventCreateWithFlags(&pipe.ev_k[i], cudaEventDisableTiming));
            TSG_CUDA(cudaEventCreateWithFlags(&pipe.ev_out[i], cudaEventDisableTiming));
        }
        pipe.device = dev_now;
    }
    cudaStream_t s_in = pipe.s_in, s_k = pipe.s_k, s_out = pipe.s_out;
    cudaEvent_t *ev_in = pipe.ev_in, *ev_k = pipe.ev_k, *ev_out = pipe.ev_out;
    TSG_CUDA(cudaStreamSynchronize(user));  // W's mirror may have been built on the user stream
    // Slabs are multiples of 128 rows (the kernel's row tile).  PCIe is the bottleneck of a host-pointer call and the GEMM on a
    // short slab runs at a fraction of its large-M efficiency, so the schedule is a RAMP: 128-row slabs at both ends (the
    // un-overlapped head -- first H2D -- and tail -- last kernel + last D2H -- shrink with the slab) doubling towards
    // 1024-row slabs in the middle (where the kernel must not become the longest leg).  TSG_HOST_SLAB_ROWS=n forces uniform
    // n-row slabs (at most 32 of them), the round-1 schedule.
    int slab_rows[kMaxSlabs];
    const char *env_slab = getenv("TSG_HOST_SLAB_ROWS");
    const int nslab = host_slab_schedule(M, env_slab ? atoi(env_slab) : 0, slab_rows);
    int slab = 0;  // the largest one (buffer size)
    for (int i = 0; i < nslab; ++i) slab = slab_rows[i] > slab ? slab_rows[i] : slab;
    // one call = one kernel family: a short last slab must not drop into the skinny kernel (different summation order)
    const int saved_kernel = tsg_tcsc_get_kernel();
    if (saved_kernel == 0 && M >= TSG_SKINNY_M) tsg_tcsc_set_kernel(1);
    const int nbuf = nslab < 3 ? nslab : 3;
    float *dX[3] = {nullptr, nullptr, nullptr}, *dY[3] = {nullptr, nullptr, nullptr}, *dB = nullptr;
    int rc = TSG_OK;
    tsg_set_stream(s_k);
    int b_owned = 0;
    void *bdev = nullptr;
    if ((rc = tsg_shim_stage_in(B_any, (size_t)N * 4, &bdev, &b_owned))) { tsg_set_stream(user); tsg_tcsc_set_kernel(saved_kernel); return rc; }
    dB = static_cast<float *>(bdev);
    for (int i = 0; i < nbuf && !rc; ++i) {
        rc = dev_alloc_t(&dX[i], (size_t)slab * K);
        if (!rc) rc = dev_alloc_t(&dY[i], (size_t)slab * N);
    }
    if (!rc) rc = build_kstream(W);  // on s_k
    cudaStreamSynchronize(s_k);      // pool allocations above are ordered on s_k; the copy streams use them next
    cudaError_t e = cudaSuccess;
    for (int i = 0, m0 = 0; i < nslab && !rc && e == cudaSuccess; m0 += slab_rows[i], ++i) {
        const int b = i % nbuf;
        const int rows = slab_rows[i];
        if (i >= nbuf) {  // buffer reuse: the kernel that read dX[b] and the copy that drained dY[b] must be done
            e = cudaStreamWaitEvent(s_in, ev_k[b], 0);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, ev_out[b], 0);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(dX[b], X + (size_t)m0 * K, (size_t)rows * K * 4, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaEventRecord(ev_in[b], s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, ev_in[b], 0);
        if (e != cudaSuccess) break;
        rc = tsg_tcsc_gemm(W, dX[b], dB, a, use_prelu, order, dY[b], rows, N, K, N);
        if (rc) break;
        e = cudaEventRecord(ev_k[b], s_k);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_out, ev_k[b], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(Y + (size_t)m0 * N, dY[b], (size_t)rows * N * 4, cudaMemcpyDeviceToHost, s_out);
        if (e == cudaSuccess) e = cudaEventRecord(ev_out[b], s_out);
    }
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_k);
    cudaError_t e2 = cudaStreamSynchronize(s_out);
    for (int i = 0; i < nbuf; ++i) {
        dev_free(dX[i]);
        dev_free(dY[i]);
    }
    tsg_shim_release(dB, b_owned);
    cudaStreamSynchronize(s_k);
    tsg_set_stream(user);
    tsg_tcsc_set_kernel(saved_kernel);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return set_error(TSG_ECUDA, "pipelined GEMM failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return TSG_OK;
}

}  // extern "C"
