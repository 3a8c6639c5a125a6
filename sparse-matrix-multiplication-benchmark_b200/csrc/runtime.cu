// runtime.cu -- error channel, stream, stream-ordered memory, launch accounting.
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "tsg_internal.h"

namespace tsg {

static thread_local char g_err[512] = "";
static thread_local cudaStream_t g_stream = nullptr;
static thread_local long long g_launches = 0;

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }
cudaStream_t stream() { return g_stream; }
void count_launch() { ++g_launches; }

struct DevInfo {
    int checked = -1;  // device ordinal the info below is for
    int rc = TSG_ENODEV;
    int sms = 0;
};
static thread_local DevInfo g_dev;

int ensure_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(TSG_ENODEV, "no CUDA device: %s (libtsgemm_b200 has no CPU fallback)", cudaGetErrorString(e));
    }
    if (g_dev.checked == dev) {
        if (g_dev.rc != TSG_OK) set_error(g_dev.rc, "device %d is not an sm_100 GPU (libtsgemm_b200 is sm_100a-only)", dev);
        return g_dev.rc;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return set_error(TSG_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    g_dev.checked = dev;
    g_dev.sms = p.multiProcessorCount;
    if (p.major != 10) {
        g_dev.rc = TSG_ENODEV;
        return set_error(TSG_ENODEV, "device %d (%s, sm_%d%d) is not an sm_100 GPU (libtsgemm_b200 is sm_100a-only)", dev,
                         p.name, p.major, p.minor);
    }
    // keep freed blocks in the pool: GEMM workspaces are re-used call after call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    g_dev.rc = TSG_OK;
    return TSG_OK;
}

int current_device() { return g_dev.checked >= 0 ? g_dev.checked : 0; }

int num_sms() { return g_dev.sms > 0 ? g_dev.sms : 148; }

int dev_alloc(void **out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(out, bytes, g_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(TSG_ENOMEM, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    return TSG_OK;
}

// ---- per-thread workspaces that survive across calls (XT / XS): no cudaMallocAsync + cudaFreeAsync per GEMM ------------
static thread_local Workspace g_ws[5];

int ws_acquire(int slot, size_t bytes, void **out) {
    Workspace &w = g_ws[slot];
    if (bytes == 0) bytes = 16;
    if (!w.ev) {
        if (cudaEventCreateWithFlags(&w.ev, cudaEventDisableTiming) != cudaSuccess) return set_error(TSG_ECUDA, "cudaEventCreate failed");
    }
    if (w.used && w.last != g_stream) {  // the previous user ran on another stream: order this one behind it
        if (cudaStreamWaitEvent(g_stream, w.ev, 0) != cudaSuccess) return set_error(TSG_ECUDA, "cudaStreamWaitEvent failed");
    }
    if (w.cap < bytes) {
        if (w.p) cudaFreeAsync(w.p, g_stream);
        w.p = nullptr;
        w.cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMallocAsync(&w.p, want, g_stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(TSG_ENOMEM, "cudaMallocAsync(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        w.cap = want;
    }
    *out = w.p;
    return TSG_OK;
}

int ws_release(int slot) {
    Workspace &w = g_ws[slot];
    w.last = g_stream;
    w.used = true;
    if (cudaEventRecord(w.ev, g_stream) != cudaSuccess) return set_error(TSG_ECUDA, "cudaEventRecord failed");
    return TSG_OK;
}

int dev_free(void *p) {
    if (!p) return TSG_OK;
    cudaError_t e = cudaFreeAsync(p, g_stream);
    if (e != cudaSuccess) return set_error(TSG_ECUDA, "cudaFreeAsync failed: %s", cudaGetErrorString(e));
    return TSG_OK;
}

}  // namespace tsg

extern "C" {

const char *tsg_last_error(void) { return tsg::g_err; }
void tsg_clear_error(void) { tsg::clear_error(); }
const char *sparse_last_error(void) { return tsg::g_err; }
const char *tsg_version(void) { return "tsgemm_b200 0.1 (sm_100a)"; }
int tsg_device_check(void) { return tsg::ensure_device(); }
int tsg_set_stream(void *s) {
    tsg::g_stream = static_cast<cudaStream_t>(s);
    return TSG_OK;
}
void *tsg_get_stream(void) { return tsg::g_stream; }
int tsg_synchronize(void) {
    TSG_CUDA(cudaStreamSynchronize(tsg::g_stream));
    return TSG_OK;
}
long long tsg_launch_count(void) { return tsg::g_launches; }
void tsg_launch_count_reset(void) { tsg::g_launches = 0; }
int tsg_dev_alloc(void **out, size_t bytes) {
    TSG_TRY(tsg::ensure_device());
    return tsg::dev_alloc(out, bytes);
}
int tsg_dev_free(void *p) { return tsg::dev_free(p); }

}  // extern "C"
