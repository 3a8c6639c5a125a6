// gen.cu -- counter-based input generators and an fp64-accumulating dense check, all on the device.
//
// The generators are bit-identical to the CPU checker's generators (same hash, same integer arithmetic): element i of a tensor with seed s is a pure
// function of (s, i), so the CPU checker and the GPU produce the same tensors without shipping them over PCIe.
// Distributions follow the reference's rands_sparse / rands_dense (dense/utils.h:9-68) and initX
// (SparseGEMM.h:43-50); the reference itself never seeds (std::random_device / time(0)).
#include "tsg_internal.h"
#include "gen_pattern.h"

namespace tsg {

__host__ __device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ int ternary_draw(uint64_t h, uint32_t num, uint32_t den) {
    const uint32_t u = (uint32_t)(h >> 32);
    const uint32_t r = (uint32_t)(((uint64_t)u * den) >> 32);
    if (r >= num) return 0;
    return (h & 1ull) ? -1 : 1;
}

template <typename T>
__global__ void k_gen_ternary(T *W, long long n, uint64_t seed, uint32_t num, uint32_t den) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        W[i] = (T)ternary_draw(hash64(seed, (uint64_t)i), num, den);
}
__global__ void k_gen_ternary_slice(float *W, int K, int N, int col0, int ncols, uint64_t seed, uint32_t num, uint32_t den) {
    const long long total = (long long)K * ncols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / ncols, j = i % ncols;
        W[i] = (float)ternary_draw(hash64(seed, (uint64_t)(k * N + col0 + j)), num, den);
    }
}
__global__ void k_gen_uniform(float *X, long long n, uint64_t seed) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t m24 = (uint32_t)(hash64(seed, (uint64_t)i) >> 40);
        X[i] = __fsub_rn(__fmul_rn((float)m24, 1.0f / 8388608.0f), 1.0f);
    }
}
__global__ void k_gen_intvalued(float *X, long long n, uint64_t seed, int range) {
    const uint32_t span = 2u * (uint32_t)range + 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t u = (uint32_t)(hash64(seed, (uint64_t)i) >> 32);
        X[i] = (float)((int)(((uint64_t)u * span) >> 32) - range);
    }
}

// one thread per output element, double accumulation over the dense W (verification only; not a product kernel)
__global__ void k_verify_dense(const float *__restrict__ X, const float *__restrict__ Wd, const float *__restrict__ B, float a, int use_prelu,
                               const float *__restrict__ Y, int N, int K, long long ldy, int m0, int mrows, double *__restrict__ out2) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double rel = 0.0, ab = 0.0;
    if (e < (long long)mrows * N) {
        const int m = m0 + (int)(e / N), n = (int)(e % N);
        double y = 0.0;
        for (int k = 0; k < K; ++k) y += (double)X[(size_t)m * K + k] * (double)Wd[(size_t)k * N + n];
        y += (double)B[n];
        if (use_prelu && y < 0.0) y *= (double)a;
        ab = fabs((double)Y[(size_t)m * ldy + n] - y);
        rel = ab / fmax(fabs(y), 1.0);
    }
    // block max -> global max (doubles are non-negative: compare as integers)
    for (int d = 16; d > 0; d >>= 1) {
        rel = fmax(rel, __shfl_xor_sync(0xffffffffu, rel, d));
        ab = fmax(ab, __shfl_xor_sync(0xffffffffu, ab, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<unsigned long long *>(out2), (unsigned long long)__double_as_longlong(rel));
        atomicMax(reinterpret_cast<unsigned long long *>(out2) + 1, (unsigned long long)__double_as_longlong(ab));
    }
}

// generateSparseMatrix patterns (gen_pattern.h).  Window pattern: one thread per element.
template <typename T>
__global__ void k_gen_window(T *Wm, int H, int Wd, int nonZero, uint64_t seed) {
    const long long n = (long long)H * Wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(i / Wd), w = (int)(i - (long long)h * Wd);
        Wm[i] = (T)gp_window_value(seed, h, w, Wd, nonZero);
    }
}
// Skewed pattern: one warp per row; the two selection thresholds by bisection over the 64-bit key space, the count of
// keys below the probe spread over the lanes
__device__ __forceinline__ uint64_t kth_key_warp(uint64_t seed, int h, int Wd, int L, int lane) {
    uint64_t lo = 0, hi = ~0ull;
    while (lo < hi) {  // warp-uniform: every lane sees the same count
        const uint64_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int w = lane; w < Wd; w += 32) c += (gp_key(seed, h, w, Wd) <= mid);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= L) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
template <typename T>
__global__ void __launch_bounds__(128) k_gen_skewed(T *Wm, int H, int Wd, int nonZero, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const int h = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (h >= H) return;  // whole warps leave together
    int n_plus, n_minus;
    gp_row_limits(seed, h, Wd, nonZero, &n_plus, &n_minus);
    const uint64_t t_plus = n_plus > 0 ? kth_key_warp(seed, h, Wd, n_plus, lane) : 0ull;
    const uint64_t t_all = n_plus + n_minus > 0 ? kth_key_warp(seed, h, Wd, n_plus + n_minus, lane) : 0ull;
    for (int w = lane; w < Wd; w += 32)
        Wm[(size_t)h * Wd + w] = (T)gp_skew_value(gp_key(seed, h, w, Wd), n_plus, n_minus, t_plus, t_all);
}

static int grid_for(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tsg

using namespace tsg;

template <typename T>
static int gen_sparse_pattern(T *Wm, int H, int Wd, int nonZero, int uniform, uint64_t seed) {
    TSG_TRY(ensure_device());
    if (!Wm || H < 0 || Wd < 0 || nonZero < 1) return set_error(TSG_EINVAL, "tsg_gen_sparse_pattern: bad arguments");
    if (uniform && nonZero < 2) return set_error(TSG_EINVAL, "tsg_gen_sparse_pattern: the window pattern needs nonZero >= 2 (the reference loops forever at 1)");
    if (!uniform && Wd > (1 << 20)) return set_error(TSG_EUNSUPPORTED, "tsg_gen_sparse_pattern: the skewed pattern supports W <= 2^20");
    if (H == 0 || Wd == 0) return TSG_OK;
    if (uniform) {
        k_gen_window<T><<<grid_for((long long)H * Wd), 256, 0, stream()>>>(Wm, H, Wd, nonZero, seed);
        TSG_KERNEL_CHECK("k_gen_window");
    } else {
        k_gen_skewed<T><<<(unsigned)((H + 3) / 4), 128, 0, stream()>>>(Wm, H, Wd, nonZero, seed);
        TSG_KERNEL_CHECK("k_gen_skewed");
    }
    return TSG_OK;
}

extern "C" {

int tsg_gen_ternary_f32(float *W, long long n, uint64_t seed, uint32_t num, uint32_t den) {
    TSG_TRY(ensure_device());
    k_gen_ternary<float><<<grid_for(n), 256, 0, stream()>>>(W, n, seed, num, den);
    TSG_KERNEL_CHECK("k_gen_ternary");
    return TSG_OK;
}
int tsg_gen_ternary_i32(int *W, long long n, uint64_t seed, uint32_t num, uint32_t den) {
    TSG_TRY(ensure_device());
    k_gen_ternary<int><<<grid_for(n), 256, 0, stream()>>>(W, n, seed, num, den);
    TSG_KERNEL_CHECK("k_gen_ternary");
    return TSG_OK;
}
int tsg_gen_ternary_slice_f32(float *W, int K, int N, int col0, int ncols, uint64_t seed, uint32_t num, uint32_t den) {
    TSG_TRY(ensure_device());
    k_gen_ternary_slice<<<grid_for((long long)K * ncols), 256, 0, stream()>>>(W, K, N, col0, ncols, seed, num, den);
    TSG_KERNEL_CHECK("k_gen_ternary_slice");
    return TSG_OK;
}
int tsg_gen_uniform_f32(float *X, long long n, uint64_t seed) {
    TSG_TRY(ensure_device());
    k_gen_uniform<<<grid_for(n), 256, 0, stream()>>>(X, n, seed);
    TSG_KERNEL_CHECK("k_gen_uniform");
    return TSG_OK;
}
int tsg_gen_intvalued_f32(float *X, long long n, uint64_t seed, int range) {
    TSG_TRY(ensure_device());
    k_gen_intvalued<<<grid_for(n), 256, 0, stream()>>>(X, n, seed, range);
    TSG_KERNEL_CHECK("k_gen_intvalued");
    return TSG_OK;
}

int tsg_gen_sparse_pattern_i32(int *W, int H, int Wd, int nonZero, int uniform, uint64_t seed) {
    return gen_sparse_pattern<int>(W, H, Wd, nonZero, uniform, seed);
}
int tsg_gen_sparse_pattern_f32(float *W, int H, int Wd, int nonZero, int uniform, uint64_t seed) {
    return gen_sparse_pattern<float>(W, H, Wd, nonZero, uniform, seed);
}

int tsg_verify_dense_f64(const float *X, const float *Wd, const float *B, float a, int use_prelu, const float *Y, int M, int N, int K,
                         long long ldy, int m0, int mrows, double *out2) {
    TSG_TRY(ensure_device());
    if (m0 < 0 || mrows < 0 || m0 + mrows > M) return set_error(TSG_EINVAL, "tsg_verify_dense_f64: row range outside Y");
    double *d = nullptr;
    TSG_TRY(dev_alloc_t(&d, 2));
    TSG_CUDA(cudaMemsetAsync(d, 0, 16, stream()));
    const long long total = (long long)mrows * N;
    if (total > 0) {
        k_verify_dense<<<(unsigned)((total + 255) / 256), 256, 0, stream()>>>(X, Wd, B, a, use_prelu, Y, N, K, ldy, m0, mrows, d);
        TSG_KERNEL_CHECK("k_verify_dense");
    }
    TSG_CUDA(cudaMemcpyAsync(out2, d, 16, cudaMemcpyDeviceToHost, stream()));
    TSG_CUDA(cudaStreamSynchronize(stream()));
    return dev_free(d);
}

}  // extern "C"
