// Microbenchmarks that decide the gather-add kernel design on B200 (sm_100a).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench ubench.cu
// Each test runs one CTA per SM (148) x NT threads and reports warp-level "adds per clock per SM"
// (32-lane FADDs retired per SM clock; FP32 peak = 4.0), measured with clock64() inside the kernel (max over CTAs)
// and cross-checked with CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NSM_MAX = 160;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------------------
// A. register-register FADD / FADD2 issue rate
// ------------------------------------------------------------------------------------------------------------
template <int MODE>  // 0: add.f32, 1: add.f32x2 (packed), 2: fma.rn.f32 with 1.0 multiplier
__global__ void k_fadd(float *out, long long *cyc, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float x0 = out[0], x1 = out[1];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x0));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("{ .reg .b64 acc, xx; mov.b64 acc, {%0, %1}; mov.b64 xx, {%2, %3}; add.f32x2 acc, acc, xx; mov.b64 {%0, %1}, acc; }"
                             : "+f"(a[i]), "+f"(a[i + 1]) : "f"(x0), "f"(x1));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x0), "f"(x1));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[2] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------------------------
// B/C. shared-memory gather: Xs[k][TM] k-major, warp-uniform k, each lane loads VEC consecutive floats
//      optional SHFL per gather (index broadcast), optional uniform-address LDS per 8 gathers
// ------------------------------------------------------------------------------------------------------------
template <int VEC, int EXTRA>  // EXTRA: 0 none, 1 one SHFL per gather, 2 one uniform LDS.128 per 8 gathers, 3 PRMT-address from packed bytes
__global__ void k_lds(float *out, long long *cyc, int iters, int KC) {
    extern __shared__ __align__(16) float xs[];
    const int TM = 32 * VEC;
    for (int i = threadIdx.x; i < KC * TM; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    __shared__ __align__(16) uint32_t idxs[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idxs[i] = (i * 2654435761u) >> 8;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[j][v] = 0.f;
    uint32_t k = (threadIdx.x >> 5) * 7 + 1;
    uint32_t carried = lane * 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint4 packed = make_uint4(0, 0, 0, 0);
        if (EXTRA == 2 || EXTRA == 3) packed = *reinterpret_cast<const uint4 *>(&idxs[(it * 4) & 1020]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t kk;
            if (EXTRA == 1) {
                carried = __shfl_sync(0xffffffffu, carried + u, (it + u) & 31);
                kk = (k + (carried & 1)) & 127u;
                k = k * 5 + 3;
            } else if (EXTRA == 2 || EXTRA == 3) {
                uint32_t w = (u < 2) ? packed.x : (u < 4) ? packed.y : (u < 6) ? packed.z : packed.w;
                kk = (w >> ((u & 1) * 16)) & 127u;
            } else {
                k = k * 5 + 3;  // warp-uniform LCG
                kk = (k >> 4) & 127u;
            }
            const float *p = xs + kk * TM + lane * VEC;
            if (VEC == 1) {
                acc[u & 3][0] += *p;
            } else if (VEC == 2) {
                float2 v = *reinterpret_cast<const float2 *>(p);
                acc[u & 3][0] += v.x; acc[u & 3][1] += v.y;
            } else {
                float4 v = *reinterpret_cast<const float4 *>(p);
                acc[u & 3][0] += v.x; acc[u & 3][1] += v.y; acc[u & 3][2] += v.z; acc[u & 3][3] += v.w;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) s += acc[j][v];
    if (s == 12345.678f) out[2] = s + carried;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------------------------
// D/E. TMEM gather: X tile in tensor memory, tcgen05.ld.32x32b.xN at a warp-uniform dynamic column
//      MIX: warps >= NW_TMEM do the shared-memory LDS.128 gather instead (concurrent datapaths?)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, float &a) {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr));
    a = __uint_as_float(r0);
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, float &a, float &b) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
    a = __uint_as_float(r0); b = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, float &a, float &b, float &c, float &d) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2); d = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int XN, int DEPTH>
__global__ void k_tmem(float *out, long long *cyc, int iters, int nw_tmem, int KC_smem) {
    extern __shared__ __align__(16) float xs[];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KC_smem * 128; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // fill TMEM: every warp writes its lane quarter (duplicates across warps with the same quarter are harmless)
    if (warp < 4) {
        for (int c = 0; c < 512; ++c) {
            uint32_t v = __float_as_uint((float)((c * 31 + lane) % 89) * 0.01f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + lane_base + c), "r"(v));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[j][v] = 0.f;
    uint32_t k = warp * 7 + 1;
    long long t0 = clock64();
    if (warp < nw_tmem) {
        for (int it = 0; it < iters; ++it) {
            float v[DEPTH][4];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                k = k * 5 + 3;
                uint32_t col = ((k >> 4) & (512u / XN - 1u)) * XN;
                uint32_t ta = tbase + lane_base + col;
                if (XN == 1) tmem_ld_x1(ta, v[u][0]);
                else if (XN == 2) tmem_ld_x2(ta, v[u][0], v[u][1]);
                else tmem_ld_x4(ta, v[u][0], v[u][1], v[u][2], v[u][3]);
            }
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < DEPTH; ++u)
#pragma unroll
                for (int q = 0; q < XN; ++q) acc[u & 3][q] += v[u][q];
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                k = k * 5 + 3;
                uint32_t kk = (k >> 4) & 127u;
                float4 v = *reinterpret_cast<const float4 *>(xs + kk * 128 + lane * 4);
                acc[u & 3][0] += v.x; acc[u & 3][1] += v.y; acc[u & 3][2] += v.z; acc[u & 3][3] += v.w;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) s += acc[j][v];
    if (s == 12345.678f) out[2] = s;
    if (lane == 0) atomicMax((unsigned long long *)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static int g_nsm = 148;
static float *d_out;
static long long *d_cyc;

template <typename F>
static void run(const char *name, double warp_adds_per_cta, int smem, F launch) {
    CK(cudaMemset(d_cyc, 0, sizeof(long long) * NSM_MAX));
    launch();  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(d_cyc, 0, sizeof(long long) * NSM_MAX));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[NSM_MAX];
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * NSM_MAX, cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0;
    for (int i = 0; i < g_nsm; ++i) { if (h[i] > mx) mx = h[i]; avg += (double)h[i] / g_nsm; }
    printf("{\"test\": \"%s\", \"warp_adds_per_clk_per_sm\": %.3f, \"frac_fp32_peak\": %.3f, \"cycles_avg\": %.0f, \"cycles_max\": %lld, \"ms\": %.4f, \"implied_mhz\": %.0f, \"smem_bytes\": %d}\n",
           name, warp_adds_per_cta / avg, warp_adds_per_cta / avg / 4.0, avg, mx, ms, mx / (ms * 1e3), smem);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, g_nsm, p.major, p.minor, p.clockRate);
    CK(cudaMalloc(&d_out, 1024));
    CK(cudaMemset(d_out, 0, 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * NSM_MAX));
    const int it = 20000;

    // A
    for (int nt : {256, 512, 1024}) {
        char nm[64];
        double wadds = (double)it * 16 * (nt / 32);
        snprintf(nm, 64, "fadd_f32_nt%d", nt);   run(nm, wadds, 0, [&] { k_fadd<0><<<g_nsm, nt>>>(d_out, d_cyc, it); });
        snprintf(nm, 64, "fadd_f32x2_nt%d", nt); run(nm, wadds, 0, [&] { k_fadd<1><<<g_nsm, nt>>>(d_out, d_cyc, it); });
        snprintf(nm, 64, "ffma_f32_nt%d", nt);   run(nm, wadds, 0, [&] { k_fadd<2><<<g_nsm, nt>>>(d_out, d_cyc, it); });
    }
    // B/C
    {
        const int KC = 208;
        auto go = [&](const char *nm, auto kern, int vec, int nt) {
            int smem = KC * 32 * vec * 4;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            char full[96];
            snprintf(full, 96, "%s_nt%d", nm, nt);
            run(full, (double)it / 4 * 8 * vec * (nt / 32), smem, [&] { kern<<<g_nsm, nt, smem>>>(d_out, d_cyc, it / 4, KC); });
        };
        for (int nt : {256, 512, 1024}) {
            go("lds32_gather", k_lds<1, 0>, 1, nt);
            go("lds64_gather", k_lds<2, 0>, 2, nt);
            go("lds128_gather", k_lds<4, 0>, 4, nt);
            go("lds128_gather_shfl", k_lds<4, 1>, 4, nt);
            go("lds64_gather_shfl", k_lds<2, 1>, 2, nt);
            go("lds32_gather_shfl", k_lds<1, 1>, 1, nt);
            go("lds128_gather_uniform_idx", k_lds<4, 2>, 4, nt);
            go("lds64_gather_uniform_idx", k_lds<2, 2>, 2, nt);
        }
    }
    // D: TMEM only
    {
        auto go = [&](const char *nm, auto kern, int xn, int depth, int nt, int nw_tmem, int kc_smem) {
            int smem = kc_smem * 128 * 4 + 16;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            char full[96];
            snprintf(full, 96, "%s_nt%d_tmemwarps%d", nm, nt, nw_tmem);
            int nw = nt / 32;
            int nws = nw - nw_tmem;
            double wadds = (double)(it / 4) * depth * (xn * nw_tmem + 4 * (nws > 0 ? nws : 0));
            run(full, wadds, smem, [&] { kern<<<g_nsm, nt, smem>>>(d_out, d_cyc, it / 4, nw_tmem, kc_smem); });
        };
        for (int nt : {128, 256, 512}) {
            go("tmem_x1_d8", k_tmem<1, 8>, 1, 8, nt, nt / 32, 1);
            go("tmem_x2_d8", k_tmem<2, 8>, 2, 8, nt, nt / 32, 1);
            go("tmem_x4_d8", k_tmem<4, 8>, 4, 8, nt, nt / 32, 1);
            go("tmem_x4_d4", k_tmem<4, 4>, 4, 4, nt, nt / 32, 1);
        }
        // E: mixed -- half the warps gather from TMEM (x4), half from shared memory (LDS.128)
        go("mix_tmem_x4_lds128", k_tmem<4, 8>, 4, 8, 512, 8, 208);
        go("mix_tmem_x4_lds128", k_tmem<4, 8>, 4, 8, 1024, 16, 208);
        go("mix_tmem_x4_lds128", k_tmem<4, 8>, 4, 8, 768, 8, 208);
        go("mix_tmem_x2_lds128", k_tmem<2, 8>, 2, 8, 512, 8, 208);
        go("mix_tmem_x1_lds128", k_tmem<1, 8>, 1, 8, 512, 8, 208);
        go("lds128_only_in_mix_kernel", k_tmem<4, 8>, 4, 8, 512, 0, 208);
    }
    printf("{\"done\": true}\n");
    return 0;
}
