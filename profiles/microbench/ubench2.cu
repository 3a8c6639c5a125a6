// Round-2 microbenchmarks for the gather-add kernel design on B200 (sm_100a).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench2 ubench2.cu
// One CTA per SM; every test prints "warp-adds per clock per SM" (FP32 peak = 4.0, one LDS.128 per 4 adds = 1.0).
//   A. tcgen05.ld.32x32b.xN for N = 4..32 at a warp-uniform dynamic column (TMEM as the gather source)
//   B. TMEM + shared-memory gathers issued by the SAME warp (do the two datapaths add up?)
//   C. tcgen05.cp smem->TMEM (.32x128b.warpx4 and .64x128b.warpx2::02_13): layout check and fill rate,
//      alone and while the compute warps gather
//   D. k-ordered stream with a dynamic accumulator choice (one X row feeds several columns): brx.idx dispatch vs
//      the compare tree ptxas builds for a switch, vs the static per-column lists the product uses
// tmem_ld_gen.h is generated (see git history of this directory); wrappers only.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "tmem_ld_gen.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NSM_MAX = 160;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t tbase) {
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}
__device__ __forceinline__ void tmem_fill_quarter(uint32_t tbase, int warp, int lane) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
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
}

// ------------------------------------------------------------------------------------------------------------
// A. wide TMEM gathers
// ------------------------------------------------------------------------------------------------------------
template <int XN, int DEPTH>
__global__ void __launch_bounds__(512, 1) k_tmem_wide(float *out, long long *cyc, int iters) {
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tbase = tmem_alloc_all(&tslot);
    tmem_fill_quarter(tbase, warp, lane);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float acc[XN];
#pragma unroll
    for (int q = 0; q < XN; ++q) acc[q] = 0.f;
    uint32_t k = warp * 7 + 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[DEPTH][XN];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            k = k * 5 + 3;
            const uint32_t col = ((k >> 4) & (512u / XN - 1u)) * XN;
            const uint32_t ta = tbase + lane_base + col;
            if (XN == 4) tmem_ld_32x32b_x4(ta, reinterpret_cast<uint32_t(&)[4]>(v[u]));
            else if (XN == 8) tmem_ld_32x32b_x8(ta, reinterpret_cast<uint32_t(&)[8]>(v[u]));
            else if (XN == 16) tmem_ld_32x32b_x16(ta, reinterpret_cast<uint32_t(&)[16]>(v[u]));
            else tmem_ld_32x32b_x32(ta, reinterpret_cast<uint32_t(&)[32]>(v[u]));
        }
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < DEPTH; ++u)
#pragma unroll
            for (int q = 0; q < XN; ++q) acc[q] += __uint_as_float(v[u][q]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int q = 0; q < XN; ++q) s += acc[q];
    if (s == 12345.678f) out[2] = s;
    if (lane == 0) atomicMax((unsigned long long *)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
    tmem_free_all(tbase);
}

// ------------------------------------------------------------------------------------------------------------
// B. the same warp gathers NT rows from TMEM (x4) and NS rows from shared memory (LDS.128) per trip
// ------------------------------------------------------------------------------------------------------------
template <int NT, int NS>
__global__ void __launch_bounds__(512, 1) k_mixwarp(float *out, long long *cyc, int iters, int KC) {
    extern __shared__ __align__(16) float xs[];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KC * 128; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    const uint32_t tbase = tmem_alloc_all(&tslot);
    tmem_fill_quarter(tbase, warp, lane);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t xb = smem_u32(xs) + lane * 16;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[j][v] = 0.f;
    uint32_t k = warp * 7 + 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t vt[NT > 0 ? NT : 1][4];
        float4 vs[NS > 0 ? NS : 1];
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            k = k * 5 + 3;
            tmem_ld_32x32b_x4(tbase + lane_base + ((k >> 4) & 127u) * 4, reinterpret_cast<uint32_t(&)[4]>(vt[u]));
        }
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            k = k * 5 + 3;
            const uint32_t a = xb + ((k >> 4) & 127u) * 512u;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vs[u].x), "=f"(vs[u].y), "=f"(vs[u].z), "=f"(vs[u].w) : "r"(a));
        }
        if (NT > 0) tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < NT; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[u & 3][q] += __uint_as_float(vt[u][q]);
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            acc[u & 3][0] += vs[u].x; acc[u & 3][1] += vs[u].y; acc[u & 3][2] += vs[u].z; acc[u & 3][3] += vs[u].w;
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
    tmem_free_all(tbase);
}

// ------------------------------------------------------------------------------------------------------------
// C. tcgen05.cp: K-major X rows (512 B = 32 lanes x 16 B) -> TMEM columns 4k..4k+3 of every lane quarter
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    // UMMA shared-memory matrix descriptor, no swizzle: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void utccp_x4(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void utccp_x2(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mode 0: .32x128b.warpx4 (row k -> all four quarters); mode 1: .64x128b.warpx2::02_13 (rows 2i -> quarters 0,2; 2i+1 -> 1,3)
// res[0] = mismatches, res[1] = fill cycles for KR rows (issue -> mbarrier), res[2] = rows checked
__global__ void __launch_bounds__(128, 1) k_utccp_check(long long *res, int KR, int mode, uint32_t lbo16, uint32_t sbo16) {
    extern __shared__ __align__(1024) float xs1k[]; float *xs = xs1k;
    __shared__ uint32_t tslot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KR * 128; i += blockDim.x) xs[i] = (float)i;  // row k, position p -> k*128+p
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const uint32_t tbase = tmem_alloc_all(&tslot);
    // zero TMEM so stale data cannot pass the check
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 512; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + lane_base + c), "r"(0u));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        if (mode == 0) {
            for (int k = 0; k < KR; ++k) utccp_x4(tbase + 4 * k, make_desc(smem_u32(xs) + k * 512, lbo16, sbo16));
        } else {
            for (int k = 0; k < KR; k += 2) utccp_x2(tbase + 4 * (k / 2), make_desc(smem_u32(xs) + k * 512, lbo16, sbo16));
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;");
    int bad = 0;
    const int slots = (mode == 0) ? KR : KR / 2;
    for (int s = 0; s < slots; ++s) {
        uint32_t r[4];
        tmem_ld_32x32b_x4(tbase + lane_base + 4 * s, r);
        tmem_wait_ld();
        const int k = (mode == 0) ? s : 2 * s + (warp & 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) bad += (__uint_as_float(r[v]) != (float)(k * 128 + lane * 4 + v));
    }
    if (bad) atomicAdd((unsigned long long *)&res[0], (unsigned long long)bad);
    if (threadIdx.x == 0 && blockIdx.x == 0) { res[1] = t1 - t0; res[2] = slots; }
    tmem_free_all(tbase);
}

// fill rate under load: 16 compute warps run the mixed gather of test B while one extra warp keeps copying KR rows per round
// into the half of TMEM the gathers do not touch.  rounds[blockIdx] = completed fill rounds.
template <int NT, int NS>
__global__ void __launch_bounds__(544, 1) k_fill_under_load(float *out, long long *cyc, int *rounds, int iters, int KC, int KR) {
    extern __shared__ __align__(1024) float xs1k[]; float *xs = xs1k;
    __shared__ uint32_t tslot;
    __shared__ __align__(8) uint64_t bar;
    __shared__ int done_warps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KC * 128; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); done_warps = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const uint32_t tbase = tmem_alloc_all(&tslot);
    tmem_fill_quarter(tbase, warp, lane);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
        if (lane == 0) {
            int r = 0;
            while (*(volatile int *)&done_warps < 16) {
                for (int k = 0; k < KR; ++k) utccp_x4(tbase + 256 + 4 * (k & 63), make_desc(smem_u32(xs) + (k % KC) * 512, 1, 8));
                tc_commit(&bar);
                mbar_wait(&bar, r & 1);
                ++r;
            }
            rounds[blockIdx.x] = r;
        }
    } else {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t xb = smem_u32(xs) + lane * 16;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[j][v] = 0.f;
        uint32_t k = warp * 7 + 1;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t vt[NT > 0 ? NT : 1][4];
            float4 vs[NS > 0 ? NS : 1];
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                k = k * 5 + 3;
                tmem_ld_32x32b_x4(tbase + lane_base + ((k >> 4) & 63u) * 4, reinterpret_cast<uint32_t(&)[4]>(vt[u]));
            }
#pragma unroll
            for (int u = 0; u < NS; ++u) {
                k = k * 5 + 3;
                const uint32_t a = xb + ((k >> 4) & 127u) * 512u;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vs[u].x), "=f"(vs[u].y), "=f"(vs[u].z), "=f"(vs[u].w) : "r"(a));
            }
            if (NT > 0) tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < NT; ++u)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[u & 3][q] += __uint_as_float(vt[u][q]);
#pragma unroll
            for (int u = 0; u < NS; ++u) {
                acc[u & 3][0] += vs[u].x; acc[u & 3][1] += vs[u].y; acc[u & 3][2] += vs[u].z; acc[u & 3][3] += vs[u].w;
            }
        }
        long long t1 = clock64();
        float s = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) s += acc[j][v];
        if (s == 12345.678f) out[2] = s;
        if (lane == 0) {
            atomicMax((unsigned long long *)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
            atomicAdd(&done_warps, 1);
        }
    }
    tmem_free_all(tbase);
}

// ------------------------------------------------------------------------------------------------------------
// D. k-ordered op stream: 16-bit ops {row:8, flags: bit5 = load the row first, col: bits 0..4 (16 = no-op)}
//    MODE 0: brx.idx dispatch (PTX), MODE 1: C switch (ptxas builds a compare tree), MODE 2: no dispatch (the add always
//    goes to a compile-time column; same loads and op decoding -- the upper bound of this loop shape)
// ------------------------------------------------------------------------------------------------------------
#define ACC_OPS(a) "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]), "+l"(a[8]), "+l"(a[9]), "+l"(a[10]), \
    "+l"(a[11]), "+l"(a[12]), "+l"(a[13]), "+l"(a[14]), "+l"(a[15]), "+l"(a[16]), "+l"(a[17]), "+l"(a[18]), "+l"(a[19]), "+l"(a[20]), "+l"(a[21]),     \
    "+l"(a[22]), "+l"(a[23]), "+l"(a[24]), "+l"(a[25]), "+l"(a[26]), "+l"(a[27]), "+l"(a[28]), "+l"(a[29]), "+l"(a[30]), "+l"(a[31])
#define CASE(j, a0, a1) "L" #j ": add.rn.f32x2 %" #a0 ", %" #a0 ", %32; add.rn.f32x2 %" #a1 ", %" #a1 ", %33; bra.uni DONE;\n"
// one op: %34 = word holding two ops, SHIFT selects; %35 = xbase (shared address of this lane's 16 bytes of row 0)
#define DISPATCH_OP(SHIFT)                                                                                                         \
    asm volatile("{\n"                                                                                                             \
                 ".reg .pred p;\n.reg .b32 op, j, a;\n"                                                                            \
                 "bfe.u32 op, %34, " #SHIFT ", 16;\n"                                                                              \
                 "and.b32 j, op, 31;\n"                                                                                            \
                 "and.b32 a, op, 32;\n"                                                                                            \
                 "setp.ne.u32 p, a, 0;\n"                                                                                          \
                 "shr.u32 a, op, 8;\n"                                                                                             \
                 "mad.lo.u32 a, a, 512, %35;\n"                                                                                    \
                 "@p ld.shared.v2.b64 {%32, %33}, [a];\n"                                                                          \
                 "ts: .branchtargets L0, L1, L2, L3, L4, L5, L6, L7, L8, L9, L10, L11, L12, L13, L14, L15, DONE;\n"                \
                 "brx.idx.uni j, ts;\n" CASE(0, 0, 1) CASE(1, 2, 3) CASE(2, 4, 5) CASE(3, 6, 7) CASE(4, 8, 9) CASE(5, 10, 11)      \
                     CASE(6, 12, 13) CASE(7, 14, 15) CASE(8, 16, 17) CASE(9, 18, 19) CASE(10, 20, 21) CASE(11, 22, 23)             \
                         CASE(12, 24, 25) CASE(13, 26, 27) CASE(14, 28, 29) CASE(15, 30, 31) "DONE:\n"                             \
                 "}\n"                                                                                                             \
                 : ACC_OPS(acc), "+l"(x01), "+l"(x23)                                                                              \
                 : "r"(word), "r"(xbase)                                                                                           \
                 : "memory")

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_dispatch(const uint32_t *__restrict__ ops_g, float *out, long long *cyc, int words_per_warp, int iters) {
    extern __shared__ __align__(16) uint8_t sm[];
    float *xs = reinterpret_cast<float *>(sm);  // 208 rows x 128 floats
    uint32_t *ops = reinterpret_cast<uint32_t *>(sm + 208 * 512);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 208 * 128; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    for (int i = threadIdx.x; i < words_per_warp * 16; i += blockDim.x) ops[i] = ops_g[i];
    __syncthreads();
    unsigned long long acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0ull;
    unsigned long long x01 = 0ull, x23 = 0ull;
    const uint32_t xbase = smem_u32(xs) + lane * 16;
    const uint4 *my = reinterpret_cast<const uint4 *>(ops + warp * words_per_warp);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int q = 0; q < words_per_warp / 4; ++q) {
            const uint4 w4 = my[q];  // uniform LDS.128 = 8 ops
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t word = (h == 0) ? w4.x : (h == 1) ? w4.y : (h == 2) ? w4.z : w4.w;
                if (MODE == 0) {
                    DISPATCH_OP(0);
                    DISPATCH_OP(16);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const uint32_t op = (word >> (16 * e)) & 0xffffu;
                        if (op & 32u) {
                            const uint32_t a = xbase + (op >> 8) * 512u;
                            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x01), "=l"(x23) : "r"(a));
                        }
                        const uint32_t j = op & 31u;
                        if (MODE == 1) {
                            switch (j) {
#define SW(c) case c: asm volatile("add.rn.f32x2 %0, %0, %2; add.rn.f32x2 %1, %1, %3;" : "+l"(acc[2 * c]), "+l"(acc[2 * c + 1]) : "l"(x01), "l"(x23)); break;
                                SW(0) SW(1) SW(2) SW(3) SW(4) SW(5) SW(6) SW(7) SW(8) SW(9) SW(10) SW(11) SW(12) SW(13) SW(14) SW(15)
                                default: break;
                            }
                        } else {
                            if (j != 16u) {  // static column: position in the unrolled body
                                const int c = (2 * h + e);
                                asm volatile("add.rn.f32x2 %0, %0, %2; add.rn.f32x2 %1, %1, %3;" : "+l"(acc[2 * c]), "+l"(acc[2 * c + 1]) : "l"(x01), "l"(x23));
                            }
                        }
                    }
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s ^= acc[j];
    if (s == 0x123456789abcull) out[2] = 1.f;
    if (lane == 0) atomicMax((unsigned long long *)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
}

// ------------------------------------------------------------------------------------------------------------
// E. does an L1-resident global gather (LDG.128, ld.global.ca) add bandwidth on top of the shared-memory gather?
// ------------------------------------------------------------------------------------------------------------
template <int NG, int NS>
__global__ void __launch_bounds__(512, 1) k_ldgmix(const float *__restrict__ xg, float *out, long long *cyc, int iters, int KC, int KG) {
    extern __shared__ __align__(16) float xs[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KC * 128; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    __syncthreads();
    const uint32_t xb = smem_u32(xs) + lane * 16;
    const float *gb = xg + (size_t)blockIdx.x * KG * 128 + lane * 4;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[j][v] = 0.f;
    uint32_t k = warp * 7 + 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float4 vg[NG > 0 ? NG : 1];
        float4 vs[NS > 0 ? NS : 1];
#pragma unroll
        for (int u = 0; u < NG; ++u) {
            k = k * 5 + 3;
            const float *a = gb + (size_t)((k >> 4) % (uint32_t)KG) * 128;
            asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vg[u].x), "=f"(vg[u].y), "=f"(vg[u].z), "=f"(vg[u].w) : "l"(a));
        }
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            k = k * 5 + 3;
            const uint32_t a = xb + ((k >> 4) % (uint32_t)KC) * 512u;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vs[u].x), "=f"(vs[u].y), "=f"(vs[u].z), "=f"(vs[u].w) : "r"(a));
        }
#pragma unroll
        for (int u = 0; u < NG; ++u) {
            acc[u & 3][0] += vg[u].x; acc[u & 3][1] += vg[u].y; acc[u & 3][2] += vg[u].z; acc[u & 3][3] += vg[u].w;
        }
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            acc[u & 3][0] += vs[u].x; acc[u & 3][1] += vs[u].y; acc[u & 3][2] += vs[u].z; acc[u & 3][3] += vs[u].w;
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
}

// ------------------------------------------------------------------------------------------------------------
// F. dense regime: register-resident X strip (8 rows per lane, 2 x LDS.128 per k), warp-uniform predicated FADD2 per
//    (k, column) from an 8-bit mask; does a predicated-off FADD2 cost one issue slot or the pipe's two cycles?
//    THRESH: a mask bit is set when (hash & 255) < THRESH, i.e. per-sign density THRESH/256
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) k_predfadd2(float *out, long long *cyc, int iters, int KC, uint32_t thresh) {
    extern __shared__ __align__(16) float xs[];  // [KC][256]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < KC * 256; i += blockDim.x) xs[i] = (float)(i % 97) * 0.01f;
    __shared__ uint32_t masks[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        uint32_t m = 0;
        for (int j = 0; j < 8; ++j) {
            uint32_t h = (i * 8 + j) * 2654435761u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            m |= ((h & 255u) < thresh ? 1u : 0u) << j;
        }
        masks[i] = m;
    }
    __syncthreads();
    float2 acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[j][v] = make_float2(0.f, 0.f);
    const uint32_t xb = smem_u32(xs) + lane * 32;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t *mp = masks + ((it * 16 + warp * 64) & 2047 & ~15);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const uint32_t m = mp[k];  // uniform address
            float4 a, b;
            const uint32_t addr = xb + ((it * 16 + k) & (KC - 1)) * 1024u;  // KC is a power of two
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(addr));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(addr + 16));
            const float2 x0 = make_float2(a.x, a.y), x1 = make_float2(a.z, a.w), x2 = make_float2(b.x, b.y), x3 = make_float2(b.z, b.w);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (m & (1u << j)) {  // warp-uniform: predicated FADD2s
                    acc[j][0] = __fadd2_rn(acc[j][0], x0);
                    acc[j][1] = __fadd2_rn(acc[j][1], x1);
                    acc[j][2] = __fadd2_rn(acc[j][2], x2);
                    acc[j][3] = __fadd2_rn(acc[j][3], x3);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v) s += acc[j][v].x + acc[j][v].y;
    if (s == 12345.678f) out[2] = s;
    if (lane == 0) atomicMax((unsigned long long *)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static int g_nsm = 148;
static float *d_out;
static long long *d_cyc;

template <typename F>
static void run(const char *name, double warp_adds_per_cta, int smem, F launch, const char *extra = "") {
    CK(cudaMemset(d_cyc, 0, sizeof(long long) * NSM_MAX));
    launch();
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
    printf("{\"test\": \"%s\", \"warp_adds_per_clk_per_sm\": %.3f, \"frac_fp32_peak\": %.3f, \"cycles_avg\": %.0f, \"ms\": %.4f, \"smem_bytes\": %d%s}\n",
           name, warp_adds_per_cta / avg, warp_adds_per_cta / avg / 4.0, avg, ms, smem, extra);
    fflush(stdout);
}

// op stream for one warp: rows 0..kc-1, each of ncols columns holds an entry with probability p
static std::vector<uint32_t> make_stream(int kc, int ncols, double p, uint32_t seed, long long *entries, long long *rows_used) {
    std::vector<uint16_t> ops;
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 33) & 0x7fffffff) / 2147483648.0; };
    for (int k = 0; k < kc; ++k) {
        bool first = true;
        for (int j = 0; j < ncols; ++j) {
            if (rnd() < p) {
                ops.push_back((uint16_t)((k << 8) | (first ? 32 : 0) | j));
                if (first) ++*rows_used;
                first = false;
                ++*entries;
            }
        }
    }
    while (ops.size() % 8) ops.push_back(16);
    std::vector<uint32_t> w(ops.size() / 2);
    for (size_t i = 0; i < w.size(); ++i) w[i] = ops[2 * i] | ((uint32_t)ops[2 * i + 1] << 16);
    return w;
}

int main(int argc, char **argv) {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, g_nsm, p.major, p.minor, p.clockRate);
    CK(cudaMalloc(&d_out, 1024));
    CK(cudaMemset(d_out, 0, 1024));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * NSM_MAX));
    const int it = 5000;
    // E
    if (argc > 1 && argv[1][0] == 'F') {
        const int KC = 64, smem = KC * 1024;
        CK(cudaFuncSetAttribute(k_predfadd2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int iters = 4000;
        for (uint32_t thresh : {13u, 26u, 43u, 64u, 85u, 128u, 256u}) {
            // adds per warp per k: 8 rows x (popcount of the mask); expected popcount = 8 * thresh / 256
            std::vector<uint32_t> hm(2048);
            double bits = 0;
            for (int i = 0; i < 2048; ++i)
                for (int j = 0; j < 8; ++j) { uint32_t h = (i * 8 + j) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; bits += ((h & 255u) < thresh); }
            const double mean_pop = bits / 2048.0;
            char nm[96], extra[96];
            snprintf(nm, 96, "predfadd2_density%.3f", thresh / 256.0);
            snprintf(extra, 96, ", \"mean_popcount_of_8\": %.3f", mean_pop);
            run(nm, (double)iters * 16 * 16 * 8 * mean_pop, smem, [&] { k_predfadd2<<<g_nsm, 512, smem>>>(d_out, d_cyc, iters, KC, thresh); }, extra);
        }
        printf("{\"done\": true}\n");
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'E') {
        const int KG = 96;  // 48 KB per SM in L1
        float *d_xg;
        CK(cudaMalloc(&d_xg, (size_t)g_nsm * KG * 512));
        CK(cudaMemset(d_xg, 0, (size_t)g_nsm * KG * 512));
        auto go = [&](const char *nm, auto kern, int ng, int ns, int KC) {
            const int smem = KC * 512;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            char full[96];
            snprintf(full, 96, "%s_kc%d", nm, KC);
            run(full, (double)it * 4 * (ng + ns) * 16, smem, [&] { kern<<<g_nsm, 512, smem>>>(d_xg, d_out, d_cyc, it, KC, KG); });
        };
        for (int KC : {32, 208}) {
            go("ldgmix_g0_s8", k_ldgmix<0, 8>, 0, 8, KC);
            go("ldgmix_g8_s0", k_ldgmix<8, 0>, 8, 0, KC);
            go("ldgmix_g4_s4", k_ldgmix<4, 4>, 4, 4, KC);
            go("ldgmix_g2_s6", k_ldgmix<2, 6>, 2, 6, KC);
        }
        printf("{\"done\": true}\n");
        return 0;
    }

    // C first: layout checks (cheap, and everything else depends on them)
    {
        long long *d_res;
        CK(cudaMalloc(&d_res, 64));
        const int KR = 128;
        const int smem = KR * 512;
        CK(cudaFuncSetAttribute(k_utccp_check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const uint32_t variants[][2] = {{1, 8}, {0, 8}, {8, 1}, {8, 8}, {32, 8}};
        for (int mode = 0; mode < 2; ++mode)
            for (auto &v : variants) {
                CK(cudaMemset(d_res, 0, 64));
                k_utccp_check<<<1, 128, smem>>>(d_res, KR, mode, v[0], v[1]);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[3] = {-1, -1, -1};
                if (e == cudaSuccess) CK(cudaMemcpy(h, d_res, 24, cudaMemcpyDeviceToHost));
                printf("{\"test\": \"utccp_%s_lbo%u_sbo%u\", \"err\": \"%s\", \"mismatches\": %lld, \"fill_cycles\": %lld, \"slots\": %lld, \"rows\": %d, \"bytes_per_clk\": %.1f}\n",
                       mode == 0 ? "32x128b_warpx4" : "64x128b_warpx2_02_13", v[0], v[1], cudaGetErrorString(e), h[0], h[1], h[2], KR,
                       h[1] > 0 ? (double)KR * 512 / h[1] : 0.0);
                fflush(stdout);
                if (e != cudaSuccess) { printf("{\"fatal\": \"sticky error, stopping\"}\n"); return 1; }
            }
        // whole-chip fill timing with the working variant (all SMs at once)
        for (int mode = 0; mode < 2; ++mode) {
            CK(cudaMemset(d_res, 0, 64));
            k_utccp_check<<<g_nsm, 128, smem>>>(d_res, KR, mode, 1, 8);
            CK(cudaDeviceSynchronize());
            long long h[3];
            CK(cudaMemcpy(h, d_res, 24, cudaMemcpyDeviceToHost));
            printf("{\"test\": \"utccp_allsm_mode%d\", \"mismatches\": %lld, \"fill_cycles\": %lld, \"bytes_per_clk\": %.1f}\n", mode, h[0], h[1], (double)KR * 512 / h[1]);
        }
    }
    // A
    {
        auto go = [&](const char *nm, auto kern, int xn, int depth, int nt) {
            char full[96];
            snprintf(full, 96, "%s_nt%d", nm, nt);
            run(full, (double)it * depth * xn * (nt / 32), 0, [&] { kern<<<g_nsm, nt>>>(d_out, d_cyc, it); });
        };
        for (int nt : {256, 512}) {
            go("tmemw_x4_d8", k_tmem_wide<4, 8>, 4, 8, nt);
            go("tmemw_x4_d4", k_tmem_wide<4, 4>, 4, 4, nt);
            go("tmemw_x4_d2", k_tmem_wide<4, 2>, 4, 2, nt);
            go("tmemw_x4_d1", k_tmem_wide<4, 1>, 4, 1, nt);
            go("tmemw_x8_d4", k_tmem_wide<8, 4>, 8, 4, nt);
            go("tmemw_x8_d2", k_tmem_wide<8, 2>, 8, 2, nt);
            go("tmemw_x16_d2", k_tmem_wide<16, 2>, 16, 2, nt);
            go("tmemw_x16_d1", k_tmem_wide<16, 1>, 16, 1, nt);
            go("tmemw_x32_d1", k_tmem_wide<32, 1>, 32, 1, nt);
        }
    }
    // B
    {
        const int KC = 128;
        const int smem = KC * 512;
        auto go = [&](const char *nm, auto kern, int ntm, int ns) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            run(nm, (double)it * 4 * (ntm + ns) * 16, smem, [&] { kern<<<g_nsm, 512, smem>>>(d_out, d_cyc, it, KC); });
        };
        go("mixwarp_t0_s8", k_mixwarp<0, 8>, 0, 8);
        go("mixwarp_t8_s0", k_mixwarp<8, 0>, 8, 0);
        go("mixwarp_t4_s4", k_mixwarp<4, 4>, 4, 4);
        go("mixwarp_t4_s2", k_mixwarp<4, 2>, 4, 2);
        go("mixwarp_t2_s4", k_mixwarp<2, 4>, 2, 4);
        go("mixwarp_t6_s2", k_mixwarp<6, 2>, 6, 2);
        go("mixwarp_t2_s6", k_mixwarp<2, 6>, 2, 6);
        go("mixwarp_t8_s4", k_mixwarp<8, 4>, 8, 4);
        go("mixwarp_t8_s8", k_mixwarp<8, 8>, 8, 8);
    }
    // C under load
    {
        const int KC = 128;
        const int smem = KC * 512;
        int *d_rounds;
        CK(cudaMalloc(&d_rounds, sizeof(int) * NSM_MAX));
        auto go = [&](const char *nm, auto kern, int ntm, int ns, int KR) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CK(cudaMemset(d_rounds, 0, sizeof(int) * NSM_MAX));
            char extra[128];
            run(nm, (double)it * 4 * (ntm + ns) * 16, smem, [&] { kern<<<g_nsm, 544, smem>>>(d_out, d_cyc, d_rounds, it, KC, KR); });
            int hr[NSM_MAX];
            CK(cudaMemcpy(hr, d_rounds, sizeof(int) * NSM_MAX, cudaMemcpyDeviceToHost));
            long long hc[NSM_MAX];
            CK(cudaMemcpy(hc, d_cyc, sizeof(long long) * NSM_MAX, cudaMemcpyDeviceToHost));
            snprintf(extra, 128, "{\"test\": \"%s_fill\", \"rounds_sm0\": %d, \"rows_per_round\": %d, \"fill_bytes_per_clk\": %.1f}", nm, hr[0], KR,
                     (double)hr[0] * KR * 512 / (double)hc[0]);
            printf("%s\n", extra);
        };
        go("fillload_t0_s8_kr64", k_fill_under_load<0, 8>, 0, 8, 64);
        go("fillload_t4_s4_kr64", k_fill_under_load<4, 4>, 4, 4, 64);
        go("fillload_t8_s0_kr64", k_fill_under_load<8, 0>, 8, 0, 64);
    }
    // D
    {
        const int KC = 208;
        for (double dens : {0.05, 0.10, 0.17, 0.25}) {
            long long entries = 0, rows_used = 0;
            std::vector<std::vector<uint32_t>> st;
            size_t mx = 0;
            for (int w = 0; w < 16; ++w) {
                st.push_back(make_stream(KC, 16, dens, 1000 + w, &entries, &rows_used));
                mx = st.back().size() > mx ? st.back().size() : mx;
            }
            mx = (mx + 3) & ~(size_t)3;
            std::vector<uint32_t> all(mx * 16, 16u | (16u << 16));
            for (int w = 0; w < 16; ++w) std::copy(st[w].begin(), st[w].end(), all.begin() + w * mx);
            uint32_t *d_ops;
            CK(cudaMalloc(&d_ops, all.size() * 4));
            CK(cudaMemcpy(d_ops, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
            const int smem = KC * 512 + (int)all.size() * 4;
            const int iters = 200;
            auto go = [&](const char *nm, auto kern) {
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                char full[96], extra[96];
                snprintf(full, 96, "%s_dens%.2f", nm, dens);
                snprintf(extra, 96, ", \"entries_per_row_load\": %.3f, \"slots_per_entry\": %.3f", (double)entries / rows_used, (double)mx * 2 * 16 / entries);
                run(full, (double)iters * entries * 4, smem, [&] { kern<<<g_nsm, 512, smem>>>(d_ops, d_out, d_cyc, (int)mx, iters); }, extra);
            };
            go("dispatch_brx", k_dispatch<0>);
            go("dispatch_switch", k_dispatch<1>);
            go("dispatch_static", k_dispatch<2>);
            CK(cudaFree(d_ops));
        }
    }
    printf("{\"done\": true}\n");
    return 0;
}
