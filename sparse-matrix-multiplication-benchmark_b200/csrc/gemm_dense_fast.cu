// gemm_dense_fast.cu -- the dense regime of the opt-in TSG_ORDER_FAST path (sparsity <= ~70 %).
//
// Same math as tcsc_sgemm_prelu_basic (sparse/tcsc.c:143-165) under the tolerance contract of TSG_ORDER_FAST (DESIGN.md
// section 4): NOT the reference's order -- entries are applied in ascending k with +1 and -1 interleaved -- so this kernel
// only ever runs when the caller opted in (tsg_set_fast_order / TSG_FAST_ORDER / order == TSG_ORDER_FAST).
//
// Why a second kernel: the gather kernel fetches one shared-memory operand word per add (25 % of the FP32-add peak at
// best, 22.6 % measured at 50-66 % sparsity).  When a quarter or more of W is non-zero it is cheaper to walk EVERY k and let
// the multiplier say what happens: acc = fma(x, w, acc) with w in {+1, 0, -1} -- exact for w = +-1 (x * 1 is exact, one
// rounding in the add), a no-op for w = 0 (for finite x).  One FFMA2 does two rows, costs the same two pipe cycles whether w is
// 0 or not (profiles/microbench F: predication does not make the zeros cheaper), and needs no index decoding.  The loop alone
// would run at density x ~0.9 x FP32 peak; with the shared-memory traffic of the real kernel (per k and warp 8 wavefronts of X,
// 4 broadcast wavefronts of W2, plus the TMA fill: 80 % of the crossbar) the FFMA2 pipe reaches 58 %: measured 3.16 ms at
// 4096^3 / 50 % sparsity (29 % of the FP32-add peak, gather kernel 4.08 ms = 22.6 %), break-even near 60 % sparsity.
//
// Mapping: a unit = 256 rows of X x 128 columns of W.  Lanes own rows (two of the 128-row K-major XT tiles the gather kernel
// uses: lane l holds rows l, l+32, l+64, l+96 of each half = 8 rows = four packed fp32x2), a warp owns 8 columns -> 64
// accumulators.  Per k a warp loads its 8 rows with two conflict-free LDS.128, the eight {w, w} pairs of its columns with four
// uniform-address LDS.128, and issues 32 FFMA2.  X chunks and the W2 stream (below) arrive through the same two-stage
// TMA/mbarrier ring as in gemm_tcsc.cu, one producer thread, setmaxnreg hand-off.
// W2 stream (private, built once per matrix from the TCSC arrays): W2[tile128][k][128] as float2 {w, w}.
#include "tsg_internal.h"
#include "tsg_f32x2.cuh"
#include "tsg_ptx.cuh"

namespace tsg {

constexpr int DF_TM = 256, DF_TN = 128, DF_NWARP = 16, DF_CW = 8;
constexpr int DF_THREADS = (DF_NWARP + 4) * 32;
constexpr int DF_REGS_COMPUTE = 112, DF_REGS_PRODUCER = 24;

struct DenseFastParams {
    const float *XT;    // K-major 128-row tiles (even number of tiles: the odd one out is zero)
    const float2 *W2;   // [ntile][K][128]
    const float *B;
    float *Y;
    long long ldy;
    int M, N, K, kc, nchunk, mtiles, ntiles;
    float a;
    int use_prelu;
};

// one warp per column: scatter +-1 into the zero-filled stream
__global__ void k_w2_fill(const int *__restrict__ csp, const int *__restrict__ csn, const int *__restrict__ rip, const int *__restrict__ rin,
                          int N, int K, float2 *__restrict__ W2) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    float2 *col = W2 + (size_t)(n / DF_TN) * K * DF_TN + (n % DF_TN);
    for (int t = csp[n] + lane; t < csp[n + 1]; t += 32) col[(size_t)rip[t] * DF_TN] = make_float2(1.f, 1.f);
    for (int t = csn[n] + lane; t < csn[n + 1]; t += 32) col[(size_t)rin[t] * DF_TN] = make_float2(-1.f, -1.f);
}

int build_w2(tsg_tcsc *W) {
    std::lock_guard<std::mutex> lk(W->mu);
    if (W->w2) return TSG_OK;
    const int ntiles = (W->cols + DF_TN - 1) / DF_TN;
    const size_t elems = (size_t)ntiles * (W->rows > 0 ? W->rows : 1) * DF_TN;
    float2 *w2 = nullptr;
    TSG_TRY(dev_alloc_t(&w2, elems));
    TSG_CUDA(cudaMemsetAsync(w2, 0, elems * sizeof(float2), stream()));
    if (W->cols > 0 && (W->n_pos + W->n_neg) > 0) {
        k_w2_fill<<<(unsigned)(((size_t)W->cols * 32 + 255) / 256), 256, 0, stream()>>>(W->csp, W->csn, W->rip, W->rin, W->cols, W->rows, w2);
        TSG_KERNEL_CHECK("k_w2_fill");
    }
    W->w2 = w2;
    return TSG_OK;
}

__global__ void __launch_bounds__(DF_THREADS, 1) k_tcsc_dense_fast(const DenseFastParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t xhalf_bytes = (uint32_t)p.kc * 512u;            // one 128-row half of the X chunk
    const uint32_t stage_bytes = 2u * xhalf_bytes + (uint32_t)p.kc * 1024u;  // + the W2 chunk (128 columns x 8 bytes per k)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 2 * (size_t)stage_bytes);
    uint64_t *empty = full + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], DF_NWARP);
        mbar_init(&empty[1], DF_NWARP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int units = p.mtiles * p.ntiles;

    if (warp >= DF_NWARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DF_REGS_PRODUCER));
        if (warp == DF_NWARP && lane == 0) {
            uint32_t it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int mt = u / p.ntiles, nt = u % p.ntiles;
                for (int c = 0; c < p.nchunk; ++c, ++it) {
                    const uint32_t s = it & 1u;
                    mbar_wait_relaxed(&empty[s], ((it >> 1) & 1u) ^ 1u);
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    const int rows = min(p.kc, p.K - c * p.kc);
                    const uint32_t xb = (uint32_t)rows * 512u, wb = (uint32_t)rows * 1024u;
                    mbar_arrive_expect_tx(&full[s], 2u * xb + wb);
                    bulk_g2s(st, p.XT + ((size_t)(2 * mt) * p.K + (size_t)c * p.kc) * 128, xb, &full[s]);
                    bulk_g2s(st + xhalf_bytes, p.XT + ((size_t)(2 * mt + 1) * p.K + (size_t)c * p.kc) * 128, xb, &full[s]);
                    bulk_g2s(st + 2 * xhalf_bytes, p.W2 + ((size_t)nt * p.K + (size_t)c * p.kc) * DF_TN, wb, &full[s]);
                }
            }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DF_REGS_COMPUTE));
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int mt = u / p.ntiles, nt = u % p.ntiles;
        float2 acc[DF_CW][4];
#pragma unroll
        for (int j = 0; j < DF_CW; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[j][v] = make_float2(0.f, 0.f);
        for (int c = 0; c < p.nchunk; ++c, ++it) {
            const uint32_t s = it & 1u;
            mbar_wait(&full[s], (it >> 1) & 1u);
            const uint32_t base = smem_addr(smem + (size_t)s * stage_bytes);
            const uint32_t xa = base + lane * 16, xb = base + xhalf_bytes + lane * 16;
            const uint32_t wa = base + 2 * xhalf_bytes + warp * (DF_CW * 8);
            const int rows = min(p.kc, p.K - c * p.kc);
            // (an explicitly double-buffered k loop was tried: 24 more live registers, spills, 3.85 instead of 3.16 ms at 4096^3 / 50 %)
#pragma unroll 2
            for (int k = 0; k < rows; ++k) {
                float4 x0, x1, w01, w23, w45, w67;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(xa + k * 512));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(xb + k * 512));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w01.x), "=f"(w01.y), "=f"(w01.z), "=f"(w01.w) : "r"(wa + k * 1024));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w23.x), "=f"(w23.y), "=f"(w23.z), "=f"(w23.w) : "r"(wa + k * 1024 + 16));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w45.x), "=f"(w45.y), "=f"(w45.z), "=f"(w45.w) : "r"(wa + k * 1024 + 32));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w67.x), "=f"(w67.y), "=f"(w67.z), "=f"(w67.w) : "r"(wa + k * 1024 + 48));
                const float2 xs[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y), make_float2(x1.z, x1.w)};
                const float2 ws[DF_CW] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w),
                                          make_float2(w45.x, w45.y), make_float2(w45.z, w45.w), make_float2(w67.x, w67.y), make_float2(w67.z, w67.w)};
#pragma unroll
                for (int j = 0; j < DF_CW; ++j)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[j][v] = ffma2(xs[v], ws[j], acc[j][v]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // epilogue: bias last, PReLU, store.  acc[j][v]: v = 0,1 -> rows l, l+32 / l+64, l+96 of the first half; v = 2,3 -> of the second
        const int nbase = nt * DF_TN + warp * DF_CW;
        const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) && (nbase + DF_CW <= p.N);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int m = mt * DF_TM + (r >> 2) * 128 + lane + 32 * (r & 3);
            if (m >= p.M) continue;
            float out[DF_CW];
#pragma unroll
            for (int j = 0; j < DF_CW; ++j) {
                const float2 pr = acc[j][r >> 1];
                float y = ((r & 1) ? pr.y : pr.x) + ((nbase + j < p.N) ? __ldg(p.B + nbase + j) : 0.f);
                if (p.use_prelu) y = (y < 0.0f) ? p.a * y : y;
                out[j] = y;
            }
            float *row = p.Y + (size_t)m * p.ldy + nbase;
            if (vec) {
                *reinterpret_cast<float4 *>(row) = make_float4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<float4 *>(row + 4) = make_float4(out[4], out[5], out[6], out[7]);
            } else {
#pragma unroll
                for (int j = 0; j < DF_CW; ++j)
                    if (nbase + j < p.N) row[j] = out[j];
            }
        }
    }
}

// *handled = 0: not the dense regime (or K == 0): the caller runs the gather kernel's fast order
int tcsc_gemm_dense_fast(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled) {
    *handled = 0;
    static const int env_dense = getenv("TSG_DENSE_FAST") ? atoi(getenv("TSG_DENSE_FAST")) : 1;
    const double density = (K > 0 && N > 0) ? ((double)W->n_pos + W->n_neg) / ((double)K * N) : 0.0;
    if (!env_dense || K <= 0 || density < 0.40) return TSG_OK;  // measured: 1.3x over the gather kernel at 50 % sparsity, break-even near 60 %
    TSG_TRY(build_w2(W));
    DenseFastParams p;
    p.mtiles = (M + DF_TM - 1) / DF_TM;
    p.ntiles = (N + DF_TN - 1) / DF_TN;
    const int tiles128 = 2 * p.mtiles;  // even: the transposer zero-fills rows >= M, so an odd last half is a tile of zeros
    float *XT = nullptr;
    WsHold ws(0);
    TSG_TRY(ws.acquire((size_t)tiles128 * K * 128 * sizeof(float), reinterpret_cast<void **>(&XT)));
    TSG_TRY(transpose_x_tiles(X, XT, M, K, tiles128));
    p.XT = XT; p.W2 = W->w2; p.B = B; p.Y = Y; p.ldy = ldy; p.M = M; p.N = N; p.K = K;
    p.a = a; p.use_prelu = use_prelu;
    p.kc = 52;  // 2 KB per k (two X halves + the W2 row): two stages of 104 KB
    if (p.kc > K) p.kc = K;
    p.nchunk = (K + p.kc - 1) / p.kc;
    const size_t smem = 2 * ((size_t)p.kc * 2048) + 64;
    static std::atomic<unsigned long long> attr_done{0};
    TSG_TRY(once_per_device(attr_done, [] {
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_dense_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        return (int)TSG_OK;
    }));
    const int units = p.mtiles * p.ntiles;
    const int grid = units < num_sms() ? units : num_sms();
    profile_mark(true);
    k_tcsc_dense_fast<<<grid, DF_THREADS, smem, stream()>>>(p);
    TSG_KERNEL_CHECK("k_tcsc_dense_fast");
    profile_mark(false);
    *handled = 1;
    return ws.release();
}

}  // namespace tsg
