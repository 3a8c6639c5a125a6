// gemm_bcsr.cu -- BCSR variant:  Y = [PReLU](X*W + B), W stored as dense r x c float blocks.
//
// Replaces bcsr_sgemm_basic / _avx / _avx2 (sparse/bcsr.c:141-175,222-261,316-385) and computes the north-star
// math for bcsr_sgemm_prelu_basic / _prelu_avx (bcsr.c:177-218,264-312) by default.  use_prelu == 2 selects the
// reference's LITERAL loop instead: the activation `res > 0 ? res : a*res` runs after every single partial update
// (bcsr.c:208-212, :296-302) and outputs no block touches keep the raw bias -- per output element still a sequential walk
// in ascending k, so the plain kernel below reproduces it bit for bit (see include/sparse/bcsr.h).
//
// The reference walks block-rows and scatters into Y (read-modify-write of Y per block, bcsr.c:168).  On a GPU the
// output tile lives in registers instead, so the kernel needs the blocks of one block-COLUMN in ascending block-row
// order: a private column-major index (tsg_bcsr::cptr/crow/cblk, handles.cu) is derived once per matrix.  Per output
// element the partial products are then accumulated starting from the bias in ascending k -- the same order as
// bcsr.c:141-175 -- with FFMA (x*val+y, one rounding, as the reference's AVX2 path and g++ -O3 -march=native do;
// for ternary block values the product is exact either way).
//
// Mapping: a CTA owns 128 rows of X (the K-major tile XT[mtile][k][128] shared with the TCSC kernel); every warp
// owns one block-column at a time (C output columns), lane l holds rows l, l+32, l+64, l+96 (one float4 of XT per k) -> C x 4 accumulators per thread.
// X rows are read straight from the L2-resident XT tile with coalesced 512-byte warp loads.
#include "tsg_internal.h"

namespace tsg {

// one partial update of an output element; LIT: the reference's literal prelu loop (activation after every update)
template <bool LIT>
__device__ __forceinline__ float bcsr_upd(float x, float w, float y, float a) {
    const float res = fmaf(x, w, y);
    if (LIT) return (res > 0.0f) ? res : a * res;
    return res;
}

template <int C, bool LIT>
__global__ void __launch_bounds__(256) k_bcsr_gemm(const float *__restrict__ XT, const int *__restrict__ cptr, const int *__restrict__ crow,
                                                   const int *__restrict__ cblk, const float *__restrict__ values, const float *__restrict__ B,
                                                   float a, int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int K, int r,
                                                   int bc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mt = blockIdx.y;
    const float *xt = XT + (size_t)mt * K * 128 + lane * 4;
    const int mbase = mt * 128 + lane;  // XT tile layout: lane l's float4 holds rows l, l+32, l+64, l+96 (gemm_tcsc.cu)
    for (int col = blockIdx.x * 8 + warp; col < bc; col += gridDim.x * 8) {
        float acc[C][4];
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const float b = __ldg(B + col * C + j);  // bcsr.c:146-150: Y starts as the bias
            acc[j][0] = b; acc[j][1] = b; acc[j][2] = b; acc[j][3] = b;
        }
        const int e0 = __ldg(cptr + col), e1 = __ldg(cptr + col + 1);
        if (r == 1) {
            // one-row blocks (the blocking the reference tests, test_bcsr.cpp:16-17): batches of four blocks with all
            // loads issued before the first FMA, so four X rows and four value rows are in flight per warp
            int e = e0;
            for (; e + 4 <= e1; e += 4) {
                int br4[4], bk4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { br4[u] = __ldg(crow + e + u); bk4[u] = __ldg(cblk + e + u); }
                float4 x4[4];
                float w4[4][C];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    x4[u] = __ldg(reinterpret_cast<const float4 *>(xt + (size_t)br4[u] * 128));
                    const float *blk = values + (size_t)bk4[u] * C;
#pragma unroll
                    for (int j = 0; j < C; ++j) w4[u][j] = __ldg(blk + j);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // ascending block-row = ascending k, as bcsr.c:152-170 accumulates
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        acc[j][0] = bcsr_upd<LIT>(x4[u].x, w4[u][j], acc[j][0], a);
                        acc[j][1] = bcsr_upd<LIT>(x4[u].y, w4[u][j], acc[j][1], a);
                        acc[j][2] = bcsr_upd<LIT>(x4[u].z, w4[u][j], acc[j][2], a);
                        acc[j][3] = bcsr_upd<LIT>(x4[u].w, w4[u][j], acc[j][3], a);
                    }
                }
            }
            for (; e < e1; ++e) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(xt + (size_t)__ldg(crow + e) * 128));
                const float *blk = values + (size_t)__ldg(cblk + e) * C;
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const float w = __ldg(blk + j);
                    acc[j][0] = bcsr_upd<LIT>(x.x, w, acc[j][0], a);
                    acc[j][1] = bcsr_upd<LIT>(x.y, w, acc[j][1], a);
                    acc[j][2] = bcsr_upd<LIT>(x.z, w, acc[j][2], a);
                    acc[j][3] = bcsr_upd<LIT>(x.w, w, acc[j][3], a);
                }
            }
        } else {
            for (int e = e0; e < e1; ++e) {
                const int brow = __ldg(crow + e);
                const float *blk = values + (size_t)__ldg(cblk + e) * r * C;
                for (int i = 0; i < r; ++i) {
                    const float4 x = __ldg(reinterpret_cast<const float4 *>(xt + (size_t)(brow * r + i) * 128));
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        const float w = __ldg(blk + i * C + j);
                        acc[j][0] = bcsr_upd<LIT>(x.x, w, acc[j][0], a);
                        acc[j][1] = bcsr_upd<LIT>(x.y, w, acc[j][1], a);
                        acc[j][2] = bcsr_upd<LIT>(x.z, w, acc[j][2], a);
                        acc[j][3] = bcsr_upd<LIT>(x.w, w, acc[j][3], a);
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int m = mbase + 32 * v;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < C; ++j) {
                float y = acc[j][v];
                if (!LIT && use_prelu) y = (y < 0.0f) ? a * y : y;
                Y[(size_t)m * ldy + col * C + j] = y;
            }
        }
    }
}

// any block width: one thread per output element (slow path, correctness only)
__global__ void k_bcsr_gemm_generic(const float *__restrict__ X, const int *__restrict__ cptr, const int *__restrict__ crow,
                                    const int *__restrict__ cblk, const float *__restrict__ values, const float *__restrict__ B, float a,
                                    int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int K, int r, int c, int bc) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)M * N) return;
    const int m = (int)(e / N), n = (int)(e % N);
    float y = B[n];
    const int col = n / c, j = n % c;
    if (col < bc) {
        for (int t = cptr[col]; t < cptr[col + 1]; ++t) {
            const int brow = crow[t];
            const float *blk = values + (size_t)cblk[t] * r * c;
            for (int i = 0; i < r; ++i) {
                y = fmaf(X[(size_t)m * K + brow * r + i], blk[i * c + j], y);
                if (use_prelu == 2) y = (y > 0.0f) ? y : a * y;  // the reference's literal loop
            }
        }
    }
    if (use_prelu == 1) y = (y < 0.0f) ? a * y : y;
    Y[(size_t)m * ldy + n] = y;
}

// columns >= bc*c (cols % c remainder, bcsr.c:25) only ever receive the bias
__global__ void k_bcsr_tail_bias(const float *__restrict__ B, float a, int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int n0) {
    const int w = N - n0;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)M * w) return;
    const int m = (int)(e / w), n = n0 + (int)(e % w);
    float y = B[n];
    if (use_prelu) y = (y < 0.0f) ? a * y : y;
    Y[(size_t)m * ldy + n] = y;
}

}  // namespace tsg

using namespace tsg;

static thread_local int g_bcsr_kernel = 0;  // 0 = default (ring kernel unless TSG_BCSR_RING=0), 1 = plain, 2 = ring

extern "C" int tsg_bcsr_set_kernel(int which) {
    if (which < 0 || which > 2) return set_error(TSG_EINVAL, "tsg_bcsr_set_kernel: 0 (default), 1 (plain) or 2 (ring)");
    g_bcsr_kernel = which;
    return TSG_OK;
}

// what the reference-named bcsr_sgemm_prelu_* entry points compute: 0 = PReLU(X*W + B) (default), 1 = the reference's literal loop
static thread_local int g_prelu_literal = -1;  // -1: not decided yet (environment TSG_BCSR_PRELU_LITERAL)
extern "C" int tsg_bcsr_set_prelu_literal(int on) {
    g_prelu_literal = on ? 1 : 0;
    return TSG_OK;
}
extern "C" int tsg_bcsr_get_prelu_literal(void) {
    if (g_prelu_literal < 0) {
        const char *e = getenv("TSG_BCSR_PRELU_LITERAL");
        g_prelu_literal = (e && atoi(e) != 0) ? 1 : 0;
    }
    return g_prelu_literal;
}

extern "C" int tsg_bcsr_gemm(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy) {
    TSG_TRY(ensure_device());
    if (!W || !X || !B || !Y) return set_error(TSG_EINVAL, "tsg_bcsr_gemm: null argument");
    if (W->bc * W->c > N || W->br * W->r > K) return set_error(TSG_EINVAL, "tsg_bcsr_gemm: W covers %d x %d but K=%d, N=%d", W->br * W->r, W->bc * W->c, K, N);
    if (ldy < N) return set_error(TSG_EINVAL, "tsg_bcsr_gemm: ldy < N");
    if (M <= 0 || N <= 0) return TSG_OK;
    if (use_prelu < 0 || use_prelu > 2) return set_error(TSG_EINVAL, "tsg_bcsr_gemm: use_prelu must be 0, 1 or 2 (the reference's literal loop)");
    const bool literal = use_prelu == 2;  // sequential by definition: the plain kernels only
    TSG_TRY(bcsr_build_cols(W));
    cudaStream_t st = stream();
    const int c = W->c, r = W->r, bc = W->bc;
    const int ncov = bc * c;
    if (ncov < N) {
        const long long total = (long long)M * (N - ncov);
        k_bcsr_tail_bias<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, a, literal ? 0 : use_prelu, Y, ldy, M, N, ncov);  // literal: untouched outputs keep the raw bias
        TSG_KERNEL_CHECK("k_bcsr_tail_bias");
    }
    if (bc == 0) return TSG_OK;
    static const int env_decode = getenv("TSG_DECODE") ? atoi(getenv("TSG_DECODE")) : 1;
    if (env_decode && M < TSG_SKINNY_M && g_bcsr_kernel == 0 && !literal) {  // decode shape: lanes over the block list (decode_bcsr.cu), tolerance contract
        int handled = 0;
        TSG_TRY(bcsr_decode(W, X, B, a, use_prelu, Y, M, N, K, ldy, &handled));
        if (handled) return TSG_OK;
    }
    if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) {
        const int mtiles = (M + 127) / 128;
        float *XT = nullptr;
        WsHold ws(0);
        TSG_TRY(ws.acquire((size_t)mtiles * K * 128 * sizeof(float), reinterpret_cast<void **>(&XT)));
        TSG_TRY(transpose_x_tiles(X, XT, M, K));
        static const int env_ring = getenv("TSG_BCSR_RING") ? atoi(getenv("TSG_BCSR_RING")) : 1;
        if (!literal && (g_bcsr_kernel == 2 || (g_bcsr_kernel == 0 && env_ring))) {
            int handled = 0;
            const int rc = bcsr_gemm_ring(W, XT, B, a, use_prelu, Y, M, N, K, ldy, &handled);
            if (rc != TSG_OK || handled) {
                const int rc2 = ws.release();
                return rc ? rc : rc2;
            }
        }
        int gx = (bc + 7) / 8;
        const int cap = (num_sms() * 8 + mtiles - 1) / mtiles;
        if (gx > cap) gx = cap;
        dim3 grid(gx < 1 ? 1 : gx, mtiles);
#define TSG_BCSR_LAUNCH(CC)                                                                                                              \
    do {                                                                                                                                 \
        if (literal) k_bcsr_gemm<CC, true><<<grid, 256, 0, st>>>(XT, W->cptr, W->crow, W->cblk, W->values, B, a, use_prelu, Y, ldy, M, N, K, r, bc); \
        else k_bcsr_gemm<CC, false><<<grid, 256, 0, st>>>(XT, W->cptr, W->crow, W->cblk, W->values, B, a, use_prelu, Y, ldy, M, N, K, r, bc);        \
    } while (0)
        switch (c) {
            case 1: TSG_BCSR_LAUNCH(1); break;
            case 2: TSG_BCSR_LAUNCH(2); break;
            case 4: TSG_BCSR_LAUNCH(4); break;
            case 8: TSG_BCSR_LAUNCH(8); break;
            default: TSG_BCSR_LAUNCH(16); break;
        }
#undef TSG_BCSR_LAUNCH
        TSG_KERNEL_CHECK("k_bcsr_gemm");
        return ws.release();
    }
    const long long total = (long long)M * ncov;
    // generic path writes columns [0, ncov): reuse the element kernel with N restricted via ldy addressing
    k_bcsr_gemm_generic<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, W->cptr, W->crow, W->cblk, W->values, B, a, use_prelu, Y, ldy, M,
                                                                       ncov, K, r, c, bc);
    TSG_KERNEL_CHECK("k_bcsr_gemm_generic");
    return TSG_OK;
}
