// decode_bcsr.cu -- the decode shape (M < TSG_SKINNY_M rows of X) of the BCSR GEMM (sparse/bcsr.c:141-218): HBM-bound on the
// block values (4 bytes per stored element, zeros included -- that is what BCSR stores, bcsr.c:122-134).
//
// The reference's only BCSR GEMM test is this shape (test/test_bcsr.cpp:13-17: M=1, K=512, N=2048, 1x8 blocks).  The tiled
// kernels give every lane rows of X, which leaves them one useful lane in 32 here.  Instead:
//   * the rows of X live in shared memory ([K][MT], MT <= 8 rows interleaved per k);
//   * the block values are read from a private copy in block-COLUMN order (cval, built once per matrix), so the eight
//     warps of a CTA stream one block-column's values as contiguous 512-byte rows -- full HBM bursts instead of the
//     scattered 32-byte blocks of the row-major b_values;
//   * lanes split the block list, every lane keeps MT x 4 partial sums, a warp shuffle tree and one shared-memory pass
//     over the eight warps (fixed order) finish the column.
// Summation is therefore a tree: like the TCSC decode kernel this path meets the tolerance contract
// (max |y - y64| / max(|y64|,1) <= 1e-5), not bit-exactness; M >= TSG_SKINNY_M keeps the bit-exact kernels.
#include "tsg_internal.h"

namespace tsg {

// cval[t][r*c] = values[cblk[t]][r*c]: block values in column-major block order
__global__ void k_bcsr_permute(const float *__restrict__ values, const int *__restrict__ cblk, long long total, int rc, float *__restrict__ cval) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long t = i / rc;
    const int x = (int)(i - t * rc);
    cval[i] = __ldg(values + (size_t)__ldg(cblk + t) * rc + x);
}

int bcsr_build_cval(tsg_bcsr *W) {
    std::lock_guard<std::recursive_mutex> lk(W->mu);
    if (W->cval) return TSG_OK;
    TSG_TRY(bcsr_build_cols(W));
    const long long total = (long long)W->k * W->r * W->c;
    float *cv = nullptr;
    TSG_TRY(dev_alloc_t(&cv, (size_t)(total > 0 ? total : 1)));
    if (total > 0) {
        k_bcsr_permute<<<(unsigned)((total + 255) / 256), 256, 0, stream()>>>(W->values, W->cblk, total, W->r * W->c, cv);
        TSG_KERNEL_CHECK("k_bcsr_permute");
    }
    W->cval = cv;
    return TSG_OK;
}

constexpr int BD_WARPS = 8;

template <int C, int MT>
__global__ void __launch_bounds__(BD_WARPS * 32) k_bcsr_decode(const float *__restrict__ X, const int *__restrict__ cptr, const int *__restrict__ crow,
                                                               const float *__restrict__ cval, const float *__restrict__ B, float a, int use_prelu,
                                                               float *__restrict__ Y, long long ldy, int M, int K, int r, int bc) {
    constexpr int LPE = C / 4;       // lanes per entry (an entry = one block row: 1 x C values at one k)
    constexpr int EPW = 32 / LPE;    // entries per warp step
    extern __shared__ __align__(16) float smem_f[];
    float *xs = smem_f;                                  // [K][MT]
    float *part = smem_f + (size_t)K * MT;               // [BD_WARPS][MT][C] partial sums of the warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = blockIdx.y * MT;
    const int mrows = min(MT, M - m0);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float v[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) v[i] = (i < mrows) ? __ldg(X + (size_t)(m0 + i) * K + k) : 0.f;
        float *dst = xs + (size_t)k * MT;
        if (MT == 1) dst[0] = v[0];
        else if (MT == 2) *reinterpret_cast<float2 *>(dst) = make_float2(v[0], v[1 % MT]);
        else {
            *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1 % MT], v[2 % MT], v[3 % MT]);
            if (MT == 8) *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4 % MT], v[5 % MT], v[6 % MT], v[7 % MT]);
        }
    }
    __syncthreads();
    const int sub = lane % LPE;  // which float4 of the entry's C values
    for (int col = blockIdx.x; col < bc; col += gridDim.x) {
        const int e0 = __ldg(cptr + col), e1 = __ldg(cptr + col + 1);
        const long long E = (long long)(e1 - e0) * r;  // entries of this block-column
        float acc[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
        const float *cv = cval + (size_t)e0 * r * C;
        for (long long ent = warp * EPW + lane / LPE; ent < E; ent += 2 * BD_WARPS * EPW) {
            // two independent entries per trip: both index loads and both value loads are in flight together
            const long long ent2 = ent + BD_WARPS * EPW;
            const bool has2 = ent2 < E;
            const int blk1 = (r == 1) ? (int)ent : (int)(ent / r), blk2 = has2 ? ((r == 1) ? (int)ent2 : (int)(ent2 / r)) : blk1;
            const int br1 = __ldg(crow + e0 + blk1), br2 = __ldg(crow + e0 + blk2);
            const float4 w1 = __ldg(reinterpret_cast<const float4 *>(cv + ent * C) + sub);
            const float4 w2 = has2 ? __ldg(reinterpret_cast<const float4 *>(cv + ent2 * C) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int k1 = br1 * r + (int)(ent - (long long)blk1 * r), k2 = br2 * r + (int)((has2 ? ent2 : ent) - (long long)blk2 * r);
            const float *x1 = xs + (size_t)k1 * MT, *x2 = xs + (size_t)k2 * MT;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const float a1 = x1[i], a2 = x2[i];
                acc[i][0] = fmaf(a1, w1.x, acc[i][0]); acc[i][1] = fmaf(a1, w1.y, acc[i][1]);
                acc[i][2] = fmaf(a1, w1.z, acc[i][2]); acc[i][3] = fmaf(a1, w1.w, acc[i][3]);
                acc[i][0] = fmaf(a2, w2.x, acc[i][0]); acc[i][1] = fmaf(a2, w2.y, acc[i][1]);
                acc[i][2] = fmaf(a2, w2.z, acc[i][2]); acc[i][3] = fmaf(a2, w2.w, acc[i][3]);
            }
        }
        // lanes holding the same float4 slot of different entries: butterfly over the entry index
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int d = 16; d >= LPE; d >>= 1) acc[i][q] += __shfl_xor_sync(0xffffffffu, acc[i][q], d);
        if (lane < LPE) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
                *reinterpret_cast<float4 *>(part + ((size_t)warp * MT + i) * C + 4 * lane) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (threadIdx.x < MT * C) {
            const int i = threadIdx.x / C, j = threadIdx.x % C;
            if (i < mrows) {
                float y = 0.f;
#pragma unroll
                for (int w = 0; w < BD_WARPS; ++w) y += part[((size_t)w * MT + i) * C + j];
                y += __ldg(B + col * C + j);
                if (use_prelu) y = (y < 0.0f) ? a * y : y;
                Y[(size_t)(m0 + i) * ldy + col * C + j] = y;
            }
        }
        __syncthreads();
    }
}

template <int C, int MT>
static int launch_bcsr_decode(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int K, long long ldy, int groups) {
    const size_t smem = ((size_t)K * MT + (size_t)BD_WARPS * MT * C) * 4;
    static std::atomic<unsigned long long> attr_done{0};
    TSG_TRY(once_per_device(attr_done, [] {
        TSG_CUDA(cudaFuncSetAttribute(k_bcsr_decode<C, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        return (int)TSG_OK;
    }));
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int gx = W->bc;
    const int cap = (num_sms() * per_sm + groups - 1) / groups;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    k_bcsr_decode<C, MT><<<dim3(gx, groups), BD_WARPS * 32, smem, stream()>>>(X, W->cptr, W->crow, W->cval, B, a, use_prelu, Y, ldy, M, K, W->r, W->bc);
    TSG_KERNEL_CHECK("k_bcsr_decode");
    return TSG_OK;
}

// *handled = 0 when the shape is outside this kernel (block width not 4/8/16, or a row of X too long for shared memory)
int bcsr_decode(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled) {
    *handled = 0;
    (void)N;
    const int c = W->c;
    if (!(c == 4 || c == 8 || c == 16) || W->bc <= 0 || K <= 0) return TSG_OK;
    constexpr size_t kSmemMax = 216 * 1024;
    if ((size_t)K * 4 > kSmemMax) return TSG_OK;
    int mt = 8;
    while (mt > 1 && (mt / 2 >= M || ((size_t)K * mt + (size_t)BD_WARPS * mt * c) * 4 > kSmemMax)) mt /= 2;
    const int groups = (M + mt - 1) / mt;
    TSG_TRY(bcsr_build_cval(W));
    int rc = TSG_OK;
#define TSG_BD(CC, MM) rc = launch_bcsr_decode<CC, MM>(W, X, B, a, use_prelu, Y, M, K, ldy, groups)
#define TSG_BD_MT(CC)            \
    switch (mt) {                \
        case 1: TSG_BD(CC, 1); break; \
        case 2: TSG_BD(CC, 2); break; \
        case 4: TSG_BD(CC, 4); break; \
        default: TSG_BD(CC, 8); break; \
    }
    if (c == 4) { TSG_BD_MT(4) } else if (c == 8) { TSG_BD_MT(8) } else { TSG_BD_MT(16) }
#undef TSG_BD_MT
#undef TSG_BD
    if (rc == TSG_OK) *handled = 1;
    return rc;
}

}  // namespace tsg
