// decode_tcsc.cu -- the decode shape of the sparse ternary GEMM (M < TSG_SKINNY_M rows of X): HBM/L2-bound on the index stream.
//
// Same math as tcsc_sgemm_prelu_basic (sparse/tcsc.c:143-165) and the other TCSC entry points; summation is a tree (lanes
// stride over a column's non-zeros, then a warp reduction), so the contract is the tolerance one of DESIGN.md section 2
// (max |y - y64| / max(|y64|, 1) <= 1e-5), exactly as for the previous skinny kernel.
//
// What changed against the first skinny kernel (gemm_tcsc.cu, kept as the fall-back for very large K): the rows of X live in
// SHARED memory for the whole CTA ([K][MT] floats, MT <= 8 rows interleaved per k so one 4*MT-byte load fetches a k for all
// rows), so the dependent half of every gather (index -> X) is a 29-cycle LDS instead of an L2 round trip, the public int32
// index arrays are streamed with four independent coalesced loads in flight per lane, and CTAs are sized to fill the SMs'
// thread slots (one column per warp at 4096 columns).  Algorithmic bytes: 4 B per non-zero + 8 B per column + X + Y.
#include "tsg_internal.h"

namespace tsg {

template <int MT>
__device__ __forceinline__ void dec_add(float (&acc)[MT], const float *xs, int k, float sign) {
    const float *p = xs + (size_t)k * MT;
    if (MT == 1) {
        acc[0] += sign * p[0];
    } else if (MT == 2) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        acc[0] += sign * v.x; acc[1 % MT] += sign * v.y;
    } else {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        acc[0] += sign * v.x; acc[1 % MT] += sign * v.y; acc[2 % MT] += sign * v.z; acc[3 % MT] += sign * v.w;
        if (MT == 8) {
            const float4 w = *reinterpret_cast<const float4 *>(p + 4);
            acc[4 % MT] += sign * w.x; acc[5 % MT] += sign * w.y; acc[6 % MT] += sign * w.z; acc[7 % MT] += sign * w.w;
        }
    }
}

// lanes stride over [lo, hi): up to four independent index loads, then their (shared-memory) gathers
template <int MT>
__device__ __forceinline__ void dec_accumulate(float (&acc)[MT], const float *xs, const int *__restrict__ idx, int lo, int hi, int lane, float sign) {
    for (int t = lo + lane; t < hi; t += 128) {
        int k[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = (t + 32 * j < hi) ? __ldg(idx + t + 32 * j) : -1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k[j] >= 0) dec_add<MT>(acc, xs, k[j], sign);
    }
}

template <int MT>
__global__ void __launch_bounds__(1024, 1) k_tcsc_decode(const float *__restrict__ X, const int *__restrict__ csp, const int *__restrict__ csn,
                                                         const int *__restrict__ rip, const int *__restrict__ rin, const float *__restrict__ B, float a,
                                                         int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int K) {
    extern __shared__ __align__(16) float xs[];  // [K][MT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int m0 = blockIdx.y * MT;
    const int mrows = min(MT, M - m0);
    // column pointers first: their latency overlaps the staging of X
    int n = blockIdx.x * nw + warp;
    int p0 = 0, p1 = 0, q0 = 0, q1 = 0;
    if (n < N) { p0 = __ldg(csp + n); p1 = __ldg(csp + n + 1); q0 = __ldg(csn + n); q1 = __ldg(csn + n + 1); }
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
    const int stride = gridDim.x * nw;
    while (n < N) {
        float acc[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) acc[i] = 0.f;
        dec_accumulate<MT>(acc, xs, rip, p0, p1, lane, 1.0f);
        dec_accumulate<MT>(acc, xs, rin, q0, q1, lane, -1.0f);
        const int nn = n + stride;  // next column's pointers: in flight during the reduction
        if (nn < N) { p0 = __ldg(csp + nn); p1 = __ldg(csp + nn + 1); q0 = __ldg(csn + nn); q1 = __ldg(csn + nn + 1); }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
        }
        if (lane == 0) {
            const float b = __ldg(B + n);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                if (i < mrows) {
                    float y = acc[i] + b;
                    if (use_prelu) y = (y < 0.0f) ? a * y : y;
                    Y[(size_t)(m0 + i) * ldy + n] = y;
                }
            }
        }
        n = nn;
    }
}

// *handled = 0 when the rows of X do not fit shared memory even one at a time (K > ~56 K): the caller falls back
int tcsc_decode(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy, int *handled) {
    *handled = 0;
    constexpr size_t kSmemMax = 224 * 1024;
    if (K <= 0 || (size_t)K * 4 > kSmemMax) return TSG_OK;
    int mt = 8;
    while (mt > 1 && (mt / 2 >= M || (size_t)K * mt * 4 > kSmemMax)) mt /= 2;  // smallest power of two >= M that fits (<= 8)
    const int groups = (M + mt - 1) / mt;
    const size_t smem = (size_t)K * mt * 4;
    // warps: one column per warp where that still fills the machine; 1024-thread CTAs when the X copy allows one CTA per SM only
    const int per_sm = (smem <= 100 * 1024) ? 2 : 1;
    int threads = (per_sm == 2) ? 512 : 1024;
    const int nw = threads / 32;
    int gx = (N + nw - 1) / nw;
    const int cap = (num_sms() * per_sm + groups - 1) / groups;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    static std::atomic<unsigned long long> attr_done{0};
    TSG_TRY(once_per_device(attr_done, [] {
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_decode<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_decode<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_decode<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_decode<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        return (int)TSG_OK;
    }));
    dim3 grid(gx, groups);
    cudaStream_t st = stream();
    switch (mt) {
        case 1: k_tcsc_decode<1><<<grid, threads, smem, st>>>(X, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        case 2: k_tcsc_decode<2><<<grid, threads, smem, st>>>(X, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        case 4: k_tcsc_decode<4><<<grid, threads, smem, st>>>(X, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        default: k_tcsc_decode<8><<<grid, threads, smem, st>>>(X, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
    }
    TSG_KERNEL_CHECK("k_tcsc_decode");
    *handled = 1;
    return TSG_OK;
}

}  // namespace tsg
