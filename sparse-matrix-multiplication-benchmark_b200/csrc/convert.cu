// convert.cu -- dense -> TCSC and dense -> BCSR conversion on the device.
//
// Replaces the reference builders tcsc_from_dense (sparse/tcsc.c:6-66), SparseFormat::SparseFormat
// (SparseGEMM.h:20-39) and bcsr_from_dense (sparse/bcsr.c:19-139).  The outputs are bit-identical to the
// reference's arrays: entries inside a column are in ascending row order, columns in ascending order.
//
// TCSC pipeline (HBM-bound: the dense matrix is read exactly once, 4*K*N bytes):
//   1. k_tcsc_masks   one thread per (32-row strip, column): 32 coalesced row reads -> a 32-bit "+1" ballot mask and
//                     a "-1" ballot mask per strip and column (1/16 of the dense bytes), popcounts accumulated per
//                     column with one packed 64-bit atomic (pos in the low word, neg in the high word).
//   2. k_scan_columns exclusive prefix scan over the N packed column counts -> col_start_pos / col_start_neg (+ totals)
//   3. k_tcsc_fill    per 32-column slab: prefix over the strips of each column, then every (strip, column) expands
//                     its two masks into ascending row indices at its offset.
#include "tsg_internal.h"

namespace tsg {

// ---- predicates ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_pos(float v) { return v == 1.0f; }   // tcsc.c:14
__device__ __forceinline__ bool is_neg(float v) { return v == -1.0f; }  // tcsc.c:16
__device__ __forceinline__ bool is_pos(int v) { return v >= 1; }        // SparseGEMM.h:26
__device__ __forceinline__ bool is_neg(int v) { return v <= -1; }       // SparseGEMM.h:30

// ---- 1. masks + column counts -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_tcsc_masks(const T *__restrict__ dense, int K, int N, uint32_t *__restrict__ posmask,
                                                    uint32_t *__restrict__ negmask, unsigned long long *__restrict__ colcnt) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    const int s = blockIdx.y;
    if (n >= N) return;
    const int row0 = s * 32;
    const T *p = dense + (size_t)row0 * N + n;
    uint32_t pm = 0, nm = 0;
    const int nrows = min(32, K - row0);
    if (nrows == 32) {
        T v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = __ldg(p + (size_t)r * N);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            pm |= (uint32_t)is_pos(v[r]) << r;
            nm |= (uint32_t)is_neg(v[r]) << r;
        }
    } else {
        for (int r = 0; r < nrows; ++r) {
            T v = __ldg(p + (size_t)r * N);
            pm |= (uint32_t)is_pos(v) << r;
            nm |= (uint32_t)is_neg(v) << r;
        }
    }
    posmask[(size_t)s * N + n] = pm;
    negmask[(size_t)s * N + n] = nm;
    unsigned long long packed = (unsigned long long)__popc(pm) | ((unsigned long long)__popc(nm) << 32);
    if (packed) atomicAdd(colcnt + n, packed);
}

// ---- 2. exclusive scan over columns (single CTA, packed pos|neg<<32) ----------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_columns(const unsigned long long *__restrict__ colcnt, int N, int *__restrict__ csp,
                                                       int *__restrict__ csn, int *__restrict__ totals) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int n = base + tid;
        unsigned long long v = (n < N) ? colcnt[n] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            warp_sums[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long excl = carry + warp_sums[warp] + (incl - v);
        if (n < N) {
            csp[n] = (int)(uint32_t)(excl & 0xffffffffull);
            csn[n] = (int)(uint32_t)(excl >> 32);
        }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long tot = carry_s;
        csp[N] = (int)(uint32_t)(tot & 0xffffffffull);
        csn[N] = (int)(uint32_t)(tot >> 32);
        totals[0] = csp[N];
        totals[1] = csn[N];
    }
}

// ---- 3. fill: 32 columns x 32 strip-groups per CTA ------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tcsc_fill(const uint32_t *__restrict__ posmask, const uint32_t *__restrict__ negmask, int K,
                                                    int N, int S, const int *__restrict__ csp, const int *__restrict__ csn,
                                                    int *__restrict__ rip, int *__restrict__ rin) {
    __shared__ int pos_part[32][33];
    __shared__ int neg_part[32][33];
    const int x = threadIdx.x;  // column inside the slab
    const int y = threadIdx.y;  // strip group
    const int n = blockIdx.x * 32 + x;
    const int per = (S + 31) / 32;
    const int s0 = y * per, s1 = min(S, s0 + per);
    int pc = 0, nc = 0;
    if (n < N)
        for (int s = s0; s < s1; ++s) {
            pc += __popc(posmask[(size_t)s * N + n]);
            nc += __popc(negmask[(size_t)s * N + n]);
        }
    pos_part[y][x] = pc;
    neg_part[y][x] = nc;
    __syncthreads();
    if (y == 0) {  // exclusive prefix down the 32 strip groups of column x
        int ap = 0, an = 0;
        for (int g = 0; g < 32; ++g) {
            int tp = pos_part[g][x], tn = neg_part[g][x];
            pos_part[g][x] = ap;
            neg_part[g][x] = an;
            ap += tp;
            an += tn;
        }
    }
    __syncthreads();
    if (n >= N) return;
    int op = csp[n] + pos_part[y][x];
    int on = csn[n] + neg_part[y][x];
    for (int s = s0; s < s1; ++s) {
        uint32_t pm = posmask[(size_t)s * N + n], nm = negmask[(size_t)s * N + n];
        const int row0 = s * 32;
        while (pm) {
            rip[op++] = row0 + __ffs(pm) - 1;
            pm &= pm - 1;
        }
        while (nm) {
            rin[on++] = row0 + __ffs(nm) - 1;
            nm &= nm - 1;
        }
    }
}

template <typename T>
static int tcsc_from_dense_impl(const T *dense, int rows, int cols, tsg_tcsc **out) {
    *out = nullptr;
    TSG_TRY(ensure_device());
    if (rows < 0 || cols < 0 || (!dense && rows > 0 && cols > 0)) return set_error(TSG_EINVAL, "tcsc_from_dense: bad arguments");
    if ((long long)rows * cols > 0x7fffffffLL) return set_error(TSG_EINVAL, "tcsc_from_dense: rows*cols exceeds the int32 range of the TCSC format");
    cudaStream_t st = stream();
    tsg_tcsc *W = new (std::nothrow) tsg_tcsc();
    if (!W) return set_error(TSG_ENOMEM, "out of host memory");
    W->rows = rows;
    W->cols = cols;
    const int K = rows, N = cols, S = (K + 31) / 32;
    uint32_t *pm = nullptr, *nm = nullptr;
    unsigned long long *cc = nullptr;
    int *totals = nullptr;
    int rc = TSG_OK;
    bool ws_held = false;
    auto fail = [&](int code) {
        if (ws_held) ws_release(2);
        tsg_tcsc_destroy(W);
        return code;
    };
    if ((rc = dev_alloc_t(&W->csp, (size_t)N + 1))) return fail(rc);
    if ((rc = dev_alloc_t(&W->csn, (size_t)N + 1))) return fail(rc);
    {   // scratch (ballot masks, packed column counts, totals) lives in a persistent per-thread workspace: converting
        // matrix after matrix must not churn the memory pool (a pool growth costs tens of milliseconds)
        const size_t mask_bytes = (((size_t)S * N * 4) + 255) & ~(size_t)255;
        const size_t cc_bytes = ((((size_t)N + 1) * 8) + 255) & ~(size_t)255;
        void *base = nullptr;
        if ((rc = ws_acquire(2, 2 * mask_bytes + cc_bytes + 256, &base))) return fail(rc);
        ws_held = true;
        char *b = static_cast<char *>(base);
        pm = reinterpret_cast<uint32_t *>(b);
        nm = reinterpret_cast<uint32_t *>(b + mask_bytes);
        cc = reinterpret_cast<unsigned long long *>(b + 2 * mask_bytes);
        totals = reinterpret_cast<int *>(b + 2 * mask_bytes + cc_bytes);
    }
    if (cudaMemsetAsync(cc, 0, ((size_t)N + 1) * sizeof(unsigned long long), st) != cudaSuccess) return fail(set_error(TSG_ECUDA, "memset failed"));
    if (N > 0 && S > 0) {
        dim3 grid((N + 127) / 128, S);
        k_tcsc_masks<T><<<grid, 128, 0, st>>>(dense, K, N, pm, nm, cc);
        if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_tcsc_masks failed"));
        count_launch();
    }
    k_scan_columns<<<1, 1024, 0, st>>>(cc, N, W->csp, W->csn, totals);
    if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_scan_columns failed"));
    count_launch();
    int h_tot[2] = {0, 0};
    if (cudaMemcpyAsync(h_tot, totals, sizeof h_tot, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return fail(set_error(TSG_ECUDA, "tcsc_from_dense: device failure: %s", cudaGetErrorString(cudaGetLastError())));
    W->n_pos = h_tot[0];
    W->n_neg = h_tot[1];
    if ((rc = dev_alloc_t(&W->rip, (size_t)W->n_pos))) return fail(rc);
    if ((rc = dev_alloc_t(&W->rin, (size_t)W->n_neg))) return fail(rc);
    if (N > 0 && S > 0 && (W->n_pos + W->n_neg) > 0) {
        k_tcsc_fill<<<(N + 31) / 32, dim3(32, 32), 0, st>>>(pm, nm, K, N, S, W->csp, W->csn, W->rip, W->rin);
        if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_tcsc_fill failed"));
        count_launch();
    }
    ws_release(2);
    *out = W;
    return TSG_OK;
}

// =====================================================================================================================
// generic exclusive scan over uint32 (used by the BCSR builder): 4096 items per CTA, two levels
// =====================================================================================================================
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *total, uint32_t *smem33) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) smem33[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = smem33[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        smem33[lane] = wi - w;
        if (lane == 31) smem33[32] = wi;
    }
    __syncthreads();
    uint32_t r = smem33[warp] + incl - v;
    *total = smem33[32];
    __syncthreads();
    return r;
}

// level 1: per-CTA (4096 items) local exclusive scan, CTA total to sums[blockIdx]
__global__ void __launch_bounds__(1024) k_scan_l1(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long n,
                                                  uint32_t *__restrict__ sums) {
    __shared__ uint32_t sm[33];
    const long long base = (long long)blockIdx.x * 4096 + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
    uint32_t tsum = v[0] + v[1] + v[2] + v[3], total;
    uint32_t ex = block_exclusive_scan_1024(tsum, &total, sm);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// level 2: single CTA scans the CTA totals in place (exclusive), grand total to *total_out
__global__ void __launch_bounds__(1024) k_scan_l2(uint32_t *__restrict__ sums, int nb, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = (i < nb) ? sums[i] : 0u, total;
        uint32_t ex = block_exclusive_scan_1024(v, &total, sm);
        const uint32_t carry = carry_s;
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}
__global__ void __launch_bounds__(1024) k_scan_l3(uint32_t *__restrict__ out, long long n, const uint32_t *__restrict__ sums) {
    const long long base = (long long)blockIdx.x * 4096 + threadIdx.x * 4;
    const uint32_t add = sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) out[base + i] += add;
}

int scan_exclusive_u32(const uint32_t *in, uint32_t *out, long long n, uint32_t *total_dev) {
    cudaStream_t st = stream();
    const int nb = (int)((n + 4095) / 4096);
    uint32_t *sums = nullptr;
    TSG_TRY(ws_acquire(4, ((size_t)nb + 1) * sizeof(uint32_t), reinterpret_cast<void **>(&sums)));
    if (nb > 0) {
        k_scan_l1<<<nb, 1024, 0, st>>>(in, out, n, sums);
        TSG_KERNEL_CHECK("k_scan_l1");
    }
    k_scan_l2<<<1, 1024, 0, st>>>(sums, nb, total_dev);
    TSG_KERNEL_CHECK("k_scan_l2");
    if (nb > 1) {
        k_scan_l3<<<nb, 1024, 0, st>>>(out, n, sums);
        TSG_KERNEL_CHECK("k_scan_l3");
    }
    return ws_release(4);
}

// =====================================================================================================================
// BCSR builder (sparse/bcsr.c:19-139)
// =====================================================================================================================
// a block is kept iff it holds at least one +-1 (bcsr.c:62); one thread per block, row-major block order
__global__ void k_bcsr_flags(const float *__restrict__ dense, int cols, int r, int c, int br, int bc, uint32_t *__restrict__ flags) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (long long)br * bc) return;
    const int brow = (int)(b / bc), bcol = (int)(b % bc);
    const float *p = dense + (size_t)brow * r * cols + (size_t)bcol * c;
    bool keep = false;
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            float v = __ldg(p + (size_t)i * cols + j);
            keep |= (v == 1.0f) | (v == -1.0f);
        }
    flags[b] = keep ? 1u : 0u;
}
// standard CSR row pointers from the scanned flags (see include/sparse/bcsr.h for the deviation from bcsr.c:114-117)
__global__ void k_bcsr_rowptr(const uint32_t *__restrict__ scanned, const uint32_t *__restrict__ total, int br, int bc,
                              int *__restrict__ row_start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < br) row_start[i] = (int)scanned[(size_t)i * bc];
    if (i == br) row_start[br] = (int)*total;
}
// one warp-slice of threads per kept block copies the whole r x c block, zeros and non-ternary values included
// (bcsr.c:122-134), and records its block column (bcsr.c:119)
__global__ void k_bcsr_fill(const float *__restrict__ dense, int cols, int r, int c, int br, int bc, const uint32_t *__restrict__ flags,
                            const uint32_t *__restrict__ scanned, int *__restrict__ col_idx, float *__restrict__ values) {
    const long long b = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    if (b >= (long long)br * bc || !flags[b]) return;
    const int brow = (int)(b / bc), bcol = (int)(b % bc);
    const uint32_t blk = scanned[b];
    if (threadIdx.x == 0) col_idx[blk] = bcol;
    const float *p = dense + (size_t)brow * r * cols + (size_t)bcol * c;
    float *q = values + (size_t)blk * r * c;
    for (int e = threadIdx.x; e < r * c; e += blockDim.x) q[e] = __ldg(p + (size_t)(e / c) * cols + (e % c));
}

}  // namespace tsg

using namespace tsg;

extern "C" {

int tsg_tcsc_from_dense_f32(const float *dense_dev, int rows, int cols, tsg_tcsc **out) {
    return tcsc_from_dense_impl<float>(dense_dev, rows, cols, out);
}
int tsg_tcsc_from_dense_i32(const int *dense_dev, int rows, int cols, tsg_tcsc **out) {
    return tcsc_from_dense_impl<int>(dense_dev, rows, cols, out);
}

int tsg_bcsr_from_dense_f32(const float *dense, int rows, int cols, int r, int c, tsg_bcsr **out) {
    *out = nullptr;
    TSG_TRY(ensure_device());
    if (r <= 0 || c <= 0 || rows < 0 || cols < 0) return set_error(TSG_EINVAL, "bcsr_from_dense: bad arguments");
    cudaStream_t st = stream();
    tsg_bcsr *W = new (std::nothrow) tsg_bcsr();
    if (!W) return set_error(TSG_ENOMEM, "out of host memory");
    W->r = r; W->c = c; W->br = rows / r; W->bc = cols / c;  // bcsr.c:24-25
    const long long nb = (long long)W->br * W->bc;
    uint32_t *flags = nullptr, *scanned = nullptr, *total = nullptr;
    int rc;
    auto fail = [&](int code) {
        dev_free(flags); dev_free(scanned); dev_free(total);
        tsg_bcsr_destroy(W);
        return code;
    };
    if ((rc = dev_alloc_t(&flags, (size_t)nb + 1))) return fail(rc);
    if ((rc = dev_alloc_t(&scanned, (size_t)nb + 1))) return fail(rc);
    if ((rc = dev_alloc_t(&total, 1))) return fail(rc);
    if ((rc = dev_alloc_t(&W->row_start, (size_t)W->br + 1))) return fail(rc);
    if (nb > 0) {
        k_bcsr_flags<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(dense, cols, r, c, W->br, W->bc, flags);
        if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_bcsr_flags failed"));
        count_launch();
    }
    if ((rc = scan_exclusive_u32(flags, scanned, nb, total))) return fail(rc);
    uint32_t h_total = 0;
    if (cudaMemcpyAsync(&h_total, total, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return fail(set_error(TSG_ECUDA, "bcsr_from_dense: device failure: %s", cudaGetErrorString(cudaGetLastError())));
    W->k = (int)h_total;
    if ((rc = dev_alloc_t(&W->col_idx, (size_t)W->k))) return fail(rc);
    if ((rc = dev_alloc_t(&W->values, (size_t)W->k * r * c))) return fail(rc);
    if (nb > 0 && W->bc > 0) {
        k_bcsr_rowptr<<<(W->br + 1 + 255) / 256, 256, 0, st>>>(scanned, total, W->br, W->bc, W->row_start);
    } else {
        cudaMemsetAsync(W->row_start, 0, ((size_t)W->br + 1) * sizeof(int), st);
    }
    if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_bcsr_rowptr failed"));
    count_launch();
    if (W->k > 0) {
        const int tx = (r * c >= 32) ? 32 : (r * c >= 8 ? 8 : (r * c >= 4 ? 4 : 1));
        const int ty = 256 / tx;
        k_bcsr_fill<<<(unsigned)((nb + ty - 1) / ty), dim3(tx, ty), 0, st>>>(dense, cols, r, c, W->br, W->bc, flags, scanned,
                                                                             W->col_idx, W->values);
        if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_bcsr_fill failed"));
        count_launch();
    }
    dev_free(flags); dev_free(scanned); dev_free(total);
    *out = W;
    return TSG_OK;
}

}  // extern "C"
