// ktformat.cu -- builds the private "K-tiled gather stream" of a TCSC matrix (device state hung off tsg_tcsc).
//
// The public TCSC arrays (sparse/tcsc.h:6-17) stay exactly as the reference defines them.  The GEMM kernel however
// keeps only a kc-row chunk of X in shared memory at a time, so it wants, for every (sign, chunk, column), the
// chunk-local row numbers of that column's entries, contiguous in memory for a tile of columns so that one bulk
// async copy (TMA) per pipeline stage brings them on chip.  Because each column's TCSC list is sorted by row
// (tcsc.c:51-59 appends rows in ascending order), the entries of a chunk are a contiguous sub-range of the list
// found by binary search, and chunk order = ascending-k order: walking chunks 0..nchunk-1 visits a column's
// entries in exactly the reference's summation order.
//
// Layout (see tsg::KStream): plane p = sign*nchunk + chunk.
//   cnt [p][ncols_pad]      uint8   16-byte quads (4 words = 16 entries) used by the column's list in this plane
//   woff[p*ngroup + g]      uint32  word offset into `body` of 8-column group g (always a multiple of 4 words = 16 B)
//   body[...]               uint32  entries: one byte per non-zero = k - chunk*kc + 1, every list padded with 0 to a whole
//                                   quad, so that the kernel fetches a list with ONE uniform 16-byte shared-memory load
//                                   per 16 entries (the index fetch competes with the gathers for the same crossbar)
// Size: ~1.6 bytes per non-zero at 90 % sparsity (vs 4 bytes in the int32 TCSC arrays).
#include "tsg_internal.h"

namespace tsg {

__device__ __forceinline__ int lower_bound_dev(const int *a, int lo, int hi, int key) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// one thread per (plane, column)
__global__ void k_ks_count(const int *__restrict__ csp, const int *__restrict__ csn, const int *__restrict__ rip,
                           const int *__restrict__ rin, int N, int ncols_pad, int kc, int nchunk, uint8_t *__restrict__ cnt) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = blockIdx.y;
    if (n >= ncols_pad) return;
    uint8_t words = 0;
    if (n < N) {
        const int sign = plane / nchunk, c = plane % nchunk;
        const int *cs = sign ? csn : csp;
        const int *ri = sign ? rin : rip;
        const int b = cs[n], e = cs[n + 1];
        const int lo = lower_bound_dev(ri, b, e, c * kc);
        const int hi = lower_bound_dev(ri, lo, e, (c + 1) * kc);
        words = (uint8_t)((hi - lo + 15) >> 4);  // quads
    }
    cnt[(size_t)plane * ncols_pad + n] = words;
}

// words of every 8-column group, rounded up to a multiple of 4 (16-byte granule for the bulk copies)
__global__ void k_ks_group_words(const uint8_t *__restrict__ cnt, long long ngroups_total, uint32_t *__restrict__ gwords) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups_total) return;
    const uint2 v = *reinterpret_cast<const uint2 *>(cnt + g * 8);
    uint32_t s = __vsadu4(v.x, 0u) + __vsadu4(v.y, 0u);  // sum of the 8 quad counts
    gwords[g] = s * 4u;
}

__global__ void k_ks_max_tile(const uint32_t *__restrict__ woff, int nplanes, int ngroup, int *__restrict__ max_words) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // 256-column tile index inside a plane
    const int plane = blockIdx.y;
    const int ntile = ngroup / 32;
    if (t >= ntile || plane >= nplanes) return;
    const size_t base = (size_t)plane * ngroup + (size_t)t * 32;
    atomicMax(max_words, (int)(woff[base + 32] - woff[base]));
}

// one thread per (plane, column): pack the chunk-local rows, 4 per word
__global__ void k_ks_fill(const int *__restrict__ csp, const int *__restrict__ csn, const int *__restrict__ rip,
                          const int *__restrict__ rin, int N, int ncols_pad, int ngroup, int kc, int nchunk,
                          const uint8_t *__restrict__ cnt, const uint32_t *__restrict__ woff, uint32_t *__restrict__ body) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = blockIdx.y;
    if (n >= N) return;
    const int sign = plane / nchunk, c = plane % nchunk;
    const int *cs = sign ? csn : csp;
    const int *ri = sign ? rin : rip;
    const int b = cs[n], e = cs[n + 1];
    const int lo = lower_bound_dev(ri, b, e, c * kc);
    const int hi = lower_bound_dev(ri, lo, e, (c + 1) * kc);
    const uint8_t *cp = cnt + (size_t)plane * ncols_pad + (n & ~7);
    uint32_t off = woff[(size_t)plane * ngroup + (n >> 3)];
    for (int i = 0; i < (n & 7); ++i) off += 4u * cp[i];
    const int k0 = c * kc;
    const int padded = ((hi - lo + 15) >> 4) << 4;  // whole quads; the tail is 0 (padding)
    for (int t = 0; t < padded; t += 4) {
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t kk = (lo + t + q < hi) ? (uint32_t)(ri[lo + t + q] - k0 + 1) : 0u;  // k+1; 0 = padding
            w |= kk << (8 * q);
        }
        body[off++] = w;
    }
}

static void free_kstream(KStream &ks) {
    dev_free(ks.cnt); dev_free(ks.woff); dev_free(ks.body);
    ks = KStream();
}

// shared-memory budget of the GEMM kernel (gemm_tcsc.cu): two pipeline stages must fit 227 KB
static constexpr long long kSmemTotal = 232448 - 1024;
static long long stage_bytes_for(int kc, int tile_words, int areas) { return (long long)kc * 512 + areas * ((long long)tile_words * 4 + 256 + 160); }

int build_kstream(tsg_tcsc *W, int smem_reserved, int areas, int nstage) {
    // smem_reserved: shared memory the kernel keeps for something else (the separate output tile of dist mode 4); the
    // chunk height is part of the stream's layout, so a different reservation means a different stream.
    // areas: stream slices per pipeline stage -- 1 for the exact orders (a stage holds the +1 OR the -1 lists of a chunk), 2 for
    // TSG_ORDER_FAST (both); the two layouts differ in the chunk height only and live side by side (ks / ks_fast)
    std::lock_guard<std::mutex> lk(W->mu);
    KStream &slot = (areas == 2) ? W->ks_fast : W->ks;
    if (slot.built && slot.smem_reserved == smem_reserved && slot.nstage == nstage) return TSG_OK;
    if (slot.built) {
        free_kstream(slot);
        slot = KStream();
    }
    const long long kSmemBudget = kSmemTotal - smem_reserved;
    cudaStream_t st = stream();
    const int K = W->rows, N = W->cols;
    const long long nnz = (long long)W->n_pos + W->n_neg;
    const double density = (K > 0 && N > 0) ? (double)nnz / ((double)K * N) : 0.0;
    // first guess for kc from the expected tile size, then verify against the real maximum and shrink if needed
    int kc = 224;
    for (; kc > 16; kc -= 8) {
        double words_per_list = kc * density * 0.5 / 4.0 + 2.0;  // lists are padded to whole 4-word quads
        long long tile_words = (long long)(256 * words_per_list * 1.15) + 64;
        if (nstage * stage_bytes_for(kc, (int)tile_words, areas) <= kSmemBudget) break;
    }
    if (kc > K) kc = (K > 0) ? K : 1;
    for (int attempt = 0; attempt < 12; ++attempt) {
        KStream ks;
        ks.smem_reserved = smem_reserved;
        ks.nstage = nstage;
        ks.kc = kc;
        ks.nchunk = (K + kc - 1) / kc;
        if (ks.nchunk < 1) ks.nchunk = 1;
        ks.ncols_pad = ((N + 255) / 256) * 256;
        if (ks.ncols_pad == 0) ks.ncols_pad = 256;
        ks.ngroup = ks.ncols_pad / 8;
        const int nplanes = 2 * ks.nchunk;
        const long long ngroups_total = (long long)nplanes * ks.ngroup;
        ks.body_words = nnz / 4 + 4LL * nplanes * ks.ncols_pad + 64;  // every list rounds up by < 4 words
        int rc;
        uint32_t *gwords = nullptr, *total = nullptr;
        int *maxw = nullptr;
        bool ws_held = false;
        auto fail = [&](int code) {
            if (ws_held) ws_release(3);
            free_kstream(ks);
            return code;
        };
        if ((rc = dev_alloc_t(&ks.cnt, (size_t)nplanes * ks.ncols_pad))) return fail(rc);
        if ((rc = dev_alloc_t(&ks.woff, (size_t)ngroups_total + 8))) return fail(rc);
        if ((rc = dev_alloc_t(&ks.body, (size_t)ks.body_words))) return fail(rc);
        {   // scratch from the persistent workspace (no pool churn when matrices are converted again and again)
            const size_t gw_bytes = ((((size_t)ngroups_total + 8) * 4) + 255) & ~(size_t)255;
            void *base = nullptr;
            if ((rc = ws_acquire(3, gw_bytes + 512, &base))) return fail(rc);
            ws_held = true;
            gwords = static_cast<uint32_t *>(base);
            total = reinterpret_cast<uint32_t *>(static_cast<char *>(base) + gw_bytes);
            maxw = reinterpret_cast<int *>(static_cast<char *>(base) + gw_bytes + 256);
        }
        cudaMemsetAsync(maxw, 0, sizeof(int), st);
        cudaMemsetAsync(ks.body, 0, (size_t)ks.body_words * 4, st);
        cudaMemsetAsync(gwords, 0, ((size_t)ngroups_total + 8) * 4, st);
        {
            dim3 grid((ks.ncols_pad + 255) / 256, nplanes);
            k_ks_count<<<grid, 256, 0, st>>>(W->csp, W->csn, W->rip, W->rin, N, ks.ncols_pad, kc, ks.nchunk, ks.cnt);
            if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_ks_count failed"));
            count_launch();
        }
        k_ks_group_words<<<(unsigned)((ngroups_total + 255) / 256), 256, 0, st>>>(ks.cnt, ngroups_total, gwords);
        if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_ks_group_words failed"));
        count_launch();
        // scan over ngroups_total + 8 entries so that woff[ngroups_total .. +7] all hold the grand total
        if ((rc = scan_exclusive_u32(gwords, ks.woff, ngroups_total + 8, total))) return fail(rc);
        {
            dim3 grid((ks.ngroup / 32 + 63) / 64, nplanes);
            k_ks_max_tile<<<grid, 64, 0, st>>>(ks.woff, nplanes, ks.ngroup, maxw);
            if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_ks_max_tile failed"));
            count_launch();
        }
        if (N > 0 && nnz > 0) {
            dim3 grid((N + 255) / 256, nplanes);
            k_ks_fill<<<grid, 256, 0, st>>>(W->csp, W->csn, W->rip, W->rin, N, ks.ncols_pad, ks.ngroup, kc, ks.nchunk, ks.cnt,
                                            ks.woff, ks.body);
            if (cudaGetLastError() != cudaSuccess) return fail(set_error(TSG_ECUDA, "launch of k_ks_fill failed"));
            count_launch();
        }
        int h_max = 0;
        uint32_t h_total = 0;
        if (cudaMemcpyAsync(&h_max, maxw, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(&h_total, total, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return fail(set_error(TSG_ECUDA, "build_kstream: device failure: %s", cudaGetErrorString(cudaGetLastError())));
        ws_release(3);
        ws_held = false;
        if ((long long)h_total > ks.body_words) {
            free_kstream(ks);
            return set_error(TSG_ECUDA, "build_kstream: internal size bound violated (%u > %lld)", h_total, ks.body_words);
        }
        ks.max_tile_words = h_max;
        if (nstage * stage_bytes_for(kc, h_max, areas) <= kSmemBudget || kc <= 8) {
            ks.built = true;
            slot = ks;
            return TSG_OK;
        }
        free_kstream(ks);  // a tile's index block is larger than estimated: shrink the chunk and retry
        kc = (int)(kc * 0.8);
        if (kc < 8) kc = 8;
    }
    return set_error(TSG_EUNSUPPORTED, "build_kstream: could not fit the gather stream into shared memory");
}

}  // namespace tsg
