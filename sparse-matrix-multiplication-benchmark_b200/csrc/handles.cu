// handles.cu -- lifetime, import/export of the device mirrors (tsg_tcsc / tsg_bcsr).
#include "tsg_internal.h"

namespace tsg {

int is_device_pointer(const void *p) {
    if (!p) return 0;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 1 : 0;
}

static int copy_in(void *dst_dev, const void *src, size_t bytes) {
    if (bytes == 0) return TSG_OK;
    TSG_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, is_device_pointer(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream()));
    return TSG_OK;
}

// one warp per column: col_start monotone and inside [0, nnz], rows inside [0, K) and ascending inside the column (the
// gather stream finds a chunk's entries by binary search, ktformat.cu).  bad: bit 0 pointers, bit 1 row range, bit 2 order
__global__ void k_validate_tcsc(const int *__restrict__ cs, const int *__restrict__ ri, int N, int K, int nnz, int *__restrict__ bad) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const int b = cs[n], e = cs[n + 1];
    int f = 0;
    if ((n == 0 && b != 0) || b > e || b < 0 || e > nnz) f |= 1;
    else
        for (int t = b + lane; t < e; t += 32) {
            const int k = ri[t];
            if (k < 0 || k >= K) f |= 2;
            if (t > b && ri[t - 1] > k) f |= 4;
        }
    if (f) atomicOr(bad, f);
}

// thread per block-column: count (pass 0) or record (pass 1) the blocks of that column in ascending block-row order
__global__ void k_bcsr_cols(const int *__restrict__ row_start, const int *__restrict__ col_idx, int br, int bc, int pass,
                            int *__restrict__ cptr, int *__restrict__ crow, int *__restrict__ cblk) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= bc) return;
    int w = pass ? cptr[col] : 0;
    for (int brow = 0; brow < br; ++brow) {
        int lo = row_start[brow], hi = row_start[brow + 1];
        while (lo < hi) {  // col_idx is ascending inside a block-row (bcsr.c:105-119 walks bcol upwards)
            int mid = (lo + hi) >> 1;
            if (col_idx[mid] < col) lo = mid + 1;
            else hi = mid;
        }
        if (lo < row_start[brow + 1] && col_idx[lo] == col) {
            if (pass) { crow[w] = brow; cblk[w] = lo; }
            ++w;
        }
    }
    if (!pass) cptr[col] = w;
}

int bcsr_build_cols(tsg_bcsr *W) {
    std::lock_guard<std::recursive_mutex> lk(W->mu);
    if (W->col_built) return TSG_OK;
    cudaStream_t st = stream();
    TSG_TRY(dev_alloc_t(&W->cptr, (size_t)W->bc + 2));
    TSG_TRY(dev_alloc_t(&W->crow, (size_t)W->k + 1));
    TSG_TRY(dev_alloc_t(&W->cblk, (size_t)W->k + 1));
    uint32_t *counts = nullptr, *total = nullptr;
    TSG_TRY(dev_alloc_t(&counts, (size_t)W->bc + 2));
    TSG_TRY(dev_alloc_t(&total, 1));
    TSG_CUDA(cudaMemsetAsync(counts, 0, ((size_t)W->bc + 2) * 4, st));
    if (W->bc > 0) {
        k_bcsr_cols<<<(W->bc + 127) / 128, 128, 0, st>>>(W->row_start, W->col_idx, W->br, W->bc, 0, reinterpret_cast<int *>(counts), nullptr, nullptr);
        TSG_KERNEL_CHECK("k_bcsr_cols");
    }
    TSG_TRY(scan_exclusive_u32(counts, reinterpret_cast<uint32_t *>(W->cptr), (long long)W->bc + 1, total));
    if (W->bc > 0) {
        k_bcsr_cols<<<(W->bc + 127) / 128, 128, 0, st>>>(W->row_start, W->col_idx, W->br, W->bc, 1, W->cptr, W->crow, W->cblk);
        TSG_KERNEL_CHECK("k_bcsr_cols");
    }
    dev_free(counts);
    dev_free(total);
    W->col_built = true;
    return TSG_OK;
}

}  // namespace tsg

using namespace tsg;

extern "C" {

int tsg_tcsc_from_arrays(const int *csp, const int *csn, const int *rip, const int *rin, int rows, int cols, tsg_tcsc **out) {
    *out = nullptr;
    TSG_TRY(ensure_device());
    if (rows < 0 || cols < 0 || !csp || !csn) return set_error(TSG_EINVAL, "tsg_tcsc_from_arrays: bad arguments");
    tsg_tcsc *W = new (std::nothrow) tsg_tcsc();
    if (!W) return set_error(TSG_ENOMEM, "out of host memory");
    W->rows = rows;
    W->cols = cols;
    int rc;
    auto fail = [&](int code) { tsg_tcsc_destroy(W); return code; };
    if ((rc = dev_alloc_t(&W->csp, (size_t)cols + 1)) || (rc = dev_alloc_t(&W->csn, (size_t)cols + 1))) return fail(rc);
    if ((rc = copy_in(W->csp, csp, ((size_t)cols + 1) * 4)) || (rc = copy_in(W->csn, csn, ((size_t)cols + 1) * 4))) return fail(rc);
    int tot[2];
    if (cudaMemcpyAsync(&tot[0], W->csp + cols, 4, cudaMemcpyDeviceToHost, stream()) != cudaSuccess ||
        cudaMemcpyAsync(&tot[1], W->csn + cols, 4, cudaMemcpyDeviceToHost, stream()) != cudaSuccess ||
        cudaStreamSynchronize(stream()) != cudaSuccess)
        return fail(set_error(TSG_ECUDA, "tsg_tcsc_from_arrays: %s", cudaGetErrorString(cudaGetLastError())));
    W->n_pos = tot[0];
    W->n_neg = tot[1];
    if (W->n_pos < 0 || W->n_neg < 0 || (W->n_pos > 0 && !rip) || (W->n_neg > 0 && !rin))
        return fail(set_error(TSG_EINVAL, "tsg_tcsc_from_arrays: inconsistent column pointers"));
    if ((rc = dev_alloc_t(&W->rip, (size_t)W->n_pos)) || (rc = dev_alloc_t(&W->rin, (size_t)W->n_neg))) return fail(rc);
    if ((rc = copy_in(W->rip, rip, (size_t)W->n_pos * 4)) || (rc = copy_in(W->rin, rin, (size_t)W->n_neg * 4))) return fail(rc);
    // caller-assembled arrays are checked once, here: the kernels trust them afterwards
    if (cols > 0) {
        int *bad = nullptr;
        if ((rc = dev_alloc_t(&bad, 1))) return fail(rc);
        int h_bad = 0;
        cudaMemsetAsync(bad, 0, 4, stream());
        const unsigned grid = (unsigned)(((size_t)cols * 32 + 255) / 256);
        k_validate_tcsc<<<grid, 256, 0, stream()>>>(W->csp, W->rip, cols, rows, W->n_pos, bad);
        k_validate_tcsc<<<grid, 256, 0, stream()>>>(W->csn, W->rin, cols, rows, W->n_neg, bad);
        const bool ok = cudaGetLastError() == cudaSuccess && cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, stream()) == cudaSuccess &&
                        cudaStreamSynchronize(stream()) == cudaSuccess;
        dev_free(bad);
        if (!ok) return fail(set_error(TSG_ECUDA, "tsg_tcsc_from_arrays: validation failed to run: %s", cudaGetErrorString(cudaGetLastError())));
        count_launch();
        count_launch();
        if (h_bad)
            return fail(set_error(TSG_EINVAL, "tsg_tcsc_from_arrays: invalid TCSC arrays:%s%s%s", (h_bad & 1) ? " column pointers not monotone from 0" : "",
                                  (h_bad & 2) ? " row index outside [0, rows)" : "",
                                  (h_bad & 4) ? " rows not ascending inside a column (tcsc_from_dense emits them ascending, tcsc.c:51-59)" : ""));
    }
    *out = W;
    return TSG_OK;
}

void tsg_tcsc_destroy(tsg_tcsc *W) {
    if (!W) return;
    dev_free(W->csp); dev_free(W->csn); dev_free(W->rip); dev_free(W->rin);
    dev_free(W->ks.cnt); dev_free(W->ks.woff); dev_free(W->ks.body);
    dev_free(W->ks_fast.cnt); dev_free(W->ks_fast.woff); dev_free(W->ks_fast.body);
    dev_free(W->w2);
    delete W;
}

int tsg_tcsc_dims(const tsg_tcsc *W, int *rows, int *cols, int *n_pos, int *n_neg) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    if (rows) *rows = W->rows;
    if (cols) *cols = W->cols;
    if (n_pos) *n_pos = W->n_pos;
    if (n_neg) *n_neg = W->n_neg;
    return TSG_OK;
}

int tsg_tcsc_download(const tsg_tcsc *W, int *csp, int *csn, int *rip, int *rin) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    cudaStream_t st = stream();
    TSG_CUDA(cudaMemcpyAsync(csp, W->csp, ((size_t)W->cols + 1) * 4, cudaMemcpyDeviceToHost, st));
    TSG_CUDA(cudaMemcpyAsync(csn, W->csn, ((size_t)W->cols + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (W->n_pos) TSG_CUDA(cudaMemcpyAsync(rip, W->rip, (size_t)W->n_pos * 4, cudaMemcpyDeviceToHost, st));
    if (W->n_neg) TSG_CUDA(cudaMemcpyAsync(rin, W->rin, (size_t)W->n_neg * 4, cudaMemcpyDeviceToHost, st));
    TSG_CUDA(cudaStreamSynchronize(st));
    return TSG_OK;
}

int tsg_tcsc_device_arrays(const tsg_tcsc *W, const int **csp, const int **csn, const int **rip, const int **rin) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    *csp = W->csp; *csn = W->csn; *rip = W->rip; *rin = W->rin;
    return TSG_OK;
}

int tsg_tcsc_stream_info(tsg_tcsc *W, long long *bytes, int *kc, int *nchunk) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    TSG_TRY(ensure_device());
    TSG_TRY(build_kstream(W));
    if (bytes) *bytes = W->ks.body_words * 4 + (long long)2 * W->ks.nchunk * (W->ks.ncols_pad + 4LL * W->ks.ngroup);
    if (kc) *kc = W->ks.kc;
    if (nchunk) *nchunk = W->ks.nchunk;
    return TSG_OK;
}

int tsg_bcsr_from_arrays(const int *row_start, const int *col_idx, const float *values, int r, int c, int br, int bc, int k, tsg_bcsr **out) {
    *out = nullptr;
    TSG_TRY(ensure_device());
    if (r <= 0 || c <= 0 || br < 0 || bc < 0 || k < 0 || !row_start) return set_error(TSG_EINVAL, "tsg_bcsr_from_arrays: bad arguments");
    tsg_bcsr *W = new (std::nothrow) tsg_bcsr();
    if (!W) return set_error(TSG_ENOMEM, "out of host memory");
    W->r = r; W->c = c; W->br = br; W->bc = bc; W->k = k;
    int rc;
    auto fail = [&](int code) { tsg_bcsr_destroy(W); return code; };
    if ((rc = dev_alloc_t(&W->row_start, (size_t)br + 1)) || (rc = dev_alloc_t(&W->col_idx, (size_t)k)) ||
        (rc = dev_alloc_t(&W->values, (size_t)k * r * c)))
        return fail(rc);
    if ((rc = copy_in(W->row_start, row_start, ((size_t)br + 1) * 4)) || (rc = copy_in(W->col_idx, col_idx, (size_t)k * 4)) ||
        (rc = copy_in(W->values, values, (size_t)k * r * c * 4)))
        return fail(rc);
    if (cudaStreamSynchronize(stream()) != cudaSuccess) return fail(set_error(TSG_ECUDA, "tsg_bcsr_from_arrays: %s", cudaGetErrorString(cudaGetLastError())));
    *out = W;
    return TSG_OK;
}

void tsg_bcsr_destroy(tsg_bcsr *W) {
    if (!W) return;
    dev_free(W->row_start); dev_free(W->col_idx); dev_free(W->values);
    dev_free(W->cptr); dev_free(W->crow); dev_free(W->cblk); dev_free(W->cval);
    dev_free(W->bs.cnt); dev_free(W->bs.wstart); dev_free(W->bs.eoff); dev_free(W->bs.hdr); dev_free(W->bs.val);
    delete W;
}

int tsg_bcsr_dims(const tsg_bcsr *W, int *r, int *c, int *br, int *bc, int *k) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    if (r) *r = W->r;
    if (c) *c = W->c;
    if (br) *br = W->br;
    if (bc) *bc = W->bc;
    if (k) *k = W->k;
    return TSG_OK;
}

int tsg_bcsr_download(const tsg_bcsr *W, int *row_start, int *col_idx, float *values) {
    if (!W) return set_error(TSG_EINVAL, "null handle");
    cudaStream_t st = stream();
    TSG_CUDA(cudaMemcpyAsync(row_start, W->row_start, ((size_t)W->br + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (W->k) {
        TSG_CUDA(cudaMemcpyAsync(col_idx, W->col_idx, (size_t)W->k * 4, cudaMemcpyDeviceToHost, st));
        TSG_CUDA(cudaMemcpyAsync(values, W->values, (size_t)W->k * W->r * W->c * 4, cudaMemcpyDeviceToHost, st));
    }
    TSG_CUDA(cudaStreamSynchronize(st));
    return TSG_OK;
}

}  // extern "C"
