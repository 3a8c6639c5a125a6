// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
//
// extern "C" trampoline around the UNMODIFIED reference, compiled in place from /root/reference
// (see oracle/Makefile, target `ref`).  Nothing from the reference is copied into this repo: this file
// only #includes its headers via -I/root/reference and forwards calls, so that the Python tests and the
// `bench.py --impl reference` arm can reach the reference's C++-mangled symbols (the reference's headers
// have no extern "C"; sparse/tcsc.h:19-48, sparse/bcsr.h:14-39, SparseGEMM.h:13-40,104-168).
//
// The resulting oracle/_ref/libref_oracle.so is git-ignored and travels to the GPU box with the snapshot.
#include <cstdlib>
#include <cstring>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "dense/dense.h"
#include "sparse/tcsc.h"
#include "sparse/bcsr.h"
#include "SparseGEMM.h"

extern "C" {

// ---- TCSC (sparse/tcsc.h:19-48) --------------------------------------------------------------------------
void *ref_tcsc_from_dense(float *dense, int rows, int cols) { return tcsc_from_dense(dense, rows, cols); }
void ref_tcsc_free(void *W) { tcsc_free(static_cast<tcsc_t *>(W)); }
int ref_tcsc_n_pos(const void *W) { return static_cast<const tcsc_t *>(W)->n_elem_pos; }
int ref_tcsc_n_neg(const void *W) { return static_cast<const tcsc_t *>(W)->n_elem_neg; }
const int *ref_tcsc_col_start_pos(const void *W) { return static_cast<const tcsc_t *>(W)->col_start_pos; }
const int *ref_tcsc_col_start_neg(const void *W) { return static_cast<const tcsc_t *>(W)->col_start_neg; }
const int *ref_tcsc_row_index_pos(const void *W) { return static_cast<const tcsc_t *>(W)->row_index_pos; }
const int *ref_tcsc_row_index_neg(const void *W) { return static_cast<const tcsc_t *>(W)->row_index_neg; }

void ref_tcsc_sgemm_basic(float *X, const void *W, float *B, float *Y, int M, int N, int K) {
    tcsc_sgemm_basic(X, static_cast<const tcsc_t *>(W), B, Y, M, N, K);
}
void ref_tcsc_sgemm_optimized(float *X, const void *W, float *B, float *Y, int M, int N, int K) {
    tcsc_sgemm_optimized(X, static_cast<const tcsc_t *>(W), B, Y, M, N, K);
}
void ref_tcsc_sgemm_prelu_basic(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K) {
    tcsc_sgemm_prelu_basic(X, static_cast<const tcsc_t *>(W), B, a, Y, M, N, K);
}
void ref_tcsc_sgemm_prelu_optimized_separate(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K) {
    tcsc_sgemm_prelu_optimized_separate(X, static_cast<const tcsc_t *>(W), B, a, Y, M, N, K);
}
void ref_tcsc_sgemm_prelu_optimized_onthego(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K) {
    tcsc_sgemm_prelu_optimized_onthego(X, static_cast<const tcsc_t *>(W), B, a, Y, M, N, K);
}

// ---- dense helpers (dense/dense.h:10-21) -----------------------------------------------------------------
void ref_gemm_basic(float *X, float *W, float *B, float *Y, int M, int N, int K) { gemm_basic(X, W, B, Y, M, N, K); }
int ref_compare(float *result, float *target, int rows, int cols) { return compare(result, target, rows, cols) ? 1 : 0; }

// ---- BCSR (sparse/bcsr.h:14-39) --------------------------------------------------------------------------
void *ref_bcsr_from_dense(float *dense, int rows, int cols, int r, int c) { return bcsr_from_dense(dense, rows, cols, r, c); }
void ref_bcsr_free(void *Wv) {  // ownership rule of test/test_bcsr.cpp:48-51: caller frees each array + struct
    bcsr_t *W = static_cast<bcsr_t *>(Wv);
    if (!W) return;
    free(W->b_values); free(W->b_row_start); free(W->b_col_idx); free(W);
}
void ref_bcsr_dims(const void *Wv, int *out5) {
    const bcsr_t *W = static_cast<const bcsr_t *>(Wv);
    out5[0] = W->r; out5[1] = W->c; out5[2] = W->br; out5[3] = W->bc; out5[4] = W->k;
}
const int *ref_bcsr_row_start(const void *W) { return static_cast<const bcsr_t *>(W)->b_row_start; }
const int *ref_bcsr_col_idx(const void *W) { return static_cast<const bcsr_t *>(W)->b_col_idx; }
const float *ref_bcsr_values(const void *W) { return static_cast<const bcsr_t *>(W)->b_values; }
void ref_bcsr_sgemm_basic(float *X, const void *W, float *B, float *Y, int M, int N, int K) {
    bcsr_sgemm_basic(X, *static_cast<const bcsr_t *>(W), B, Y, M, N, K);
}
void ref_bcsr_sgemm_prelu_basic(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K) {
    bcsr_sgemm_prelu_basic(X, *static_cast<const bcsr_t *>(W), B, a, Y, M, N, K);
}
void ref_bcsr_sgemm_avx(float *X, const void *W, float *B, float *Y, int M, int N, int K) {
    bcsr_sgemm_avx(X, *static_cast<const bcsr_t *>(W), B, Y, M, N, K);
}
void ref_bcsr_sgemm_prelu_avx(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K) {
    bcsr_sgemm_prelu_avx(X, *static_cast<const bcsr_t *>(W), B, a, Y, M, N, K);
}
void ref_bcsr_sgemm_avx2(float *X, const void *W, float *B, float *Y, int M, int N, int K) {
    bcsr_sgemm_avx2(X, *static_cast<const bcsr_t *>(W), B, Y, M, N, K);
}

// ---- SparseGEMM.h (class SparseFormat :13-40, sparseGEMM<T> :104-119, sparseGEMM_PReLU<T> :151-168) ---------
void *ref_sparseformat_new(int *matrix, int K, int N) { return new SparseFormat(matrix, K, N); }
void ref_sparseformat_delete(void *sf) { delete static_cast<SparseFormat *>(sf); }
int ref_sparseformat_sizes(const void *sfv, int *out4) {
    const SparseFormat *sf = static_cast<const SparseFormat *>(sfv);
    out4[0] = (int)sf->col_start_pos.size(); out4[1] = (int)sf->col_start_neg.size();
    out4[2] = (int)sf->row_index_pos.size(); out4[3] = (int)sf->row_index_neg.size();
    return 0;
}
const int *ref_sparseformat_array(void *sfv, int which) {
    SparseFormat *sf = static_cast<SparseFormat *>(sfv);
    switch (which) {
        case 0: return sf->col_start_pos.data();
        case 1: return sf->col_start_neg.data();
        case 2: return sf->row_index_pos.data();
        default: return sf->row_index_neg.data();
    }
}
void ref_sparseGEMM_f32(float *X, int *csp, int *csn, int *rip, int *rin, float *b, float *Y, int M, int N, int K) {
    sparseGEMM<float>(X, csp, csn, rip, rin, b, Y, M, N, K);
}
void ref_sparseGEMM_PReLU_f32(float *X, int *csp, int *csn, int *rip, int *rin, float *b, float *Y, int M, int N, int K, float a) {
    sparseGEMM_PReLU<float>(X, csp, csn, rip, rin, b, Y, M, N, K, a);
}
void ref_GEMM_f32(float *X, float *W, float *b, float *Y, int M, int N, int K) { GEMM<float>(X, W, b, Y, M, N, K); }
void ref_GEMM_PReLU_f32(float *X, float *W, float *b, float *Y, int M, int N, int K, float a) { GEMM_PReLU<float>(X, W, b, Y, M, N, K, a); }

// ---- timing helper for bench.py --impl reference / cpu_baseline: best-of-`reps` steady_clock seconds -------
double ref_time_tcsc_sgemm_prelu_basic(float *X, const void *W, float *B, float a, float *Y, int M, int N, int K, int reps) {
    double best = 1e300;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        tcsc_sgemm_prelu_basic(X, static_cast<const tcsc_t *>(W), B, a, Y, M, N, K);
        auto t1 = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        if (s < best) best = s;
    }
    return best;
}

// the reference's multi-threaded form of the same math: sparseGEMM_PReLU<float> carries `#pragma omp parallel for` over m
// (SparseGEMM.h:151-168); only the _omp build of this wrapper (-fopenmp) runs it on more than one thread
double ref_time_sparseGEMM_PReLU_f32(float *X, int *csp, int *csn, int *rip, int *rin, float *b, float *Y, int M, int N, int K, float a,
                                     int reps) {
    double best = 1e300;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        sparseGEMM_PReLU<float>(X, csp, csn, rip, rin, b, Y, M, N, K, a);
        auto t1 = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        if (s < best) best = s;
    }
    return best;
}
int ref_omp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

const char *ref_build_flags(void) {
#ifdef REF_BUILD_FLAGS
    return REF_BUILD_FLAGS;
#else
    return "unknown";
#endif
}

}  // extern "C"
