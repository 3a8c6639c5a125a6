/*
 * sparse/tcsc.h -- Ternary Compressed Sparse Column: builder and multiplication-free GEMM entry points.
 *
 * Drop-in replacement for the reference's sparse/tcsc.h:6-48: same struct layout, same function names, same
 * argument order and types, same ownership rules (tcsc_from_dense mallocs, tcsc_free releases; struct fields
 * are plain host memory the caller may read, main.cpp:296).  Differences a maintainer should know:
 *
 *   - the functions run on an NVIDIA B200 (sm_100a).  There is no CPU fallback: without a usable CUDA device
 *     tcsc_from_dense returns NULL and the GEMM entry points leave Y untouched; sparse_last_error() says why.
 *   - X, B, Y and `dense` may be host pointers (staged over PCIe inside the call) or CUDA device pointers
 *     (used in place); the library detects which (cudaPointerGetAttributes).
 *   - every entry point reproduces the reference function's exact sequence of fp32 roundings per output element
 *     (see DESIGN.md "Summation order"), so results are bit-identical to the reference built with
 *     g++ -O3 -march=native (no -ffast-math) whenever M >= TSG_SKINNY_M; the skinny-M (decode) kernel sums a
 *     column's non-zeros in a tree and is within 1e-5 * max(|y|,1) of an fp64 accumulation instead.
 *   - declared extern "C" (the reference header has no linkage specification, so its own objects are C++-mangled;
 *     libtsgemm_b200.so also exports those mangled names, see csrc/api_cxx.cpp).
 */
#ifndef TSG_SPARSE_TCSC_H
#define TSG_SPARSE_TCSC_H

#include "../dense/dense.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference sparse/tcsc.h:6-17 */
typedef struct {
    int rows, cols;
    int n_elem_pos; /* number of matrix elements with value +1 */
    int n_elem_neg; /* number of matrix elements with value -1 */
    int *col_start_pos; /* cols+1 entries */
    int *col_start_neg; /* cols+1 entries */
    int *row_index_pos; /* n_elem_pos entries, ascending row inside each column */
    int *row_index_neg; /* n_elem_neg entries */
} tcsc_t;

/* reference sparse/tcsc.h:19, sparse/tcsc.c:6-66.  `dense` is rows x cols row-major; +1 iff ==1.0f, -1 iff ==-1.0f.
 * Returns NULL on allocation or device failure. */
tcsc_t *tcsc_from_dense(dense_t dense, int rows, int cols);

/* reference sparse/tcsc.h:21-24, sparse/tcsc.c:69-98:   Y = X*W + B, summed as  B, +pos..., -neg... */
void tcsc_sgemm_basic(const dense_t X, const tcsc_t *W, const dense_t B, dense_t Y, int M, int N, int K);

/* reference sparse/tcsc.h:26-29, sparse/tcsc.c:101-140: Y = (B + sum(pos)) - sum(neg) */
void tcsc_sgemm_optimized(const dense_t X, const tcsc_t *W, const dense_t B, dense_t Y, int M, int N, int K);

/* reference sparse/tcsc.h:31-34, sparse/tcsc.c:143-165: Y = PReLU(((0 +pos...) -neg...) + B),  PReLU(y) = y<0 ? a*y : y */
void tcsc_sgemm_prelu_basic(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y, int M, int N, int K);

/* reference sparse/tcsc.h:37-40, sparse/tcsc.c:179-227: PReLU((B + sum(pos)) - sum(neg)) */
void tcsc_sgemm_prelu_optimized_separate(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y,
                                         int M, int N, int K);

/* reference sparse/tcsc.h:43-46, sparse/tcsc.c:231-275: same value as the `separate` variant */
void tcsc_sgemm_prelu_optimized_onthego(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y,
                                        int M, int N, int K);

/* reference sparse/tcsc.h:48, sparse/tcsc.c:167-175; NULL-safe; also drops the cached device mirror of W */
void tcsc_free(tcsc_t *W);

/* extension: forget the device mirror cached for W.  The library notices a struct that was freed, rebuilt or edited
 * through W's dimensions, array pointers and a sampled content hash (TSG_MIRROR_CHECK=full hashes every word); call
 * this after editing the index arrays in place when neither would change. */
void tcsc_invalidate(const tcsc_t *W);

/* extension: message of the last failure on the calling thread ("" if none) -- the reference's GEMM entry points
 * return void, so this is the only error channel */
const char *sparse_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
