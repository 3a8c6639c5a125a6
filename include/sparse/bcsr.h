/*
 * sparse/bcsr.h -- Block CSR (dense r x c float blocks) builder and GEMM entry points.
 *
 * Drop-in replacement for the reference's sparse/bcsr.h:5-39.  Same struct, names, argument order; W is passed BY
 * VALUE like the reference does.  Ownership follows test/test_bcsr.cpp:48-51: bcsr_from_dense allocates the three
 * arrays and the struct with aligned_alloc(32, ...) and the CALLER frees each of them with free(); call
 * bcsr_release_device(W) first if you want the cached device mirror dropped as well (it is also dropped at exit).
 *
 * Semantics that differ from the reference on purpose:
 *   - b_row_start is a standard CSR pointer array (br+1 entries, empty block-rows repeat the previous value).  The
 *     reference only appends an entry for block-rows that own a block (bcsr.c:114-117) and leaves the tail
 *     uninitialised; both agree whenever no block-row is empty.
 *   - bcsr_sgemm_prelu_* return PReLU(X*W + B) with PReLU(y) = y<0 ? a*y : y by default.  The reference applies the
 *     activation after every partial update (bcsr.c:208-212, 300-306), which is not that function;
 *     tsg_bcsr_set_prelu_literal(1) (tsgemm_b200.h) or TSG_BCSR_PRELU_LITERAL=1 makes these two entry points return the
 *     reference's literal result instead, bit for bit.
 *   - the *_avx / *_avx2 names are aliases of the one CUDA kernel; their alignment and c==8 / r==c==8
 *     preconditions (bcsr.c:229-230, 316) are not required here.
 * The __restrict qualifier on the by-value struct parameter of the reference prototypes is a g++-ism with no ABI
 * effect and is omitted so that this header is valid C.
 */
#ifndef TSG_SPARSE_BCSR_H
#define TSG_SPARSE_BCSR_H

#include "../dense/dense.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef float bcsr_elem_t; /* sparse/bcsr.h:5 */

/* reference sparse/bcsr.h:7-12 */
typedef struct {
    int r, c, br, bc, k;
    int *b_row_start;
    int *b_col_idx;
    bcsr_elem_t *b_values;
} bcsr_t;

/* reference sparse/bcsr.h:14, sparse/bcsr.c:19-139.  br = rows/r, bc = cols/c (remainders dropped).  Returns NULL on
 * failure (the reference calls exit(); a library must not). */
bcsr_t *bcsr_from_dense(dense_t dense, int rows, int cols, int r, int c);

/* reference sparse/bcsr.h:16-19, sparse/bcsr.c:141-175 */
void bcsr_sgemm_basic(const dense_t __restrict X, const bcsr_t W, const dense_t __restrict B, dense_t __restrict Y,
                      int M, int N, int K);
/* reference sparse/bcsr.h:21-24 */
void bcsr_sgemm_prelu_basic(const dense_t __restrict X, const bcsr_t W, const dense_t __restrict B, float a,
                            dense_t __restrict Y, int M, int N, int K);
/* reference sparse/bcsr.h:26-29 */
void bcsr_sgemm_avx(const dense_t __restrict X, const bcsr_t W, const dense_t __restrict B, dense_t __restrict Y,
                    int M, int N, int K);
/* reference sparse/bcsr.h:31-34 */
void bcsr_sgemm_prelu_avx(const dense_t __restrict X, const bcsr_t W, const dense_t __restrict B, float a,
                          dense_t __restrict Y, int M, int N, int K);
/* reference sparse/bcsr.h:36-39 */
void bcsr_sgemm_avx2(const dense_t __restrict X, const bcsr_t W, const dense_t __restrict B, dense_t __restrict Y,
                     int M, int N, int K);

/* extension: drop the device mirror cached for W (keyed by W->b_values) */
void bcsr_release_device(const bcsr_t *W);

#ifdef __cplusplus
}
#endif
#endif
