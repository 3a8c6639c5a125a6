/*
 * dense/dense.h -- element typedefs shared by the sparse entry points.
 *
 * Drop-in for the typedef part of the reference's dense/dense.h:5-6.  Unlike the reference header (which includes
 * <cstdbool> and therefore only compiles as C++), this one is valid C11 and C++17.
 *
 * The reference's dense helpers (init_rand_dense, init_rand_sparse, compare, gemm_basic: dense/dense.h:10-21) are
 * benchmark/test support that stays on the CPU; they are declared below for C++ callers but
 * are not part of this library.
 */
#ifndef TSG_DENSE_DENSE_H
#define TSG_DENSE_DENSE_H

typedef float dense_elem_t;    /* dense/dense.h:5 */
typedef dense_elem_t *dense_t; /* dense/dense.h:6 */

#ifdef __cplusplus
/* Declarations only, so that callers which include "dense/dense.h" for the reference's CPU helpers (main.cpp:278-280,
 * 307,317; test/test_bcsr.cpp:20-22,30,36) still compile after a header swap.  The definitions keep coming from the
 * reference's own dense/dense.c, which every documented build compiles with g++ -- hence C++ linkage here. */
bool compare(const dense_t result, const dense_t target, int rows, int cols);                     /* dense/dense.h:10 */
dense_t init_rand_dense(int rows, int cols);                                                     /* dense/dense.h:12 */
dense_t init_rand_sparse(int rows, int cols, int non_zero);                                      /* dense/dense.h:13 */
void gemm_basic(const dense_t X, const dense_t W, const dense_t B, dense_t Y, int M, int N, int K); /* dense/dense.h:18-21 */
#endif

#endif
