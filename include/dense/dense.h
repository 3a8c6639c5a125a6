/*
 * dense/dense.h -- element typedefs shared by the sparse entry points.
 *
 * Drop-in for the typedef part of the reference's dense/dense.h:5-6.  Unlike the reference header (which includes
 * <cstdbool> and therefore only compiles as C++), this one is valid C11 and C++17.
 *
 * The reference's dense helpers (init_rand_dense, init_rand_sparse, compare, gemm_basic: dense/dense.h:10-21) are
 * benchmark/test support that stays on the CPU; they are not part of this library.
 */
#ifndef TSG_DENSE_DENSE_H
#define TSG_DENSE_DENSE_H

typedef float dense_elem_t;    /* dense/dense.h:5 */
typedef dense_elem_t *dense_t; /* dense/dense.h:6 */

#endif
