/*
 * common.h -- the plugin seam: function-pointer types every kernel entry point conforms to.
 * Drop-in for the reference's common.h:5-14 (its add_func template, :16-17, is declared but never defined or
 * called anywhere in the reference and is not reproduced).
 */
#ifndef TSG_COMMON_H
#define TSG_COMMON_H

#include "dense/dense.h"

/* reference common.h:6-9 */
typedef void (*gemm_func)(const dense_t X, const void *W, const dense_t B, dense_t Y, int M, int N, int K);
/* reference common.h:11-14 */
typedef void (*prelu_func)(const dense_t X, const void *W, const dense_t B, float a, dense_t Y, int M, int N, int K);

#endif
