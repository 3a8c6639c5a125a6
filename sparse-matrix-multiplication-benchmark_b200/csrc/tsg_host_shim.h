/* tsg_host_shim.h -- the thin C-ABI shim the plain-C host layer (api_tcsc.c, api_bcsr.c) calls.  Private. */
#ifndef TSG_HOST_SHIM_H
#define TSG_HOST_SHIM_H
#include <stddef.h>
#include "tsgemm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* 1 if p is a CUDA device (or managed) pointer, 0 if host memory */
int tsg_shim_is_device(const void *p);
/* make `bytes` at p visible on the device: device pointers pass through (*owned = 0); host pointers are copied into a
 * pool allocation on the current stream (*owned = 1, release with tsg_shim_release) */
int tsg_shim_stage_in(const void *p, size_t bytes, void **dev, int *owned);
/* device buffer for an output living at p: pass-through for device pointers, fresh allocation for host pointers */
int tsg_shim_stage_out_begin(void *p, size_t bytes, void **dev, int *owned);
/* copy back (if owned), synchronise the stream (if owned) and release */
int tsg_shim_stage_out_end(void *p, size_t bytes, void *dev, int owned);
int tsg_shim_release(void *dev, int owned);
/* Y = [PReLU](X*W + B) for any mix of host and device pointers.  Small host operands go through a per-thread pinned
 * arena (one H2D copy, Y written through mapped pinned memory, one synchronisation); larger ones through pool
 * allocations.  Returns after the caller's host buffers are no longer needed (and Y, if a host buffer, is complete);
 * all-device calls stay asynchronous. */
int tsg_shim_tcsc_gemm_staged(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N,
                              int K);
int tsg_shim_bcsr_gemm_staged(tsg_bcsr *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K);
/* host-pointer GEMM with the rows of X/Y pipelined over PCIe in slabs (H2D, kernel, D2H overlapped on 3 streams) */
int tsg_shim_tcsc_gemm_hostpipe(tsg_tcsc *W, const float *X_host, const float *B_any, float a, int use_prelu, int order,
                                float *Y_host, int M, int N, int K);
#ifdef __cplusplus
}
#endif
#endif
