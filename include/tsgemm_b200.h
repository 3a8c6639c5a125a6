/*
 * tsgemm_b200.h -- device-level C-ABI of libtsgemm_b200.so (extensions beyond the reference's headers).
 *
 * The reference-named entry points (sparse/tcsc.h, sparse/bcsr.h, SparseGEMM.h) are thin host wrappers over the
 * functions declared here.  Everything is plain C: pointers, sizes, opaque handles; no torch / C++ types.
 *
 * Conventions
 *   - functions returning int return 0 on success, a non-zero TSG_E* code on failure; tsg_last_error() (same
 *     string as sparse_last_error()) describes the failure on the calling thread.
 *   - "dev" pointers are CUDA device pointers on the current device.  Work is enqueued on the library's current
 *     stream (tsg_set_stream; default: the CUDA legacy default stream) and is asynchronous unless stated.
 *   - matrices are row-major fp32, indices int32, exactly as in the reference (sparse/tcsc.h:6-17).
 */
#ifndef TSGEMM_B200_H
#define TSGEMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    TSG_OK = 0,
    TSG_ECUDA = 1,    /* a CUDA runtime call failed */
    TSG_EINVAL = 2,   /* bad argument */
    TSG_ENOMEM = 3,   /* host or device allocation failed */
    TSG_ENODEV = 4,   /* no sm_100 device */
    TSG_ENCCL = 5,    /* a NCCL call failed */
    TSG_EUNSUPPORTED = 6
};

/* summation orders: which reference function's exact sequence of fp32 roundings to reproduce */
enum {
    TSG_ORDER_BIAS_FIRST = 0, /* tcsc_sgemm_basic            tcsc.c:69-98    y=B; +pos...; -neg...            */
    TSG_ORDER_BIAS_LAST = 1,  /* tcsc_sgemm_prelu_basic      tcsc.c:143-165  y=0; +pos...; -neg...; y+=B      */
                              /* sparseGEMM / sparseGEMM_PReLU SparseGEMM.h:104-119,151-168                  */
    TSG_ORDER_SPLIT = 2,      /* tcsc_sgemm_optimized, ..._prelu_optimized_{separate,onthego}
                                 tcsc.c:101-140,179-275      y=(B+fl(sum pos))-fl(sum neg)                    */
    TSG_ORDER_FAST = 3        /* opt-in, NOT bit-identical to any reference function: one sweep over K, per chunk of rows
                                 +pos... then -neg..., bias last.  Meets the tolerance contract (max |y-y64| /
                                 max(|y64|,1) <= 1e-5 or the reference's own error, whichever is larger); X streams
                                 through shared memory once instead of twice.  tsg_set_fast_order(1) makes the
                                 reference-named entry points use it.                                                   */
};

/* below this many rows of X the skinny (decode) kernel runs: lanes over a column's non-zeros + tree reduction */
#define TSG_SKINNY_M 32

/* ---- runtime ------------------------------------------------------------------------------------------------- */
const char *tsg_last_error(void);
void tsg_clear_error(void);                  /* the reference-named entry points clear it on entry */
const char *tsg_version(void);
int tsg_device_check(void);                 /* 0 iff the current device is sm_100 */
int tsg_set_stream(void *cuda_stream);      /* cudaStream_t as void*; NULL = legacy default stream */
void *tsg_get_stream(void);
int tsg_synchronize(void);                  /* cudaStreamSynchronize(current stream) */
/* counters of kernels launched by this library on this thread since the last reset (bench.py "gpu_launches") */
long long tsg_launch_count(void);
void tsg_launch_count_reset(void);
int tsg_dev_alloc(void **out, size_t bytes); /* stream-ordered pool allocation (cudaMallocAsync) */
int tsg_dev_free(void *p);

/* ---- TCSC device mirror --------------------------------------------------------------------------------------- */
typedef struct tsg_tcsc tsg_tcsc; /* opaque: device TCSC arrays + the private K-tiled gather stream */

/* dense (rows x cols, device) -> device TCSC.  Counting/compaction kernels, bit-exact w.r.t. tcsc.c:6-66 (f32:
 * ==1.0f / ==-1.0f) and SparseGEMM.h:20-39 (i32: >=1 / <=-1).  Synchronises once (to size the index arrays). */
int tsg_tcsc_from_dense_f32(const float *dense_dev, int rows, int cols, tsg_tcsc **out);
int tsg_tcsc_from_dense_i32(const int *dense_dev, int rows, int cols, tsg_tcsc **out);
/* adopt existing index arrays (host or device pointers; copied) */
int tsg_tcsc_from_arrays(const int *col_start_pos, const int *col_start_neg, const int *row_index_pos,
                         const int *row_index_neg, int rows, int cols, tsg_tcsc **out);
void tsg_tcsc_destroy(tsg_tcsc *W);
int tsg_tcsc_dims(const tsg_tcsc *W, int *rows, int *cols, int *n_pos, int *n_neg);
/* copy the four index arrays to host buffers of (cols+1, cols+1, n_pos, n_neg) ints; synchronous */
int tsg_tcsc_download(const tsg_tcsc *W, int *col_start_pos, int *col_start_neg, int *row_index_pos, int *row_index_neg);
/* device pointers of the four arrays (for device-side consumers) */
int tsg_tcsc_device_arrays(const tsg_tcsc *W, const int **csp, const int **csn, const int **rip, const int **rin);
/* bytes of the private gather stream and its tiling (kc rows of X per chunk); builds it if not built yet */
int tsg_tcsc_stream_info(tsg_tcsc *W, long long *bytes, int *kc, int *nchunk);

/* Y = [PReLU](X*W + B) on the device.  X: M x K, B: N, Y: M x N, all device, row-major; `order` is a
 * TSG_ORDER_* value; use_prelu != 0 applies y<0 ? a*y : y.  ldy = row pitch of Y in floats (>= N; lets the
 * column-partitioned path write its slab straight into the full Y). */
int tsg_tcsc_gemm(tsg_tcsc *W, const float *X_dev, const float *B_dev, float a, int use_prelu, int order,
                  float *Y_dev, int M, int N, int K, long long ldy);
/* opt-in for the reference-named entry points (tcsc_sgemm_*, sparseGEMM*): 1 = TSG_ORDER_FAST instead of the reference function's
 * exact order (per thread; default 0; also set by the environment variable TSG_FAST_ORDER=1 at first use) */
int tsg_set_fast_order(int on);
int tsg_get_fast_order(void);
/* force one kernel: 0 auto, 1 tiled shared-memory gather kernel, 2 skinny kernel */
int tsg_tcsc_set_kernel(int which);
int tsg_tcsc_get_kernel(void);
/* planning hooks of the tiled kernel -- pure host arithmetic, no device needed (tests/test_plan.py).
 * tsg_plan_units: out5 = {row tiles, 256-column tiles, full units, sub, total units}; tsg_plan_unit_at: out3 = {row tile,
 * first column, columns per warp} of unit u; tsg_plan_progress: the progress groups of the multi-GPU mode 2 and the
 * number of arrivals each group's counter will see. */
int tsg_plan_units(int M, int N, int sms, int out5[5]);
int tsg_plan_unit_at(int M, int N, int sms, int u, int out3[3]);
int tsg_plan_progress(int M, int N, int sms, int *ngroups, int gbound9[9], unsigned int target8[8]);
/* per-launch CUDA-event timing of the tiled GEMM kernel on its launching stream (bench.py's roofline numerator) */
int tsg_profile_enable(int on);
int tsg_profile_read(double *total_ms, int *launches);

/* ---- BCSR device mirror --------------------------------------------------------------------------------------- */
typedef struct tsg_bcsr tsg_bcsr;
int tsg_bcsr_from_dense_f32(const float *dense_dev, int rows, int cols, int r, int c, tsg_bcsr **out);
int tsg_bcsr_from_arrays(const int *b_row_start, const int *b_col_idx, const float *b_values, int r, int c, int br,
                         int bc, int k, tsg_bcsr **out);
void tsg_bcsr_destroy(tsg_bcsr *W);
int tsg_bcsr_dims(const tsg_bcsr *W, int *r, int *c, int *br, int *bc, int *k);
int tsg_bcsr_download(const tsg_bcsr *W, int *b_row_start, int *b_col_idx, float *b_values);
/* use_prelu: 0 = none, 1 = PReLU(X*W + B), 2 = the reference's literal bcsr_sgemm_prelu_* loop (activation after every
 * partial update, untouched outputs keep the raw bias: sparse/bcsr.c:177-218, 264-312), bit-identical to it. */
int tsg_bcsr_gemm(tsg_bcsr *W, const float *X_dev, const float *B_dev, float a, int use_prelu, float *Y_dev,
                  int M, int N, int K, long long ldy);
/* what the reference-named bcsr_sgemm_prelu_basic / _avx compute on this thread: 0 (default, also TSG_BCSR_PRELU_LITERAL
 * unset) = PReLU(X*W + B); 1 = the reference's literal loop, for callers that need its exact output. */
int tsg_bcsr_set_prelu_literal(int on);
int tsg_bcsr_get_prelu_literal(void);
/* kernel selection for A/B runs: 0 = default (the shared-memory ring kernel; the plain kernel when TSG_BCSR_RING=0 is
 * set), 1 = plain kernel, 2 = ring kernel (falls back to the plain one outside its limits: c not in {1,2,4,8,16},
 * r > 224).  Both produce the same bits. */
int tsg_bcsr_set_kernel(int which);

/* ---- device generators / verification (counter-based: element i is a pure function of (seed, i)) --------------- */
int tsg_gen_ternary_f32(float *W_dev, long long n, uint64_t seed, uint32_t num, uint32_t den);
int tsg_gen_ternary_i32(int *W_dev, long long n, uint64_t seed, uint32_t num, uint32_t den);
int tsg_gen_uniform_f32(float *X_dev, long long n, uint64_t seed);
int tsg_gen_intvalued_f32(float *X_dev, long long n, uint64_t seed, int range);
/* the two patterns of generateSparseMatrix (reference SparseGEMM.h:53-102) as counter-based generators, H x W row-major:
 * uniform != 0: every window of 2*nonZero columns of every row holds one +1 and one -1 on distinct even offsets
 * (nonZero >= 2); uniform == 0: row h holds (W/nonZero)/2 + d(h) entries +1 and (W/nonZero)/2 - d(h) entries -1 at
 * distinct random columns, d(h) uniform on [0, W/nonZero/20 + 1] (W <= 2^20). */
int tsg_gen_sparse_pattern_i32(int *W_dev, int H, int W, int nonZero, int uniform, uint64_t seed);
int tsg_gen_sparse_pattern_f32(float *W_dev, int H, int W, int nonZero, int uniform, uint64_t seed);
/* ternary generator for a column slice [col0, col0+ncols) of a K x N matrix (values identical to the full matrix) */
int tsg_gen_ternary_slice_f32(float *W_dev, int K, int N, int col0, int ncols, uint64_t seed, uint32_t num, uint32_t den);
/* fp64-accumulating dense check on the device: out[0] = max |Y - y64| / max(|y64|,1), out[1] = max |Y - y64|,
 * over rows [m0, m0+mrows) and all N columns of Y (pitch ldy); y64 = [PReLU](X*Wdense + B) in double.  Synchronous. */
int tsg_verify_dense_f64(const float *X_dev, const float *Wdense_dev, const float *B_dev, float a, int use_prelu,
                         const float *Y_dev, int M, int N, int K, long long ldy, int m0, int mrows, double *out2);

/* ---- column-partitioned multi-GPU path (one process per GPU) ---------------------------------------------------- */
typedef struct tsg_dist tsg_dist;
/* 128-byte NCCL unique id: rank 0 calls tsg_dist_unique_id, ships the bytes to the other ranks by any means
 * (torch.distributed / a file), every rank then calls tsg_dist_create.  Collective. */
int tsg_dist_unique_id(unsigned char id128[128]);
int tsg_dist_create(const unsigned char id128[128], int rank, int world, tsg_dist **out);
void tsg_dist_destroy(tsg_dist *D);
/* columns [*col0, *col0 + *ncols) of an N-column W owned by `rank` (contiguous, multiples of 32 where possible) */
void tsg_dist_partition(int N, int rank, int world, int *col0, int *ncols);
/* fused mode needs a symmetric Y: every rank allocates `bytes` (cudaMalloc) and maps every peer's buffer through
 * CUDA IPC (handles exchanged over NCCL).  Collective; returns this rank's buffer, owned by D. */
int tsg_dist_alloc_y(tsg_dist *D, size_t bytes, float **y_local);
/* the peer mappings of the symmetric Y, indexed by rank (entry `rank` is the local buffer); up to 8 entries */
int tsg_dist_peer_ptrs(tsg_dist *D, void *out[8]);
/* diagnostic (tools/peer_store_bw.py): write a rows x cols tile pattern into dst (row pitch ld floats; may be a peer
 * mapping) `iters` times; mode 0 = per-lane 64-byte row segments (the fused epilogue's pattern), mode 1 = 1 KB row
 * segments from shared memory with bulk async (TMA) stores */
int tsg_dbg_peer_store(float *dst, long long ld, int rows, int cols, int iters, int mode);
/* 1 if the symmetric Y of the last tsg_dist_alloc_y has an NVSwitch multicast mapping (mode 5 usable) */
int tsg_dist_has_multicast(tsg_dist *D);
/* stream-ordered cross-rank barrier (4-byte ncclAllReduce on the current stream) */
int tsg_dist_barrier(tsg_dist *D);
/* Y(M x N, full, on every rank) = [PReLU](X*W + B).  W_local holds this rank's column slice; B_dev is the full
 * bias (N); X_dev is valid on `root` and is broadcast in place (root < 0: every rank already holds X).
 * mode 0: the kernel writes a contiguous slab, ncclAllGather of the slabs, re-layout kernel into row-major Y.
 * mode 1: Y_dev must be the tsg_dist_alloc_y buffer; the kernel epilogue stores every finished row segment into
 *         the local Y and straight into every peer's Y over NVLink (fused all-gather).
 * mode 2: Y_dev must be the tsg_dist_alloc_y buffer; the kernel writes the local slab and bumps a progress counter
 *         per 128-row tile; per peer a copy stream waits on those counters (cuStreamWaitValue32, no SM involved) and
 *         pushes finished row blocks with strided 2-D DMA copies, so the all-gather overlaps the GEMM (M >= 32).
 * mode 3: Y_dev must be the tsg_dist_alloc_y buffer; fused all-gather through the TMA engine: every finished 128-row
 *         tile is staged in shared memory and its row segments are written to the local Y and to every peer's Y with
 *         bulk async stores while the SMs gather the next tile (M >= 32, N and the slab width multiples of 4).
 * mode 4: variant of mode 3: the staged tile gets shared memory of its own (shorter K chunks), so the stores of one unit
 *         drain while the next unit is gathered; same results (slower at N = 2: the shorter chunks cost more than the
 *         overlap returns).
 * mode 5: NVSwitch multicast: tsg_dist_alloc_y builds the symmetric Y from the virtual-memory API and binds every rank's
 *         buffer to one multicast object (tsg_dist_has_multicast() tells whether that worked); the kernel stages every
 *         finished tile in shared memory and writes it ONCE with multimem.st -- the switch replicates it into every
 *         rank's Y, this rank's included (M >= 32, N and the slab width multiples of 4). */
int tsg_dist_gemm(tsg_dist *D, tsg_tcsc *W_local, float *X_dev, int root, const float *B_dev, float a, int use_prelu,
                  int order, float *Y_dev, int M, int N, int K, int mode);

/* The same with HOST buffers that all ranks share (e.g. one POSIX shared-memory segment mapped by every process; pin it
 * with tsg_host_register in every process for full PCIe rate).  Every rank moves 1/world of the bytes over ITS OWN PCIe
 * link: rows [rank*ceil(M/world), ...) of X host->device, ncclAllGather of the row blocks over NVLink, tsg_dist_gemm into
 * the symmetric Y (tsg_dist_alloc_y must have been called with >= M*N*4 bytes; with world == 1 or mode 0 that buffer is
 * simply the device copy of Y), the same rows of the full Y device->host.  Synchronous: returns when this rank's rows of
 * Y_host are complete; Y_host is whole once every rank has returned.  Collective. */
int tsg_dist_gemm_host(tsg_dist *D, tsg_tcsc *W_local, const float *X_host, const float *B_dev, float a, int use_prelu,
                       int order, float *Y_host, int M, int N, int K, int mode);
/* cudaHostRegister / cudaHostUnregister for buffers the caller allocated itself (portable flag) */
int tsg_host_register(void *p, size_t bytes);
int tsg_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif
