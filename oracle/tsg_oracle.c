/*
 * tsg_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see tsg_oracle.h for the usage rule and parity status).
 *
 * Each function restates, in plain C, the arithmetic of one reference function and cites it.  The summation
 * ORDER is the contract (fp32 addition is not associative), so every loop nest below produces the same sequence
 * of roundings per output element as the cited reference code, even where the loop nest itself is organised
 * differently.  Build with -ffp-contract=off -fno-fast-math (oracle/Makefile) so the compiler keeps that order.
 */
#include "tsg_oracle.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ============================================================================================================
 * builders
 * ========================================================================================================== */

#define IS_POS_F32(v) ((v) == 1.0f)
#define IS_NEG_F32(v) ((v) == -1.0f)
#define IS_POS_I32(v) ((v) >= 1)
#define IS_NEG_I32(v) ((v) <= -1)

/* sparse/tcsc.c:11-19 -- a value is +1 iff it compares equal to 1.0f, -1 iff equal to -1.0f; anything else
 * (0.5, 2, NaN, -0.0) is dropped. */
void orc_tcsc_count_f32(const float *dense, int rows, int cols, int *n_pos, int *n_neg) {
    int64_t total = (int64_t)rows * cols;
    int p = 0, q = 0;
    for (int64_t e = 0; e < total; ++e) {
        float v = dense[e];
        p += IS_POS_F32(v);
        q += (!IS_POS_F32(v)) && IS_NEG_F32(v);
    }
    *n_pos = p;
    *n_neg = q;
}

/* sparse/tcsc.c:45-63 -- column j's lists start at the running counters; rows are appended in ascending order;
 * the sentinel entry [cols] holds the totals. */
void orc_tcsc_fill_f32(const float *dense, int rows, int cols,
                       int *col_start_pos, int *col_start_neg, int *row_index_pos, int *row_index_neg) {
    int p = 0, q = 0;
    for (int j = 0; j < cols; ++j) {
        col_start_pos[j] = p;
        col_start_neg[j] = q;
        const float *colp = dense + j;
        for (int i = 0; i < rows; ++i, colp += cols) {
            float v = *colp;
            if (IS_POS_F32(v)) row_index_pos[p++] = i;
            else if (IS_NEG_F32(v)) row_index_neg[q++] = i;
        }
    }
    col_start_pos[cols] = p;
    col_start_neg[cols] = q;
}

/* SparseGEMM.h:20-39 -- same structure on an int matrix with >=1 / <=-1 (so +-2 count as non-zeros). */
void orc_tcsc_count_i32(const int *dense, int rows, int cols, int *n_pos, int *n_neg) {
    int64_t total = (int64_t)rows * cols;
    int p = 0, q = 0;
    for (int64_t e = 0; e < total; ++e) {
        int v = dense[e];
        p += IS_POS_I32(v);
        q += IS_NEG_I32(v);
    }
    *n_pos = p;
    *n_neg = q;
}

void orc_tcsc_fill_i32(const int *dense, int rows, int cols,
                       int *col_start_pos, int *col_start_neg, int *row_index_pos, int *row_index_neg) {
    int p = 0, q = 0;
    for (int j = 0; j < cols; ++j) {
        col_start_pos[j] = p;
        col_start_neg[j] = q;
        const int *colp = dense + j;
        for (int i = 0; i < rows; ++i, colp += cols) {
            int v = *colp;
            if (IS_POS_I32(v)) row_index_pos[p++] = i;
            else if (IS_NEG_I32(v)) row_index_neg[q++] = i;
        }
    }
    col_start_pos[cols] = p;
    col_start_neg[cols] = q;
}

/* does block (brow,bcol) hold at least one +-1 ?  sparse/bcsr.c:52-64 (compares against the doubles -1.0/1.0,
 * which for a float operand is the same test as ==-1.0f/==1.0f) */
static int bcsr_block_kept(const float *dense, int cols, int r, int c, int brow, int bcol) {
    for (int i = 0; i < r; ++i) {
        const float *row = dense + (int64_t)(brow * r + i) * cols + (int64_t)bcol * c;
        for (int j = 0; j < c; ++j)
            if (row[j] == 1.0f || row[j] == -1.0f) return 1;
    }
    return 0;
}

int orc_bcsr_count(const float *dense, int rows, int cols, int r, int c) {
    int br = rows / r, bc = cols / c, k = 0; /* sparse/bcsr.c:24-25: integer division, remainder dropped */
    for (int brow = 0; brow < br; ++brow)
        for (int bcol = 0; bcol < bc; ++bcol)
            k += bcsr_block_kept(dense, cols, r, c, brow, bcol);
    return k;
}

void orc_bcsr_fill(const float *dense, int rows, int cols, int r, int c, int quirk, int tail_fill,
                   int *b_row_start, int *b_col_idx, float *b_values) {
    int br = rows / r, bc = cols / c;
    int k = 0;        /* blocks are numbered in row-major block order, sparse/bcsr.c:66-69 */
    int written = 0;  /* entries of b_row_start written so far (quirk mode) */
    for (int brow = 0; brow < br; ++brow) {
        int first_of_row = k;
        for (int bcol = 0; bcol < bc; ++bcol) {
            if (!bcsr_block_kept(dense, cols, r, c, brow, bcol)) continue;
            b_col_idx[k] = bcol; /* :119 */
            for (int i = 0; i < r; ++i) /* :122-134 whole block, zeros and non-ternary values included */
                for (int j = 0; j < c; ++j)
                    b_values[(int64_t)k * r * c + i * c + j] =
                        dense[(int64_t)(brow * r + i) * cols + (int64_t)bcol * c + j];
            ++k;
        }
        if (!quirk) b_row_start[brow] = first_of_row;
        else if (k != first_of_row) b_row_start[written++] = first_of_row; /* :114-117 only non-empty block-rows */
    }
    if (!quirk) {
        b_row_start[br] = k;
    } else {
        b_row_start[written++] = k; /* :137 */
        while (written < br + 1) b_row_start[written++] = tail_fill; /* undefined in the reference */
    }
}

/* ============================================================================================================
 * TCSC kernels
 * ========================================================================================================== */

/* sum of X[m, idx[t]] for t in [lo,hi), added to / subtracted from `y` one term at a time in ascending t */
static inline float gather_add(float y, const float *xrow, const int *idx, int lo, int hi) {
    for (int t = lo; t < hi; ++t) y += xrow[idx[t]];
    return y;
}
static inline float gather_sub(float y, const float *xrow, const int *idx, int lo, int hi) {
    for (int t = lo; t < hi; ++t) y -= xrow[idx[t]];
    return y;
}
static inline float prelu_lt0(float y, float a) { return (y < 0.0f) ? a * y : y; } /* tcsc.c:162 */

/* sparse/tcsc.c:69-98 : y starts at the bias, then +pos (ascending), then -neg (ascending) */
void orc_tcsc_sgemm_basic(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                          const float *B, float *Y, int M, int N, int K) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        float *yrow = Y + (int64_t)m * N;
        for (int n = 0; n < N; ++n) {
            float y = B[n];
            y = gather_add(y, xrow, rip, csp[n], csp[n + 1]);
            y = gather_sub(y, xrow, rin, csn[n], csn[n + 1]);
            yrow[n] = y;
        }
    }
}

/* shared body of tcsc_sgemm_optimized (tcsc.c:101-140), ..._prelu_optimized_separate (:179-227) and
 * ..._onthego (:231-275): Y = (B + fl(sum pos from 0)) - fl(sum neg from 0), the two partial sums rounded
 * separately.  The per-element result does not depend on the n-outer loop order of the reference, so the
 * restatement walks m-outer (cache friendly) -- same roundings, same values. */
static void tcsc_optimized_core(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                                const float *B, float *Y, int M, int N, int K, int fuse_prelu, float a) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        float *yrow = Y + (int64_t)m * N;
        for (int n = 0; n < N; ++n) {
            float acc_pos = gather_add(0.0f, xrow, rip, csp[n], csp[n + 1]);
            float acc_neg = gather_add(0.0f, xrow, rin, csn[n], csn[n + 1]);
            float y = B[n];
            y += acc_pos;
            y -= acc_neg;
            yrow[n] = fuse_prelu ? prelu_lt0(y, a) : y;
        }
    }
}

void orc_tcsc_sgemm_optimized(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                              const float *B, float *Y, int M, int N, int K) {
    tcsc_optimized_core(X, csp, csn, rip, rin, B, Y, M, N, K, 0, 0.0f);
}

/* sparse/tcsc.c:143-165 : y = 0, +pos, -neg, +bias, PReLU */
void orc_tcsc_sgemm_prelu_basic(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                                const float *B, float a, float *Y, int M, int N, int K) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        float *yrow = Y + (int64_t)m * N;
        for (int n = 0; n < N; ++n) {
            float y = 0.0f;
            y = gather_add(y, xrow, rip, csp[n], csp[n + 1]);
            y = gather_sub(y, xrow, rin, csn[n], csn[n + 1]);
            y += B[n];
            yrow[n] = prelu_lt0(y, a);
        }
    }
}

/* tcsc.c:179-227 : optimized matmul, then a separate PReLU sweep over Y (:221-226) */
void orc_tcsc_sgemm_prelu_separate(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                                   const float *B, float a, float *Y, int M, int N, int K) {
    tcsc_optimized_core(X, csp, csn, rip, rin, B, Y, M, N, K, 0, 0.0f);
    int64_t total = (int64_t)M * N;
    for (int64_t e = 0; e < total; ++e) Y[e] = prelu_lt0(Y[e], a);
}

/* tcsc.c:231-275 : PReLU applied right after the `-= acc_neg` (:268-272) */
void orc_tcsc_sgemm_prelu_onthego(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                                  const float *B, float a, float *Y, int M, int N, int K) {
    tcsc_optimized_core(X, csp, csn, rip, rin, B, Y, M, N, K, 1, a);
}

/* SparseGEMM.h:104-119 : y=0, +pos, -neg, store y+b[n] */
void orc_sparse_gemm_f32(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                         const float *b, float *Y, int M, int N, int K) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        for (int n = 0; n < N; ++n) {
            float y = 0.0f;
            y = gather_add(y, xrow, rip, csp[n], csp[n + 1]);
            y = gather_sub(y, xrow, rin, csn[n], csn[n + 1]);
            Y[(int64_t)m * N + n] = y + b[n];
        }
    }
}

/* SparseGEMM.h:151-168 : the same followed by (y<0)? a*y : y -- arithmetically identical to tcsc.c:143-165 */
void orc_sparse_gemm_prelu_f32(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                               const float *b, float *Y, int M, int N, int K, float a) {
    orc_tcsc_sgemm_prelu_basic(X, csp, csn, rip, rin, b, a, Y, M, N, K);
}

/* tolerance anchor (BASELINE.json north_star: "max relative error <= 1e-5 versus a double-precision
 * accumulation") */
void orc_tcsc_sgemm_f64(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                        const float *B, int use_prelu, double a, double *Y, int M, int N, int K) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        for (int n = 0; n < N; ++n) {
            double y = 0.0;
            for (int t = csp[n]; t < csp[n + 1]; ++t) y += (double)xrow[rip[t]];
            for (int t = csn[n]; t < csn[n + 1]; ++t) y -= (double)xrow[rin[t]];
            y += (double)B[n];
            if (use_prelu && y < 0.0) y *= a;
            Y[(int64_t)m * N + n] = y;
        }
    }
}

void orc_tcsc_abs_mass(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                       const float *B, double *S, int M, int N, int K) {
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        for (int n = 0; n < N; ++n) {
            double s = fabs((double)B[n]);
            for (int t = csp[n]; t < csp[n + 1]; ++t) s += fabs((double)xrow[rip[t]]);
            for (int t = csn[n]; t < csn[n + 1]; ++t) s += fabs((double)xrow[rin[t]]);
            S[(int64_t)m * N + n] = s;
        }
    }
}

/* ============================================================================================================
 * BCSR kernels
 * ========================================================================================================== */

static void fill_bias(float *Y, const float *B, int M, int N) {
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) Y[(int64_t)m * N + n] = B[n];
}

/* sparse/bcsr.c:141-175 : Y = bias, then for every stored block, Y[m, bc*c+j] += X[m, br*r+i] * val, walking
 * block-rows ascending => per output element the partial sums arrive in ascending k. */
void orc_bcsr_sgemm_basic(const float *X, int r, int c, int br, const int *b_row_start, const int *b_col_idx,
                          const float *b_values, const float *B, float *Y, int M, int N, int K) {
    fill_bias(Y, B, M, N);
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        float *yrow = Y + (int64_t)m * N;
        for (int brow = 0; brow < br; ++brow) {
            for (int bi = b_row_start[brow]; bi < b_row_start[brow + 1]; ++bi) {
                const float *blk = b_values + (int64_t)bi * r * c;
                float *yseg = yrow + (int64_t)b_col_idx[bi] * c;
                for (int i = 0; i < r; ++i) {
                    float x = xrow[brow * r + i];
                    for (int j = 0; j < c; ++j) {
                        float prod = x * blk[i * c + j];
                        yseg[j] += prod;
                    }
                }
            }
        }
    }
}

/* sparse/bcsr.c:177-218, LITERALLY: the activation `result>0 ? result : a*result` runs after every single
 * partial update (:208-212), and outputs never touched by a block keep the raw bias.  This is not
 * PReLU(X*W+b); it is restated only so the report can show what the reference actually returns. */
void orc_bcsr_sgemm_prelu_literal(const float *X, int r, int c, int br, const int *b_row_start, const int *b_col_idx,
                                  const float *b_values, const float *B, float a, float *Y, int M, int N, int K) {
    fill_bias(Y, B, M, N);
    for (int m = 0; m < M; ++m) {
        const float *xrow = X + (int64_t)m * K;
        float *yrow = Y + (int64_t)m * N;
        for (int brow = 0; brow < br; ++brow) {
            for (int bi = b_row_start[brow]; bi < b_row_start[brow + 1]; ++bi) {
                const float *blk = b_values + (int64_t)bi * r * c;
                float *yseg = yrow + (int64_t)b_col_idx[bi] * c;
                for (int i = 0; i < r; ++i) {
                    float x = xrow[brow * r + i];
                    for (int j = 0; j < c; ++j) {
                        float prod = x * blk[i * c + j];
                        float res = yseg[j] + prod;
                        yseg[j] = (res > 0.0f) ? res : a * res;
                    }
                }
            }
        }
    }
}

/* what BASELINE.json's north_star asks the product to compute: PReLU(X*W + b) */
void orc_bcsr_sgemm_prelu_math(const float *X, int r, int c, int br, const int *b_row_start, const int *b_col_idx,
                               const float *b_values, const float *B, float a, float *Y, int M, int N, int K) {
    orc_bcsr_sgemm_basic(X, r, c, br, b_row_start, b_col_idx, b_values, B, Y, M, N, K);
    int64_t total = (int64_t)M * N;
    for (int64_t e = 0; e < total; ++e) Y[e] = prelu_lt0(Y[e], a);
}

/* ============================================================================================================
 * dense helpers
 * ========================================================================================================== */

/* dense/dense.c:64-77 : y = 0; y += X*W over ascending k; store y + B[n] */
void orc_gemm_basic(const float *X, const float *W, const float *B, float *Y, int M, int N, int K) {
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float y = 0.0f;
            for (int k = 0; k < K; ++k) {
                float prod = X[(int64_t)m * K + k] * W[(int64_t)k * N + n];
                y += prod;
            }
            Y[(int64_t)m * N + n] = y + B[n];
        }
}

/* dense/dense.c:42-59 : absolute tolerance (1e-4 in the reference), first failure => false */
int orc_compare(const float *result, const float *target, int rows, int cols, float tol) {
    int64_t total = (int64_t)rows * cols;
    for (int64_t e = 0; e < total; ++e)
        if (fabs(result[e] - target[e]) > tol) return 0;
    return 1;
}

/* ============================================================================================================
 * generators
 * ========================================================================================================== */

uint64_t orc_hash64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* r = floor(u32 * den / 2^32) is uniform on [0,den); r<num => non-zero, sign from bit 0 of the hash */
static inline int ternary_draw(uint64_t h, uint32_t num, uint32_t den) {
    uint32_t u = (uint32_t)(h >> 32);
    uint32_t r = (uint32_t)(((uint64_t)u * den) >> 32);
    if (r >= num) return 0;
    return (h & 1ull) ? -1 : 1;
}

void orc_gen_ternary_f32(float *W, int64_t n, uint64_t seed, uint32_t num, uint32_t den) {
    for (int64_t i = 0; i < n; ++i) W[i] = (float)ternary_draw(orc_hash64(seed, (uint64_t)i), num, den);
}

void orc_gen_ternary_i32(int *W, int64_t n, uint64_t seed, uint32_t num, uint32_t den) {
    for (int64_t i = 0; i < n; ++i) W[i] = ternary_draw(orc_hash64(seed, (uint64_t)i), num, den);
}

void orc_gen_uniform_f32(float *X, int64_t n, uint64_t seed) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t m24 = (uint32_t)(orc_hash64(seed, (uint64_t)i) >> 40); /* 24 random bits */
        X[i] = (float)m24 * (1.0f / 8388608.0f) - 1.0f;                 /* exact: (m24 - 2^23) * 2^-23 */
    }
}

/* generateSparseMatrix patterns (SparseGEMM.h:53-102).  Written independently of the product's gen_pattern.h: the skewed
 * pattern sorts the row's selection keys instead of bisecting for thresholds. */
static uint32_t below(uint64_t h, uint32_t n) { return (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * n) >> 32); }
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
int orc_gen_sparse_pattern_i32(int *Wm, int H, int W, int nonZero, int uniform, uint64_t seed) {
    if (!Wm || H < 0 || W < 0 || nonZero < 1 || (uniform && nonZero < 2) || (!uniform && W > (1 << 20))) return -1;
    memset(Wm, 0, (size_t)H * (size_t)W * sizeof(int));
    if (uniform) { /* SparseGEMM.h:56-68 */
        int window = 2 * nonZero, nwin = (W + window - 1) / window;
        for (int h = 0; h < H; ++h)
            for (int g = 0; g < nwin; ++g) {
                uint64_t cell = ((uint64_t)h * (uint64_t)nwin + (uint64_t)g) * 2u;
                uint32_t plus = below(orc_hash64(seed, cell), (uint32_t)nonZero);
                uint32_t minus = below(orc_hash64(seed, cell + 1u), (uint32_t)nonZero - 1u);
                if (minus >= plus) ++minus; /* the reference redraws until the two slots differ (:62-64) */
                int wp = g * window + 2 * (int)plus, wm = g * window + 2 * (int)minus;
                if (wp < W) Wm[(size_t)h * W + wp] = 1; /* the reference writes past the row end here (:61,65) */
                if (wm < W) Wm[(size_t)h * W + wm] = -1;
            }
        return 0;
    }
    uint64_t *keys = (uint64_t *)malloc((size_t)(W > 0 ? W : 1) * sizeof(uint64_t));
    if (!keys) return -1;
    for (int h = 0; h < H; ++h) { /* SparseGEMM.h:74-98 */
        int per_row = W / nonZero, half = per_row / 2, dmax = per_row / 20 + 1;
        int d = (int)below(orc_hash64(seed + 0x632BE59BD9B4E019ull, (uint64_t)h), (uint32_t)dmax + 1u);
        int n_plus = half + d, n_minus = half - d;
        if (n_minus < 0) n_minus = 0;
        if (n_plus > W) n_plus = W;
        if (n_minus > W - n_plus) n_minus = W - n_plus;
        for (int w = 0; w < W; ++w) keys[w] = (orc_hash64(seed, (uint64_t)h * (uint64_t)W + (uint64_t)w) & ~0xFFFFFull) | (uint64_t)w;
        qsort(keys, (size_t)W, sizeof(uint64_t), cmp_u64);
        for (int i = 0; i < n_plus + n_minus; ++i) Wm[(size_t)h * W + (int)(keys[i] & 0xFFFFFull)] = (i < n_plus) ? 1 : -1;
    }
    free(keys);
    return 0;
}

void orc_gen_intvalued_f32(float *X, int64_t n, uint64_t seed, int range) {
    uint32_t span = 2u * (uint32_t)range + 1u;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)(orc_hash64(seed, (uint64_t)i) >> 32);
        int v = (int)(((uint64_t)u * span) >> 32) - range;
        X[i] = (float)v;
    }
}

/* ============================================================================================================
 * timing helper
 * ========================================================================================================== */

double orc_time_tcsc_sgemm_prelu_basic(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                                       const float *B, float a, float *Y, int M, int N, int K, int reps) {
    double best = 1e300;
    for (int rep = 0; rep < reps; ++rep) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        orc_tcsc_sgemm_prelu_basic(X, csp, csn, rip, rin, B, a, Y, M, N, K);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (s < best) best = s;
    }
    return best;
}
