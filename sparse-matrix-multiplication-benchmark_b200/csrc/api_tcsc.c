/*
 * api_tcsc.c -- plain-C host layer: the reference's TCSC entry points (include/sparse/tcsc.h; reference
 * sparse/tcsc.h:19-48) on top of the device-level C-ABI (include/tsgemm_b200.h).
 *
 * What happens here and nowhere else: argument conventions of the reference (host tcsc_t with malloc'ed arrays the
 * caller may read, tcsc.c:22-33), the side table that remembers the device mirror of every tcsc_t this library
 * handed out (or saw), and host<->device staging of X / B / Y.  No arithmetic happens on the CPU.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "sparse/tcsc.h"
#include "tsg_fingerprint.h"
#include "tsg_host_shim.h"
#include "tsgemm_b200.h"

/* ---- side table: tcsc_t* -> device mirror ------------------------------------------------------------------ */
typedef struct {
    const tcsc_t *host;
    tsg_tcsc *dev;
    /* fingerprint: a caller may free a tcsc_t and malloc may hand the same addresses out again, or edit the arrays in
     * place; dimensions, array pointers and a content hash (tsg_fingerprint.h) must all still match */
    int rows, cols, n_pos, n_neg;
    const int *csp, *csn, *rip, *rin;
    uint64_t fp;
} mirror_entry;

static mirror_entry *g_tab = NULL;
static size_t g_tab_len = 0, g_tab_cap = 0;
static pthread_mutex_t g_tab_mu = PTHREAD_MUTEX_INITIALIZER;

static uint64_t content_fp(const tcsc_t *W) {
    uint64_t h = 0x7463736300000001ull;
    const size_t nc = W->cols >= 0 ? (size_t)W->cols + 1 : 0;
    h = tsg_fp_words(W->col_start_pos, W->col_start_pos ? nc : 0, h);
    h = tsg_fp_words(W->col_start_neg, W->col_start_neg ? nc : 0, h);
    h = tsg_fp_words(W->row_index_pos, W->n_elem_pos > 0 ? (size_t)W->n_elem_pos : 0, h);
    h = tsg_fp_words(W->row_index_neg, W->n_elem_neg > 0 ? (size_t)W->n_elem_neg : 0, h);
    return h;
}

static int entry_matches(const mirror_entry *e, const tcsc_t *W, uint64_t fp) {
    return e->host == W && e->rows == W->rows && e->cols == W->cols && e->n_pos == W->n_elem_pos && e->n_neg == W->n_elem_neg &&
           e->csp == W->col_start_pos && e->csn == W->col_start_neg && e->rip == W->row_index_pos && e->rin == W->row_index_neg &&
           e->fp == fp;
}

/* an entry that shares the struct address or any array address with W describes memory that has been released and
 * handed out again (tcsc_free was bypassed, or the caller re-used a struct): it can never be valid once W exists */
static int entry_aliases(const mirror_entry *e, const tcsc_t *W) {
    return e->host == W || (W->col_start_pos && e->csp == W->col_start_pos) || (W->col_start_neg && e->csn == W->col_start_neg) ||
           (W->row_index_pos && e->rip == W->row_index_pos) || (W->row_index_neg && e->rin == W->row_index_neg);
}

#define TSG_MAX_STALE 8
/* with g_tab_mu held: drop every entry aliasing W (their mirrors go to stale[] for destruction after unlocking) */
static size_t purge_aliases_locked(const tcsc_t *W, tsg_tcsc **stale, size_t nstale) {
    for (size_t i = 0; i < g_tab_len;) {
        if (entry_aliases(&g_tab[i], W)) {
            if (nstale < TSG_MAX_STALE) stale[nstale++] = g_tab[i].dev; /* beyond that: leaked rather than freed under the lock */
            g_tab[i] = g_tab[--g_tab_len];
        } else ++i;
    }
    return nstale;
}

static int insert_locked(const tcsc_t *W, tsg_tcsc *dev, uint64_t fp) {
    if (g_tab_len == g_tab_cap) {
        size_t ncap = g_tab_cap ? 2 * g_tab_cap : 16;
        mirror_entry *nt = (mirror_entry *)realloc(g_tab, ncap * sizeof *nt);
        if (!nt) return TSG_ENOMEM;
        g_tab = nt;
        g_tab_cap = ncap;
    }
    mirror_entry e = {W, dev, W->rows, W->cols, W->n_elem_pos, W->n_elem_neg,
                      W->col_start_pos, W->col_start_neg, W->row_index_pos, W->row_index_neg, fp};
    g_tab[g_tab_len++] = e;
    return TSG_OK;
}

/* register W -> dev, replacing whatever the table believed about these addresses */
static int table_insert(const tcsc_t *W, tsg_tcsc *dev, uint64_t fp) {
    tsg_tcsc *stale[TSG_MAX_STALE];
    pthread_mutex_lock(&g_tab_mu);
    size_t ns = purge_aliases_locked(W, stale, 0);
    int rc = insert_locked(W, dev, fp);
    pthread_mutex_unlock(&g_tab_mu);
    for (size_t i = 0; i < ns; ++i) tsg_tcsc_destroy(stale[i]);
    return rc;
}

static tsg_tcsc *table_remove(const tcsc_t *W) {
    tsg_tcsc *dev = NULL;
    pthread_mutex_lock(&g_tab_mu);
    for (size_t i = g_tab_len; i-- > 0;)
        if (g_tab[i].host == W) {
            dev = g_tab[i].dev;
            g_tab[i] = g_tab[--g_tab_len];
            break;
        }
    pthread_mutex_unlock(&g_tab_mu);
    return dev;
}

/* mirror of W: cached, or (for a tcsc_t the caller assembled itself / a stale slot) built from W's host arrays.  The
 * hash and the H2D build run outside the table lock. */
static tsg_tcsc *mirror_of(const tcsc_t *W) {
    const uint64_t fp = content_fp(W);
    tsg_tcsc *dev = NULL;
    pthread_mutex_lock(&g_tab_mu);
    for (size_t i = g_tab_len; i-- > 0;) /* newest first */
        if (g_tab[i].host == W) {
            if (entry_matches(&g_tab[i], W, fp)) dev = g_tab[i].dev;
            break;
        }
    pthread_mutex_unlock(&g_tab_mu);
    if (dev) return dev;
    if (tsg_tcsc_from_arrays(W->col_start_pos, W->col_start_neg, W->row_index_pos, W->row_index_neg, W->rows, W->cols, &dev) != TSG_OK)
        return NULL;
    tsg_tcsc *stale[TSG_MAX_STALE], *winner = NULL;
    pthread_mutex_lock(&g_tab_mu);
    for (size_t i = g_tab_len; i-- > 0;) /* another thread may have mirrored the same W meanwhile */
        if (g_tab[i].host == W && entry_matches(&g_tab[i], W, fp)) { winner = g_tab[i].dev; break; }
    size_t ns = 0;
    if (winner) {
        stale[ns++] = dev;
        dev = winner;
    } else {
        ns = purge_aliases_locked(W, stale, 0);
        if (insert_locked(W, dev, fp) != TSG_OK) { stale[ns++] = dev; dev = NULL; }
    }
    pthread_mutex_unlock(&g_tab_mu);
    for (size_t i = 0; i < ns; ++i) tsg_tcsc_destroy(stale[i]);
    return dev;
}

/* explicit invalidation for callers that edit a tcsc_t's arrays in place between GEMM calls (extension) */
void tcsc_invalidate(const tcsc_t *W) {
    tsg_tcsc *dev = W ? table_remove(W) : NULL;
    if (dev) tsg_tcsc_destroy(dev);
}

/* ---- builder (reference sparse/tcsc.c:6-66) ----------------------------------------------------------------- */
tcsc_t *tcsc_from_dense(dense_t dense, int rows, int cols) {
    tsg_clear_error();
    void *ddev = NULL;
    int owned = 0;
    tsg_tcsc *dev = NULL;
    if (rows < 0 || cols < 0) return NULL;
    if (tsg_shim_stage_in(dense, (size_t)rows * (size_t)cols * sizeof(float), &ddev, &owned) != TSG_OK) return NULL;
    int rc = tsg_tcsc_from_dense_f32((const float *)ddev, rows, cols, &dev);
    tsg_shim_release(ddev, owned);
    if (rc != TSG_OK) return NULL;

    tcsc_t *W = (tcsc_t *)malloc(sizeof *W); /* tcsc.c:22 */
    if (!W) { tsg_tcsc_destroy(dev); return NULL; }
    W->rows = rows;
    W->cols = cols;
    tsg_tcsc_dims(dev, NULL, NULL, &W->n_elem_pos, &W->n_elem_neg);
    /* tcsc.c:30-33 (malloc(0) may return NULL: ask for at least one int so NULL always means failure) */
    W->col_start_pos = (int *)malloc(((size_t)cols + 1) * sizeof(int));
    W->col_start_neg = (int *)malloc(((size_t)cols + 1) * sizeof(int));
    W->row_index_pos = (int *)malloc(((size_t)W->n_elem_pos + 1) * sizeof(int));
    W->row_index_neg = (int *)malloc(((size_t)W->n_elem_neg + 1) * sizeof(int));
    int ok = W->col_start_pos && W->col_start_neg && W->row_index_pos && W->row_index_neg;
    if (ok) ok = tsg_tcsc_download(dev, W->col_start_pos, W->col_start_neg, W->row_index_pos, W->row_index_neg) == TSG_OK;
    if (ok) ok = table_insert(W, dev, content_fp(W)) == TSG_OK; /* also drops entries of freed structs whose addresses malloc re-used */
    if (!ok) { /* tcsc.c:35-43 */
        free(W->col_start_pos); free(W->col_start_neg); free(W->row_index_pos); free(W->row_index_neg);
        free(W);
        tsg_tcsc_destroy(dev);
        return NULL;
    }
    return W;
}

void tcsc_free(tcsc_t *W) { /* tcsc.c:167-175 */
    if (!W) return;
    tsg_tcsc *dev = table_remove(W);
    if (dev) tsg_tcsc_destroy(dev);
    free(W->col_start_pos);
    free(W->col_start_neg);
    free(W->row_index_pos);
    free(W->row_index_neg);
    free(W);
}

/* ---- GEMM entry points ------------------------------------------------------------------------------------------ */
static void run_gemm(const float *X, const tcsc_t *W, const float *B, float a, int use_prelu, int order, float *Y, int M, int N, int K) {
    tsg_clear_error();
    if (!W || M <= 0 || N <= 0) return;
    if (tsg_get_fast_order()) order = TSG_ORDER_FAST; /* opt-in: tolerance contract instead of the reference function's exact order */
    tsg_tcsc *dev = mirror_of(W);
    if (!dev) return; /* reason in sparse_last_error() */
    const int x_dev = tsg_shim_is_device(X), y_dev = tsg_shim_is_device(Y);
    if (!x_dev && !y_dev && (size_t)M * ((size_t)K + (size_t)N) * sizeof(float) >= ((size_t)8 << 20)) {
        /* the reference's calling convention, everything in host memory, and enough of it to be worth pipelining */
        tsg_shim_tcsc_gemm_hostpipe(dev, X, B, a, use_prelu, order, Y, M, N, K);
        return;
    }
    tsg_shim_tcsc_gemm_staged(dev, X, B, a, use_prelu, order, Y, M, N, K); /* failure: Y untouched, reason in sparse_last_error() */
}

void tcsc_sgemm_basic(const dense_t X, const tcsc_t *W, const dense_t B, dense_t Y, int M, int N, int K) {
    run_gemm(X, W, B, 0.0f, 0, TSG_ORDER_BIAS_FIRST, Y, M, N, K);
}
void tcsc_sgemm_optimized(const dense_t X, const tcsc_t *W, const dense_t B, dense_t Y, int M, int N, int K) {
    run_gemm(X, W, B, 0.0f, 0, TSG_ORDER_SPLIT, Y, M, N, K);
}
void tcsc_sgemm_prelu_basic(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y, int M, int N, int K) {
    run_gemm(X, W, B, a, 1, TSG_ORDER_BIAS_LAST, Y, M, N, K);
}
void tcsc_sgemm_prelu_optimized_separate(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y, int M, int N, int K) {
    run_gemm(X, W, B, a, 1, TSG_ORDER_SPLIT, Y, M, N, K);
}
void tcsc_sgemm_prelu_optimized_onthego(const dense_t X, const tcsc_t *W, const dense_t B, float a, dense_t Y, int M, int N, int K) {
    run_gemm(X, W, B, a, 1, TSG_ORDER_SPLIT, Y, M, N, K);
}
