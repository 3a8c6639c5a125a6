// gen_pattern.h -- pure per-element / per-row functions of the two generateSparseMatrix patterns (reference
// SparseGEMM.h:53-102) as counter-based generators.  Plain integer arithmetic, host and device: gen.cu builds its kernels
// from them, tests/test_patterns_gen.py compiles this header with g++ to run the same arithmetic on the CPU.
//
// The reference draws from rand() / mt19937(time(0)), so its matrices cannot be reproduced; what is kept is each
// pattern's structure:
//   uniformDistribution = true  (SparseGEMM.h:56-68): every row is cut into windows of 2*nonZero columns; each window gets
//     exactly one +1 and one -1, both on even offsets, at distinct positions.  (Where W is not a multiple of the window the
//     reference writes past the end of the row -- those cells are dropped here; nonZero = 1 never terminates there and is
//     refused here.)
//   uniformDistribution = false (SparseGEMM.h:70-99): row h gets half + d(h) entries +1 and half - d(h) entries -1 at
//     distinct uniformly random columns, half = (W/nonZero)/2, d(h) uniform on [0, W/nonZero/20 + 1] -- the +/- skew
//     varies from row to row (rows of W = the K index, so some k are "hot" in every column list).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define TSG_GP_HD __host__ __device__ __forceinline__
#else
#define TSG_GP_HD static inline
#endif

TSG_GP_HD uint64_t gp_hash64(uint64_t seed, uint64_t idx) {  // same finalizer as the other generators (gen.cu)
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform on [0, n) from the high 32 bits of a hash
TSG_GP_HD uint32_t gp_below(uint64_t h, uint32_t n) { return (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * n) >> 32); }

// ---- uniformDistribution = true: element (h, w) is a pure function of (seed, h, w) -----------------------------------------
TSG_GP_HD int gp_window_value(uint64_t seed, int h, int w, int W, int nonZero) {
    const int window = 2 * nonZero;
    const int g = w / window, off = w - g * window;
    if (off & 1) return 0;  // only even offsets are ever written (SparseGEMM.h:60-61: rand() % nonZero * 2)
    const int nwin = (W + window - 1) / window;
    const uint64_t cell = ((uint64_t)h * (uint64_t)nwin + (uint64_t)g) * 2u;
    const uint32_t plus = gp_below(gp_hash64(seed, cell), (uint32_t)nonZero);
    uint32_t minus = gp_below(gp_hash64(seed, cell + 1u), (uint32_t)nonZero - 1u);
    minus += (minus >= plus);  // uniform over the nonZero - 1 other slots (the reference redraws until distinct)
    const uint32_t slot = (uint32_t)off >> 1;
    return slot == plus ? 1 : (slot == minus ? -1 : 0);
}

// ---- uniformDistribution = false ---------------------------------------------------------------------------------------------
TSG_GP_HD void gp_row_limits(uint64_t seed, int h, int W, int nonZero, int *n_plus, int *n_minus) {
    const int per_row = W / nonZero, half = per_row / 2, dmax = per_row / 20 + 1;  // SparseGEMM.h:73,75-77
    const int d = (int)gp_below(gp_hash64(seed + 0x632BE59BD9B4E019ull, (uint64_t)h), (uint32_t)dmax + 1u);
    int p = half + d, n = half - d;
    if (n < 0) n = 0;       // the reference's `while (count < limitNeg)` simply does not run
    if (p > W) p = W;       // (the reference would never terminate)
    if (n > W - p) n = W - p;
    *n_plus = p;
    *n_minus = n;
}
// selection key of column w in row h: distinct inside a row (low 20 bits = w, so W <= 2^20); the n_plus smallest keys of a
// row become +1, the next n_minus become -1 -- a uniformly random choice of distinct columns, like the rejection loop
TSG_GP_HD uint64_t gp_key(uint64_t seed, int h, int w, int W) {
    return (gp_hash64(seed, (uint64_t)h * (uint64_t)W + (uint64_t)w) & ~0xFFFFFull) | (uint64_t)w;
}
TSG_GP_HD int gp_skew_value(uint64_t key, int n_plus, int n_minus, uint64_t t_plus, uint64_t t_all) {
    if (n_plus > 0 && key <= t_plus) return 1;
    if (n_plus + n_minus > 0 && key <= t_all) return -1;
    return 0;
}
// L-th smallest key of row h (L >= 1), serial form (the kernel runs the same bisection with the count spread over a warp)
TSG_GP_HD uint64_t gp_kth_key_serial(uint64_t seed, int h, int W, int L) {
    uint64_t lo = 0, hi = ~0ull;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int w = 0; w < W; ++w) c += (gp_key(seed, h, w, W) <= mid);
        if (c >= L) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
