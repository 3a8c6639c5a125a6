/* tsg_fingerprint.h -- content fingerprints for the host-side mirror tables (api_tcsc.c, api_bcsr.c, api_cxx.cpp).
 *
 * The reference hands callers plain structs with malloc'ed arrays (sparse/tcsc.h:6-17, sparse/bcsr.h:5-12) that they
 * may free(), rebuild at the same address or edit in place, and nothing tells the library.  A device mirror is
 * therefore only reused when the struct's dimensions, its array pointers AND a hash of the array contents still match.
 * Arrays of up to TSG_FP_FULL_WORDS 32-bit words are hashed completely; longer ones by their first and last 64 words
 * plus 128 evenly spaced words (about a microsecond per call for the four TCSC arrays: the reference's own shapes are a
 * 15-microsecond kernel, so the check has to stay well below that); TSG_MIRROR_CHECK=full hashes everything,
 * TSG_MIRROR_CHECK=off trusts pointers and sizes alone.  Private. */
#ifndef TSG_FINGERPRINT_H
#define TSG_FINGERPRINT_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TSG_FP_FULL_WORDS 256

static inline uint64_t tsg_fp_mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    return h ^ (h >> 31);
}

/* 0 = sampled (default), 1 = full, 2 = off */
static inline int tsg_fp_mode(void) {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("TSG_MIRROR_CHECK");
        mode = (e && !strcmp(e, "full")) ? 1 : (e && !strcmp(e, "off")) ? 2 : 0;
    }
    return mode;
}

static inline uint64_t tsg_fp_words(const void *p, size_t nwords, uint64_t h) {
    const uint32_t *w = (const uint32_t *)p;
    h = tsg_fp_mix(h, (uint64_t)nwords);
    if (!w || nwords == 0) return h;
    const int mode = tsg_fp_mode();
    if (mode == 2) return h;
    if (mode == 1 || nwords <= TSG_FP_FULL_WORDS) {
        for (size_t i = 0; i < nwords; ++i) h = tsg_fp_mix(h, w[i]);
        return h;
    }
    for (size_t i = 0; i < 64; ++i) h = tsg_fp_mix(h, w[i]);
    for (size_t i = nwords - 64; i < nwords; ++i) h = tsg_fp_mix(h, w[i]);
    const size_t step = (nwords - 128) / 128 + 1;
    for (size_t i = 64; i < nwords - 64; i += step) h = tsg_fp_mix(h, w[i]);
    return h;
}

#endif
