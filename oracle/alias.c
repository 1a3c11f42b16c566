/*
 * oracle/alias.c -- TEST INFRASTRUCTURE: Vose alias table over deg^alpha.
 *
 * `use_scale_free_distribution` ("Sample negatives proportionally to their
 * degree", /root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:101-102);
 * north_star fixes the exponent at 0.75 ("a unigram^0.75 alias table"),
 * alpha = 1 approximates Ensmallen's edge-endpoint sampling.
 *
 * Normative construction (the product's host-side builder must reproduce it
 * bit for bit):  w_i = deg_i^alpha in double (exact forms for alpha in
 * {0, 0.5, 0.75, 1}, pow() otherwise);  total = left-to-right double sum;
 * scaled_i = w_i * n / total;  small/large are LIFO stacks filled in ascending
 * node order;  thr_i = min(floor(prob_i * 2^32), 2^32 - 1).
 * Sampling:  idx = mulhi(r_a, n);  node = r_b < thr[idx] ? idx : alias[idx].
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

static double degree_weight(uint64_t deg, double alpha) {
    const double d = (double)deg;
    if (deg == 0) return 0.0;
    if (alpha == 0.0) return 1.0;
    if (alpha == 1.0) return d;
    if (alpha == 0.5) return sqrt(d);
    if (alpha == 0.75) return sqrt(sqrt(d * d * d));
    return pow(d, alpha);
}

int orc_alias_build(const int64_t *indptr, uint64_t n, double alpha, uint32_t *thr,
                    uint32_t *alias) {
    if (!indptr || !thr || !alias || n == 0 || n > 0xFFFFFFFFull) return -1;
    double *scaled = (double *)malloc(n * sizeof(double));
    uint32_t *small = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *large = (uint32_t *)malloc(n * sizeof(uint32_t));
    if (!scaled || !small || !large) { free(scaled); free(small); free(large); return -2; }
    double total = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        scaled[i] = degree_weight((uint64_t)(indptr[i + 1] - indptr[i]), alpha);
        total += scaled[i];
    }
    if (!(total > 0.0)) { free(scaled); free(small); free(large); return -3; }
    uint64_t ns = 0, nl = 0;
    for (uint64_t i = 0; i < n; ++i) {
        scaled[i] = scaled[i] * (double)n / total;
        if (scaled[i] < 1.0) small[ns++] = (uint32_t)i; else large[nl++] = (uint32_t)i;
    }
    for (uint64_t i = 0; i < n; ++i) { thr[i] = 0xFFFFFFFFu; alias[i] = (uint32_t)i; }
    while (ns > 0 && nl > 0) {
        const uint32_t s = small[--ns];
        const uint32_t l = large[--nl];
        const double t = floor(scaled[s] * 4294967296.0);
        thr[s] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        alias[s] = l;
        scaled[l] = (scaled[l] + scaled[s]) - 1.0;
        if (scaled[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    /* leftovers keep prob 1 / self alias (set above) */
    free(scaled); free(small); free(large);
    return 0;
}
