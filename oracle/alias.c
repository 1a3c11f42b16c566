/*
 * oracle/alias.c -- TEST INFRASTRUCTURE: alias table over deg^alpha for the negative draws.
 *
 * `use_scale_free_distribution` ("Sample negatives proportionally to their
 * degree", /root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:101-102);
 * north_star fixes the exponent at 0.75 ("a unigram^0.75 alias table"),
 * alpha = 1 approximates Ensmallen's edge-endpoint sampling.
 *
 * Normative construction (round 2: stated in INTEGERS, so that the product can build the very same
 * table with parallel prefix sums on the GPU -- a left-to-right double sum cannot be split):
 *   w_i  = deg_i^alpha in double (exact forms for alpha in {0, 0.5, 0.75, 1}, pow() otherwise);
 *   W_i  = floor(w_i 2^F), F = the largest value <= 40 with n (floor(w_max 2^F) + 1) < 2^62;
 *   T    = sum W_i;  c = ceil(T / n) is the capacity of a bucket;  the n c - T units that are
 *          missing to fill n buckets go to the nodes of degree > 0, q = excess / n_src each and
 *          one more to the first excess % n_src of them in node order (a relative change of
 *          at most (q + 1) 2^-F of a weight);  m_i = the resulting mass, sum m_i = n c exactly.
 *   light nodes (m_i < c) and heavy nodes (m_i >= c) are swept in node order (the FIFO form of
 *   Vose's method): a light node keeps its own mass in its bucket and is topped up by the
 *   current heavy node; a heavy node whose remainder drops below c becomes the owner of its own
 *   bucket with that remainder and is topped up by the next heavy node; what is never exhausted
 *   owns a full bucket.
 *   thr = floor(own mass 2^32 / c) by 32 steps of long division; a full bucket has
 *   thr = 2^32 - 1 and aliases itself.
 * Sampling:  idx = mulhi(r_a, n);  node = r_b < thr[idx] ? idx : alias[idx].
 * The GPU builder (csrc/alias_build.cu) computes the same table in closed form from the two
 * prefix sums (deficits of the light nodes, surpluses of the heavy ones); tests compare them
 * bit for bit, which is also what checks the closed form against this sweep.
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

/* fractional bits of the fixed-point weights */
int orc_alias_fraction_bits(uint64_t n, uint64_t max_degree, double alpha) {
    const double w_max = degree_weight(max_degree, alpha);
    int bits = 40;
    while (bits > 0 && (double)n * (floor(ldexp(w_max, bits)) + 1.0) >= 4611686018427387904.0) --bits;
    return bits;
}

/* floor(mass 2^32 / capacity) for mass < capacity < 2^62, by long division */
static uint32_t threshold(uint64_t mass, uint64_t capacity) {
    uint64_t r = mass;
    uint32_t q = 0;
    for (int bit = 0; bit < 32; ++bit) {
        r <<= 1;
        q <<= 1;
        if (r >= capacity) { r -= capacity; q |= 1u; }
    }
    return q;
}

int orc_alias_build(const int64_t *indptr, uint64_t n, double alpha, uint32_t *thr,
                    uint32_t *alias) {
    if (!indptr || !thr || !alias || n == 0 || n > 0xFFFFFFFFull) return -1;
    uint64_t *mass = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint32_t *light = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *heavy = (uint32_t *)malloc(n * sizeof(uint32_t));
    if (!mass || !light || !heavy) { free(mass); free(light); free(heavy); return -2; }
    uint64_t max_degree = 0, n_src = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t deg = (uint64_t)(indptr[i + 1] - indptr[i]);
        if (deg > max_degree) max_degree = deg;
        n_src += deg > 0;
    }
    const int bits = orc_alias_fraction_bits(n, max_degree, alpha);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        mass[i] = (uint64_t)floor(ldexp(degree_weight((uint64_t)(indptr[i + 1] - indptr[i]), alpha), bits));
        total += mass[i];
    }
    if (total == 0 || n_src == 0) { free(mass); free(light); free(heavy); return -3; }
    const uint64_t capacity = (total + n - 1) / n;
    const uint64_t excess = capacity * n - total, each = excess / n_src, first = excess % n_src;
    uint64_t rank = 0, n_light = 0, n_heavy = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (indptr[i + 1] > indptr[i]) {
            mass[i] += each + (rank < first ? 1 : 0);
            ++rank;
        }
        if (mass[i] < capacity) light[n_light++] = (uint32_t)i; else heavy[n_heavy++] = (uint32_t)i;
        thr[i] = 0xFFFFFFFFu; /* full bucket unless the sweep says otherwise */
        alias[i] = (uint32_t)i;
    }
    if (n_light && !n_heavy) { free(mass); free(light); free(heavy); return -4; } /* impossible: sum m = n c */
    uint64_t j = 0;
    uint64_t rest = n_heavy ? mass[heavy[0]] : 0; /* what the current heavy node still holds */
    for (uint64_t k = 0; k < n_light; ++k) {
        const uint32_t l = light[k];
        thr[l] = threshold(mass[l], capacity);
        alias[l] = heavy[j];
        rest -= capacity - mass[l];
        while (rest < capacity && j + 1 < n_heavy) { /* exhausted: it owns its bucket with `rest` */
            thr[heavy[j]] = threshold(rest, capacity);
            alias[heavy[j]] = heavy[j + 1];
            rest = mass[heavy[j + 1]] - (capacity - rest);
            ++j;
        }
    }
    free(mass); free(light); free(heavy);
    return 0;
}
