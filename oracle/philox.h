/*
 * oracle/philox.h -- TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
 *
 * Philox4x32-10 counter-based generator (Salmon et al., "Parallel random
 * numbers: as easy as 1, 2, 3", SC'11).  The reference tree holds no RNG for
 * this path (the arithmetic sits in the un-vendored `ensmallen` wheel,
 * /root/reference/setup.py:76); BASELINE.json:north_star prescribes "a
 * counter-based Philox RNG keyed by (seed, walk, step)", SURVEY.md App. C.1.
 * Pinned by the Random123 known-answer vectors in tests/test_philox.py.
 */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>

#define ORC_TAG_WALK1 1u /* first-order steps, 4 steps per block            */
#define ORC_TAG_WALK2 2u /* second-order trials, 2 trials per block         */
#define ORC_TAG_NEG 3u   /* negative draws                                   */
#define ORC_TAG_INIT0 4u /* table 0 initialisation                           */
#define ORC_TAG_INIT1 5u /* table 1 initialisation                           */
#define ORC_TAG_WALK3 7u /* general walks (normalize_by_degree, typed): 1 trial per block */
#define ORC_TAG_FOLD 12u /* second-order trials with the return edge folded: 1 trial per block */
#define ORC_TAG_SKIP 6u  /* stochastic_downsample_by_degree, one per centre     */

static inline void orc_philox4x32_10(uint32_t seed_lo, uint32_t seed_hi, uint32_t c0, uint32_t c1,
                                     uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = seed_lo, k1 = seed_hi;
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Lemire multiply-shift: uniform index in [0, range) from 32 random bits. */
static inline uint32_t orc_mulhi(uint32_t r, uint32_t range) {
    return (uint32_t)(((uint64_t)r * (uint64_t)range) >> 32);
}
#endif
