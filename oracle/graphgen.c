/*
 * oracle/graphgen.c -- TEST / BENCH INFRASTRUCTURE: the seeded synthetic graphs of the
 * BASELINE.json shapes (Erdos-Renyi G(n, m), R-MAT), generated on the host cores with OpenMP
 * so that the CPU arm of bench.py (`--impl reference`) builds the SAME graph as the GPU arm
 * without loading the product library.
 *
 * The reference tree has no generator (its only in-tree graph idiom is the GraphBuilder loop of
 * /root/reference/embiggen/utils/networkx_utils.py:79-113); the shapes come from BASELINE.json
 * and the definition is SURVEY.md 8(d): undirected, simple, rows sorted ascending.  Definition
 * (shared with embiggen_b200/graph.py and csrc/graph_build.cu, which tests compare): the graph
 * holds the first m distinct undirected edges, in draw order, of the Philox stream
 * (seed, draw index) -- tag 0x10: word 0/1 -> endpoints (Erdos-Renyi); tag 0x11: one word per
 * R-MAT level, four levels per block.
 *
 * Method here (differs from the sort-based GPU builder, same result): draws are consumed in
 * order, in batches of exactly `m - have` draws -- a batch can therefore never overshoot m --
 * and inserted into a lock-free open-addressing hash set; the batch that brings the set to m
 * ends on a new edge, so the prefix is the minimal one.  The CSR is a counting sort by source
 * followed by a per-row sort.
 */
#include "oracle.h"
#include "philox.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GG_EMPTY (~(uint64_t)0)
#define GG_TAG_ER 0x10u
#define GG_TAG_RMAT 0x11u

static inline uint64_t gg_draw(int kind, uint64_t n, uint32_t scale, uint32_t seed_lo, uint32_t seed_hi,
                               uint64_t t_a, uint64_t t_ab, uint64_t t_abc, uint64_t idx) {
    const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
    uint64_t u = 0, v = 0;
    uint32_t r[4];
    if (kind == 0) {
        orc_philox4x32_10(seed_lo, seed_hi, lo, hi, 0u, GG_TAG_ER << 24, r);
        u = orc_mulhi(r[0], (uint32_t)n);
        v = orc_mulhi(r[1], (uint32_t)n);
    } else {
        for (uint32_t block = 0; block * 4u < scale; ++block) {
            orc_philox4x32_10(seed_lo, seed_hi, lo, hi, block, GG_TAG_RMAT << 24, r);
            for (uint32_t level = 0; level < 4u && block * 4u + level < scale; ++level) {
                const uint64_t w = r[level];
                u = (u << 1) | (uint64_t)(w >= t_ab);
                v = (v << 1) | (uint64_t)((w >= t_a && w < t_ab) || w >= t_abc);
            }
        }
    }
    if (u == v || u >= n || v >= n) return GG_EMPTY;
    return u < v ? (u << 32) | v : (v << 32) | u;
}

static inline uint64_t gg_hash(uint64_t key) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdull;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ull;
    key ^= key >> 33;
    return key;
}

/* 1 when `key` was not in the set yet */
static inline int gg_insert(uint64_t *table, uint64_t mask, uint64_t key) {
    uint64_t slot = gg_hash(key) & mask;
    for (;;) {
        uint64_t seen = __atomic_load_n(table + slot, __ATOMIC_RELAXED);
        if (seen == key) return 0;
        if (seen == GG_EMPTY) {
            uint64_t expected = GG_EMPTY;
            if (__atomic_compare_exchange_n(table + slot, &expected, key, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                return 1;
            if (expected == key) return 0;
        }
        slot = (slot + 1) & mask;
    }
}

static int gg_cmp_u32(const void *a, const void *b) {
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

static void gg_sort_row(uint32_t *row, uint64_t len) {
    if (len < 2) return;
    if (len <= 24) {
        for (uint64_t i = 1; i < len; ++i) {
            const uint32_t x = row[i];
            uint64_t j = i;
            while (j > 0 && row[j - 1] > x) { row[j] = row[j - 1]; --j; }
            row[j] = x;
        }
        return;
    }
    qsort(row, len, sizeof(uint32_t), gg_cmp_u32);
}

int orc_synthetic_csr(int kind, uint64_t n, uint32_t scale, uint64_t m, uint64_t seed, uint64_t t_a,
                      uint64_t t_ab, uint64_t t_abc, int64_t *indptr, uint32_t *indices, uint64_t *nnz_out) {
    if (!indptr || !indices || !nnz_out || n < 2 || n >= 0xFFFFFF00ull) return -1;
    if (kind != 0 && kind != 1) return -1;
    if (kind == 1 && (scale == 0 || scale > 32 || (scale < 32 && n > ((uint64_t)1 << scale)))) return -1;
    if ((double)m > 0.25 * (double)n * (double)(n - 1)) return -1;
    const int threads = orc_get_threads();
    const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
    uint64_t capacity = 1024;
    while (capacity < 2 * m + 2) capacity <<= 1;
    const uint64_t mask = capacity - 1;
    uint64_t *table = (uint64_t *)malloc(capacity * sizeof(uint64_t));
    if (!table) return -3;
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(static)
    for (uint64_t s = 0; s < capacity; ++s) table[s] = GG_EMPTY;

    uint64_t have = 0, drawn = 0;
    for (uint64_t round = 0; have < m; ++round) {
        if (round > (1u << 22)) { free(table); return -2; }
        const uint64_t want = m - have; /* never overshoots: a draw yields at most one edge */
        uint64_t fresh = 0;
#pragma omp parallel for num_threads(threads) if (threads > 1 && want > 4096) schedule(static) reduction(+ : fresh)
        for (uint64_t k = 0; k < want; ++k) {
            const uint64_t key = gg_draw(kind, n, scale, seed_lo, seed_hi, t_a, t_ab, t_abc, drawn + k);
            if (key != GG_EMPTY) fresh += (uint64_t)gg_insert(table, mask, key);
        }
        drawn += want;
        have += fresh;
    }

    /* counting sort by source: degrees, offsets, scatter, then sort every row */
    uint32_t *cursor = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
    if (!cursor) { free(table); return -3; }
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(static)
    for (uint64_t s = 0; s < capacity; ++s) {
        const uint64_t key = table[s];
        if (key == GG_EMPTY) continue;
        __atomic_fetch_add(cursor + (key >> 32), 1u, __ATOMIC_RELAXED);
        __atomic_fetch_add(cursor + (uint32_t)key, 1u, __ATOMIC_RELAXED);
    }
    indptr[0] = 0;
    for (uint64_t v = 0; v < n; ++v) {
        indptr[v + 1] = indptr[v] + (int64_t)cursor[v];
        cursor[v] = 0;
    }
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(static)
    for (uint64_t s = 0; s < capacity; ++s) {
        const uint64_t key = table[s];
        if (key == GG_EMPTY) continue;
        const uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        indices[indptr[a] + (int64_t)__atomic_fetch_add(cursor + a, 1u, __ATOMIC_RELAXED)] = b;
        indices[indptr[b] + (int64_t)__atomic_fetch_add(cursor + b, 1u, __ATOMIC_RELAXED)] = a;
    }
    free(table);
    free(cursor);
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(dynamic, 4096)
    for (uint64_t v = 0; v < n; ++v) gg_sort_row(indices + indptr[v], (uint64_t)(indptr[v + 1] - indptr[v]));
    *nnz_out = 2 * m;
    return 0;
}
