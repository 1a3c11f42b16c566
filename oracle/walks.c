/*
 * oracle/walks.c -- TEST INFRASTRUCTURE: CPU reference sampler for DeepWalk /
 * node2vec walks.  The GPU walk kernel must match it bit for bit.
 *
 * Follows: the p/q semantics of the reference docstrings
 * (/root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:58-71,
 * return_weight = 1/p, explore_weight = 1/q), the sorted-CSR hand-off of
 * /root/reference/embiggen/embedders/pecanpy_embedders/node2vec.py:139-163, and
 * the rejection sampler north_star prescribes (KnightKing, SOSP'19): uniform
 * neighbour proposal, integer accept test.  DeepWalk = both weights 1
 * (/root/reference/embiggen/embedders/ensmallen_embedders/deepwalk_skipgram.py:6-139).
 *
 * The oracle always classifies a proposal the plain way (return / common /
 * explore, full binary search); the counters additionally model the
 * shortcuts a fast implementation may take (return test first, then
 * lower-/upper-bound pre-decision) so that algorithmic bytes can be
 * accounted.  Shortcuts never change a decision, which is exactly what
 * bit-exact parity of the GPU kernel demonstrates.
 */
#include "oracle.h"
#include "philox.h"
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, c2, c3, out);
}

/* blocks (c0 = first_c0 + i, c1, c2, c3) for i < count; out holds 4 * count words */
void orc_philox_range(uint64_t seed, uint32_t first_c0, uint64_t count, uint32_t c1, uint32_t c2,
                      uint32_t c3, uint32_t *out) {
    for (uint64_t i = 0; i < count; ++i)
        orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), first_c0 + (uint32_t)i, c1, c2, c3,
                          out + 4 * i);
}

uint64_t orc_sources(const int64_t *indptr, uint64_t n, uint32_t *out) {
    uint64_t count = 0;
    for (uint64_t v = 0; v < n; ++v) {
        if (indptr[v + 1] > indptr[v]) {
            if (out) out[count] = (uint32_t)v;
            ++count;
        }
    }
    return count;
}

void orc_thresholds(float return_weight, float explore_weight, uint64_t thr[3]) {
    const double w[3] = {(double)return_weight, 1.0, (double)explore_weight};
    double wmax = w[0];
    if (w[1] > wmax) wmax = w[1];
    if (w[2] > wmax) wmax = w[2];
    for (int i = 0; i < 3; ++i) {
        if (w[i] >= wmax) {
            thr[i] = 4294967296ull;
        } else {
            double t = floor(w[i] / wmax * 4294967296.0);
            thr[i] = t >= 4294967296.0 ? 4294967296ull : (uint64_t)t;
        }
    }
}

/*
 * Return-edge folding (KnightKing's "outlier" folding, SOSP'19 section 4.2).  With
 * return_weight > max(1, explore_weight) the plain envelope max(rw, 1, ew) is set by ONE edge of
 * the row -- the way back -- and every other proposal is accepted with probability <= 1 / rw.
 * On an undirected, unweighted graph the way back always exists, so its excess is cut off the
 * envelope and appended to the row as a virtual slot of mass e = (rw - wenv) / wenv, where
 * wenv = max(1, ew) is the envelope of the other classes:
 *   trial k of step t: ONE Philox block (tag 8, c2 = t - 1, c3 = k) -> words x, y, z;
 *     x <  T_out(d)  : the dart fell into the virtual slot: go back (accepted at once);
 *     otherwise      : propose idx = mulhi(y, d); accept iff z < thrf[class], with
 *                      thrf = {2^32, floor(2^32 / wenv), floor(2^32 ew / wenv)} (2^32 when the
 *                      ratio is 1);
 *   T_out(d) = floor(2^32 E / (d 2^20 + E)),  E = min(floor(e 2^20), 2^31)   -- integers only.
 * P(back) : P(common y) : P(explore y) = (e + 1) : 1 / wenv : ew / wenv = rw : 1 : ew, as before,
 * in about half the trials (C3: 3.7 -> 1.9 per step).  Directed or weighted graphs, typed walks
 * and rw <= wenv keep the plain envelope (tag 2).
 */
void orc_fold_thresholds(float return_weight, float explore_weight, uint64_t thrf[3], uint64_t *excess) {
    const double rw = (double)return_weight, ew = (double)explore_weight;
    const double wenv = ew > 1.0 ? ew : 1.0;
    const double w[3] = {wenv, 1.0, ew};
    for (int i = 0; i < 3; ++i) {
        if (w[i] >= wenv) {
            thrf[i] = 4294967296ull;
        } else {
            double t = floor(w[i] / wenv * 4294967296.0);
            thrf[i] = t >= 4294967296.0 ? 4294967296ull : (uint64_t)t;
        }
    }
    double e = rw > wenv ? floor((rw - wenv) / wenv * 1048576.0) : 0.0;
    if (e > 2147483648.0) e = 2147483648.0;
    *excess = (uint64_t)e;
}

/* 1 when every edge has its mirror (u in N(v) <=> v in N(u)) */
static int row_contains(const uint32_t *row, uint64_t len, uint32_t key);
int orc_is_undirected(const int64_t *indptr, const uint32_t *indices, uint64_t n) {
    if (!indptr || !indices) return 0;
    int symmetric = 1;
    const int threads = orc_get_threads();
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(dynamic, 4096) reduction(&& : symmetric)
    for (uint64_t u = 0; u < n; ++u) {
        if (!symmetric) continue;
        for (int64_t e = indptr[u]; e < indptr[u + 1]; ++e) {
            const uint32_t v = indices[e];
            if (v >= n || !row_contains(indices + indptr[v], (uint64_t)(indptr[v + 1] - indptr[v]), (uint32_t)u)) {
                symmetric = 0;
                break;
            }
        }
    }
    return symmetric;
}

/* sorted-row membership, plain lower-bound bisection */
static int row_contains(const uint32_t *row, uint64_t len, uint32_t key) {
    uint64_t lo = 0, hi = len;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (row[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < len && row[lo] == key;
}

static uint64_t probe_sectors(uint64_t len) {
    /* ceil(log2(len + 1)) - 2, at least 1: the last three levels share a 32 B sector */
    uint64_t levels = 0;
    while (((uint64_t)1 << levels) < len + 1) ++levels;
    return levels > 3 ? levels - 2 : 1;
}

/*
 * Edge weights ("next" row f-2; the reference's classes advertise them:
 * /root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py:119-129).  Proposals are
 * drawn proportionally to the static weight in O(1) through a per-row Vose alias table; second
 * order keeps the p/q accept test on top (KnightKing: proposal by the static weight, acceptance
 * by the dynamic bias).  Normative construction, per row of d edges (the product's host-side
 * builder must reproduce it bit for bit): scaled_i = w_i * d / total with total the
 * left-to-right double sum of the row; small / large are LIFO stacks filled in ascending edge
 * order; thr_i = min(floor(prob_i * 2^32), 2^32 - 1); leftovers keep thr = 2^32 - 1 and alias =
 * self, except a leftover of weight zero, which gets thr = 0 and the row's heaviest edge as alias
 * (a zero-weight edge is never proposed); a row of total weight zero is uniform.  The table holds
 * two words per edge: {thr, alias index inside the row}.
 * Sampling with ONE random word r: u = r * d (64 bit); i = u >> 32; the low word of u is the
 * coin: edge = low < thr_i ? i : alias_i.
 */
int orc_edge_alias(const int64_t *indptr, const float *weights, uint64_t n, uint32_t *table) {
    if (!indptr || !weights || !table) return -1;
    uint64_t max_degree = 0;
    for (uint64_t v = 0; v < n; ++v)
        if ((uint64_t)(indptr[v + 1] - indptr[v]) > max_degree) max_degree = (uint64_t)(indptr[v + 1] - indptr[v]);
    double *scaled = (double *)malloc((max_degree + 1) * sizeof(double));
    uint32_t *small = (uint32_t *)malloc((max_degree + 1) * sizeof(uint32_t));
    uint32_t *large = (uint32_t *)malloc((max_degree + 1) * sizeof(uint32_t));
    if (!scaled || !small || !large) { free(scaled); free(small); free(large); return -3; }
    int status = 0;
    for (uint64_t v = 0; v < n && status == 0; ++v) {
        const int64_t begin = indptr[v];
        const uint64_t d = (uint64_t)(indptr[v + 1] - begin);
        uint32_t *row = table + 2 * begin;
        double total = 0.0;
        uint64_t heaviest = 0;
        for (uint64_t i = 0; i < d; ++i) {
            const float w = weights[begin + i];
            if (!(w >= 0.0f)) { status = -2; break; } /* negative or NaN */
            total += (double)w;
            if (w > weights[begin + heaviest]) heaviest = i;
        }
        if (status) break;
        uint64_t ns = 0, nl = 0;
        for (uint64_t i = 0; i < d; ++i) {
            scaled[i] = total > 0.0 ? (double)weights[begin + i] * (double)d / total : 1.0;
            row[2 * i] = 0xFFFFFFFFu;
            row[2 * i + 1] = (uint32_t)i;
            if (scaled[i] < 1.0) small[ns++] = (uint32_t)i; else large[nl++] = (uint32_t)i;
        }
        while (ns > 0 && nl > 0) {
            const uint32_t s = small[--ns];
            const uint32_t l = large[--nl];
            const double t = floor(scaled[s] * 4294967296.0);
            row[2 * s] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
            row[2 * s + 1] = l;
            scaled[l] = (scaled[l] + scaled[s]) - 1.0;
            if (scaled[l] < 1.0) small[ns++] = l; else large[nl++] = l;
        }
        for (uint64_t k = 0; k < ns; ++k) { /* leftovers of weight zero are never proposed */
            const uint32_t s = small[k];
            if (total > 0.0 && weights[begin + s] == 0.0f) { row[2 * s] = 0; row[2 * s + 1] = (uint32_t)heaviest; }
        }
    }
    free(scaled); free(small); free(large);
    return status;
}

/* index of the proposal inside a row: uniform (table == NULL) or by the row's alias table */
static uint32_t propose(const uint32_t *table_row, uint32_t deg, uint32_t r) {
    const uint64_t u = (uint64_t)r * deg;
    const uint32_t i = (uint32_t)(u >> 32);
    if (!table_row) return i;
    return (uint32_t)u < table_row[2 * i] ? i : table_row[2 * i + 1];
}

static int walks_plain(const int64_t *indptr, const uint32_t *indices, const uint32_t *table, uint64_t n,
                       const uint32_t *sources, uint64_t n_src, uint64_t seed, uint64_t first_walk,
                       uint64_t n_walks, uint64_t walk_id_stride, uint32_t walk_length,
                       float return_weight, float explore_weight, int undirected, uint32_t *out,
                       orc_walk_counters *counters);

int orc_walks(const int64_t *indptr, const uint32_t *indices, uint64_t n, const uint32_t *sources,
              uint64_t n_src, uint64_t seed, uint64_t first_walk, uint64_t n_walks,
              uint64_t walk_id_stride, uint32_t walk_length, float return_weight,
              float explore_weight, int undirected, uint32_t *out, orc_walk_counters *counters) {
    return orc_walks_weighted(indptr, indices, NULL, n, sources, n_src, seed, first_walk, n_walks,
                              walk_id_stride, walk_length, return_weight, explore_weight, undirected, out,
                              counters);
}

int orc_walks_weighted(const int64_t *indptr, const uint32_t *indices, const uint32_t *table, uint64_t n,
                       const uint32_t *sources, uint64_t n_src, uint64_t seed, uint64_t first_walk,
                       uint64_t n_walks, uint64_t walk_id_stride, uint32_t walk_length,
                       float return_weight, float explore_weight, int undirected, uint32_t *out,
                       orc_walk_counters *counters) {
    return walks_plain(indptr, indices, table, n, sources, n_src, seed, first_walk, n_walks,
                       walk_id_stride, walk_length, return_weight, explore_weight, undirected, out, counters);
}

/*
 * Typed walks ("next" row f-2, .../node2vec_skipgram.py:72-77).  Every transition, the first one
 * included, is a trial loop on its own Philox stream (tag 7, ONE block per trial: c2 = t - 1,
 * c3 = trial):
 *   word 0  proposal (uniform, or proportional to the edge weight);
 *   word 2  node-type test: the weight of v -> x is multiplied by change_node_type_weight when
 *           type(x) != type(v); accept iff r2 < q_node[changed];
 *   word 3  edge-type test (from the second transition on): multiplied by
 *           change_edge_type_weight when type(v -> x) != type(prev -> v); accept iff r3 < q_edge[changed];
 *   word 1  p/q test: accept iff r1 < thr[class]   (thr = 2^32 for the first transition).
 * q[changed] = floor(w / max(1, w) * 2^32), q[same] = floor(1 / max(1, w) * 2^32).  The three
 * tests use independent words, so the acceptance probability is the product of the three
 * ratios and the walk follows  weight * bias_pq * type factors  exactly (up to 2^-32).
 * `normalize_by_degree` (.../node2vec_skipgram.py:94-96: the transition weight divided by the
 * degree of the destination) is not a rejection test at all: it is folded into the proposal,
 * whose per-edge table is built over  weight / max(deg(destination), 1)  (orc_edge_alias), so it
 * costs no extra trials however skewed the degrees are.
 * Integer arithmetic only; the cheap tests come first and the adjacency search is counted only
 * when it is reached and undecided.
 */
void orc_type_thresholds(float change_weight, uint64_t q[2]) {
    const double w = (double)change_weight, m = w > 1.0 ? w : 1.0;
    const double ratio[2] = {1.0 / m, w / m}; /* [same, changed] */
    for (int i = 0; i < 2; ++i) {
        double t = floor(ratio[i] * 4294967296.0);
        q[i] = t >= 4294967296.0 ? 4294967296ull : (uint64_t)t;
    }
}

static int walks_general(const int64_t *indptr, const uint32_t *indices, const uint32_t *table,
                         const uint32_t *node_types, const uint32_t *edge_types,
                         float change_node_type_weight, float change_edge_type_weight,
                         const uint32_t *sources, uint64_t n_src, uint64_t seed,
                         uint64_t first_walk, uint64_t n_walks, uint64_t walk_id_stride,
                         uint32_t walk_length, float return_weight, float explore_weight,
                         uint32_t *out, orc_walk_counters *counters) {
    const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
    uint64_t thr[3], qn[2], qe[2];
    orc_thresholds(return_weight, explore_weight, thr);
    orc_type_thresholds(node_types ? change_node_type_weight : 1.0f, qn);
    orc_type_thresholds(edge_types ? change_edge_type_weight : 1.0f, qe);
    const int use_nt = node_types && qn[0] != qn[1];
    const int use_et = edge_types && qe[0] != qe[1];
    const uint64_t thr_lo = thr[1] < thr[2] ? thr[1] : thr[2];
    const uint64_t thr_hi = thr[1] < thr[2] ? thr[2] : thr[1];
    uint64_t n_steps = 0, n_trials = 0, n_searches = 0, n_probe = 0, n_capped = 0;
    const int threads = orc_get_threads();
#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(dynamic, 256) \
    reduction(+ : n_steps, n_trials, n_searches, n_probe, n_capped)
    for (uint64_t i = 0; i < n_walks; ++i) {
        const uint64_t wid = first_walk + i * walk_id_stride;
        const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
        uint32_t *walk = out + i * (uint64_t)walk_length;
        uint32_t cur = sources[wid % n_src], prev = ORC_PAD_TOKEN, prev_etype = 0;
        walk[0] = cur;
        uint32_t t = 1;
        for (; t < walk_length; ++t) {
            const int64_t off = indptr[cur];
            const uint64_t deg = (uint64_t)(indptr[cur + 1] - off);
            if (deg == 0) break;
            uint32_t next, trial = 0;
            int64_t e;
            for (;;) {
                uint32_t rnd[4];
                orc_philox4x32_10(seed_lo, seed_hi, wid_lo, wid_hi, t - 1,
                                  (ORC_TAG_WALK3 << 24) | trial, rnd);
                e = off + propose(table ? table + 2 * off : NULL, (uint32_t)deg, rnd[0]);
                next = indices[e];
                ++n_trials;
                int accept = 1;
                if (use_nt && (uint64_t)rnd[2] >= qn[node_types[next] != node_types[cur]]) accept = 0;
                if (accept && use_et && t > 1 && (uint64_t)rnd[3] >= qe[edge_types[e] != prev_etype])
                    accept = 0;
                if (accept) {
                    const uint64_t lhs = rnd[1];
                    uint64_t limit = 4294967296ull; /* first transition: no p/q bias */
                    if (t > 1) {
                        int cls;
                        if (next == prev) {
                            cls = 0;
                        } else {
                            const int64_t poff = indptr[prev];
                            const uint64_t pdeg = (uint64_t)(indptr[prev + 1] - poff);
                            cls = row_contains(indices + poff, pdeg, next) ? 1 : 2;
                            if (lhs >= thr_lo && lhs < thr_hi) {
                                ++n_searches;
                                n_probe += probe_sectors(pdeg);
                            }
                        }
                        limit = thr[cls];
                    }
                    accept = lhs < limit;
                }
                if (accept) break;
                ++trial;
                if (trial >= ORC_MAX_TRIALS) { ++n_capped; break; }
            }
            ++n_steps;
            walk[t] = next;
            if (edge_types) prev_etype = edge_types[e];
            prev = cur;
            cur = next;
        }
        for (; t < walk_length; ++t) walk[t] = ORC_PAD_TOKEN;
    }
    if (counters) {
        counters->steps = n_steps; counters->trials = n_trials; counters->first_order = 0;
        counters->searches = n_searches; counters->probe_sectors = n_probe; counters->capped = n_capped;
    }
    return 0;
}

int orc_walks_typed(const int64_t *indptr, const uint32_t *indices, const uint32_t *table,
                    const uint32_t *node_types, const uint32_t *edge_types,
                    float change_node_type_weight, float change_edge_type_weight, uint64_t n,
                    const uint32_t *sources, uint64_t n_src, uint64_t seed, uint64_t first_walk,
                    uint64_t n_walks, uint64_t walk_id_stride, uint32_t walk_length,
                    float return_weight, float explore_weight, int undirected, uint32_t *out,
                    orc_walk_counters *counters) {
    if (!indptr || !indices || !sources || !out || n_src == 0 || walk_length == 0) return -1;
    const int typed = (node_types && change_node_type_weight != 1.0f) ||
                      (edge_types && change_edge_type_weight != 1.0f);
    if (!typed)
        return orc_walks_weighted(indptr, indices, table, n, sources, n_src, seed, first_walk, n_walks,
                                  walk_id_stride, walk_length, return_weight, explore_weight, undirected, out,
                                  counters);
    return walks_general(indptr, indices, table, node_types, edge_types, change_node_type_weight,
                         change_edge_type_weight, sources, n_src, seed, first_walk, n_walks,
                         walk_id_stride, walk_length, return_weight, explore_weight, out, counters);
}

static int walks_plain(const int64_t *indptr, const uint32_t *indices, const uint32_t *table, uint64_t n,
                       const uint32_t *sources, uint64_t n_src, uint64_t seed, uint64_t first_walk,
                       uint64_t n_walks, uint64_t walk_id_stride, uint32_t walk_length,
                       float return_weight, float explore_weight, int undirected, uint32_t *out,
                       orc_walk_counters *counters) {
    if (!indptr || !indices || !sources || !out || n_src == 0 || walk_length == 0) return -1;
    (void)n;
    const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
    const int second_order = !(return_weight == 1.0f && explore_weight == 1.0f);
    uint64_t thr[3];
    orc_thresholds(return_weight, explore_weight, thr);
    uint64_t excess = 0;
    {
        uint64_t thrf[3];
        orc_fold_thresholds(return_weight, explore_weight, thrf, &excess);
        if (!(undirected && !table && second_order)) excess = 0;
        if (excess) { thr[0] = thrf[0]; thr[1] = thrf[1]; thr[2] = thrf[2]; }
    }
    const uint64_t thr_lo = thr[1] < thr[2] ? thr[1] : thr[2];
    const uint64_t thr_hi = thr[1] < thr[2] ? thr[2] : thr[1];
    uint64_t n_steps = 0, n_trials = 0, n_first = 0, n_searches = 0, n_probe = 0, n_capped = 0;
    const int threads = orc_get_threads();

#pragma omp parallel for num_threads(threads) if (threads > 1) schedule(dynamic, 256) \
    reduction(+ : n_steps, n_trials, n_first, n_searches, n_probe, n_capped)
    for (uint64_t i = 0; i < n_walks; ++i) {
        orc_walk_counters c = {0, 0, 0, 0, 0, 0};
        const uint64_t wid = first_walk + i * walk_id_stride;
        const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
        uint32_t *walk = out + i * (uint64_t)walk_length;
        uint32_t cur = sources[wid % n_src];
        uint32_t prev = ORC_PAD_TOKEN;
        walk[0] = cur;
        uint32_t rnd[4] = {0, 0, 0, 0};
        uint32_t t = 1;
        for (; t < walk_length; ++t) {
            const int64_t off = indptr[cur];
            const uint64_t deg = (uint64_t)(indptr[cur + 1] - off);
            if (deg == 0) break;
            uint32_t next;
            if (!second_order || t == 1) {
                const uint32_t s = t - 1;
                orc_philox4x32_10(seed_lo, seed_hi, wid_lo, wid_hi, s >> 2, ORC_TAG_WALK1 << 24,
                                  rnd);
                next = indices[off + propose(table ? table + 2 * off : NULL, (uint32_t)deg, rnd[s & 3])];
                ++c.first_order;
            } else {
                const int64_t poff = indptr[prev];
                const uint64_t pdeg = (uint64_t)(indptr[prev + 1] - poff);
                uint32_t trial = 0;
                const uint64_t t_out = excess ? (excess << 32) / ((deg << 20) + excess) : 0;
                for (;;) {
                    uint32_t r0, r1;
                    if (excess) { /* folded return edge: one block per trial */
                        orc_philox4x32_10(seed_lo, seed_hi, wid_lo, wid_hi, t - 1,
                                          (ORC_TAG_FOLD << 24) | trial, rnd);
                        if ((uint64_t)rnd[0] < t_out) { next = prev; ++c.trials; break; }
                        r0 = rnd[1];
                        r1 = rnd[2];
                    } else {
                        if ((trial & 1u) == 0)
                            orc_philox4x32_10(seed_lo, seed_hi, wid_lo, wid_hi, t - 1,
                                              (ORC_TAG_WALK2 << 24) | (trial >> 1), rnd);
                        r0 = rnd[2 * (trial & 1u)];
                        r1 = rnd[2 * (trial & 1u) + 1];
                    }
                    next = indices[off + propose(table ? table + 2 * off : NULL, (uint32_t)deg, r0)];
                    ++c.trials;
                    int cls;
                    if (next == prev) {
                        cls = 0;
                    } else {
                        cls = row_contains(indices + poff, pdeg, next) ? 1 : 2;
                        if ((uint64_t)r1 >= thr_lo && (uint64_t)r1 < thr_hi) {
                            ++c.searches;
                            c.probe_sectors += probe_sectors(pdeg);
                        }
                    }
                    if ((uint64_t)r1 < thr[cls]) break;
                    ++trial;
                    if (trial >= ORC_MAX_TRIALS) { ++c.capped; break; }
                }
            }
            ++c.steps;
            walk[t] = next;
            prev = cur;
            cur = next;
        }
        for (; t < walk_length; ++t) walk[t] = ORC_PAD_TOKEN;
        n_steps += c.steps; n_trials += c.trials; n_first += c.first_order;
        n_searches += c.searches; n_probe += c.probe_sectors; n_capped += c.capped;
    }
    if (counters) {
        counters->steps = n_steps; counters->trials = n_trials; counters->first_order = n_first;
        counters->searches = n_searches; counters->probe_sectors = n_probe;
        counters->capped = n_capped;
    }
    return 0;
}
