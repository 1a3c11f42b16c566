/*
 * oracle/sgns.c -- TEST INFRASTRUCTURE: CPU restatement of the SkipGram / CBOW
 * negative-sampling SGD that `ensmallen.models.SkipGram/CBOW.fit_transform`
 * performs below /root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py:99.
 *
 * Kwarg semantics follow the reference docstrings
 * (/root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:37-119:
 * window trimmed at the borders :55-57, clipping_value :45-47, negatives :48-50,
 * learning_rate / decay :82-85, normalize_learning_rate_by_degree :99-100,
 * use_scale_free_distribution :101-102); the model structure (two tables,
 * centre embedding against output weights; CBOW = mean of the context
 * embeddings) is cross-checked with
 * /root/reference/embiggen/embedders/tensorflow_embedders/skipgram.py:28-61 and
 * cbow.py:28-60.  The exact recipe is the normative spec in DESIGN.md
 * (SURVEY.md App. C.6-C.9): PARITY UNPINNED against Ensmallen, see oracle.h.
 *
 * Floating point is specified operation by operation (fmaf / single IEEE
 * ops, a fixed warp-shaped reduction order) so that a GPU launch that
 * processes walks in the same order reproduces the tables bit for bit.
 * Compile with -ffp-contract=off.
 */
#include "oracle.h"
#include "philox.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void orc_set_threads(int threads) { g_threads = threads < 1 ? 1 : threads; }
int orc_get_threads(void) { return g_threads; }

/* exp(y) from IEEE single ops only: Cody-Waite reduction + degree-6 polynomial */
float orc_exp_det(float y) {
    if (y > 80.0f) y = 80.0f;
    if (y < -80.0f) y = -80.0f;
    const float k = rintf(y * 1.44269504088896341f);
    float r = fmaf(k, -0.693145751953125f, y);
    r = fmaf(k, -1.42860682030941723212e-6f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, r * r, r);
    p = p + 1.0f;
    union { uint32_t u; float f; } scale;
    scale.u = (uint32_t)((int32_t)k + 127) << 23;
    return p * scale.f;
}

float orc_sigmoid(float x) { return 1.0f / (1.0f + orc_exp_det(-x)); }

/* warp-shaped dot: lane l owns float4 chunks l, l+32, ...; xor-butterfly 16,8,4,2,1 */
float orc_dot(const float *a, const float *b, uint32_t row_stride) {
    const uint32_t chunks = row_stride / 4;
    float part[32];
    for (uint32_t lane = 0; lane < 32; ++lane) {
        float p = 0.0f;
        for (uint32_t ch = lane; ch < chunks; ch += 32) {
            const float *x = a + 4 * ch, *y = b + 4 * ch;
            p = fmaf(x[0], y[0], p);
            p = fmaf(x[1], y[1], p);
            p = fmaf(x[2], y[2], p);
            p = fmaf(x[3], y[3], p);
        }
        part[lane] = p;
    }
    /* lanes l and l^off end up with the same sum (IEEE add commutes), so one half suffices */
    for (uint32_t off = 16; off >= 1; off >>= 1)
        for (uint32_t lane = 0; lane < off; ++lane) part[lane] = part[lane] + part[lane + off];
    return part[0];
}

int orc_init_tables(uint64_t n, uint32_t embedding_size, uint32_t row_stride, uint64_t seed,
                    float *t0, float *t1) {
    if (!t0 || !t1 || embedding_size == 0 || row_stride < embedding_size || (row_stride & 3))
        return -1;
    const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
    const float dim = (float)embedding_size;
    for (int table = 0; table < 2; ++table) {
        float *t = table ? t1 : t0;
        const uint32_t tag = (table ? ORC_TAG_INIT1 : ORC_TAG_INIT0) << 24;
        /* every element is a pure function of (seed, table, i, j): the thread count cannot matter */
#pragma omp parallel for num_threads(g_threads) if (g_threads > 1) schedule(static)
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t rnd[4];
            for (uint32_t j = 0; j < row_stride; ++j) {
                if ((j & 3) == 0)
                    orc_philox4x32_10(seed_lo, seed_hi, (uint32_t)i, (uint32_t)(i >> 32), j >> 2,
                                      tag, rnd);
                float v = 0.0f;
                if (j < embedding_size) {
                    const float u01 = (float)(rnd[j & 3] >> 8) * 5.9604644775390625e-8f; /* 2^-24 */
                    v = (u01 - 0.5f) / dim;
                }
                t[i * row_stride + j] = v;
            }
        }
    }
    return 0;
}

static double softplus(double z) { return z > 30.0 ? z : log1p(exp(z)); }

typedef struct {
    double loss;
    uint64_t pairs, targets;
} train_acc;

#define MAX_TARGETS 65

/* negatives for one draw site; returns validity mask semantics via valid[] */
static void draw_negatives(const orc_sgns_cfg *cfg, uint64_t seed, uint64_t wid, uint32_t site,
                           uint64_t n, const uint32_t *thr, const uint32_t *alias, uint32_t centre,
                           uint32_t context, uint32_t *neg, int *valid) {
    const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
    for (uint32_t k = 0; k < cfg->negatives; ++k) {
        uint32_t rnd[4];
        orc_philox4x32_10(seed_lo, seed_hi, (uint32_t)wid, (uint32_t)(wid >> 32), site,
                          (ORC_TAG_NEG << 24) | k, rnd);
        const uint32_t idx = orc_mulhi(rnd[0], (uint32_t)n);
        uint32_t node = idx;
        if (cfg->use_alias) node = rnd[1] < thr[idx] ? idx : alias[idx];
        neg[k] = node;
        int ok = node != centre && node != context;
        for (uint32_t q = 0; q < k && ok; ++q) ok = neg[q] != node;
        valid[k] = ok;
    }
}

/*
 * Score `h` against targets (rows of t1), batch semantics: every dot uses the
 * pre-update rows; then targets are applied in order.  acc receives
 * sum_k g_k * row_k (pre-update).
 */
/* cfg->fast_math: the dot as a plain vectorisable loop and libm's expf -- the same algorithm at the
 * speed a CPU implementation would be written for; used only to TIME the CPU baseline (bench.py),
 * never for parity (it is not bit-comparable with the GPU's warp-shaped arithmetic). */
static float fast_dot(const float *a, const float *b, uint32_t count) {
    float sum = 0.0f;
#pragma omp simd reduction(+ : sum)
    for (uint32_t e = 0; e < count; ++e) sum += a[e] * b[e];
    return sum;
}

static void apply_targets(const orc_sgns_cfg *cfg, float lr, float inv_scale, const float *h,
                          float *t1, const uint32_t *target, const int *valid, uint32_t count,
                          float *acc, train_acc *out) {
    const uint32_t stride = cfg->row_stride;
    float f[MAX_TARGETS];
    for (uint32_t k = 0; k < count; ++k) {
        if (!valid[k]) continue;
        f[k] = cfg->fast_math ? fast_dot(h, t1 + (uint64_t)target[k] * stride, stride)
                              : orc_dot(h, t1 + (uint64_t)target[k] * stride, stride);
        if (cfg->scale_by_sqrt_dim) f[k] = f[k] * inv_scale;
    }
    for (uint32_t k = 0; k < count; ++k) {
        if (!valid[k]) continue;
        ++out->targets;
        if (fabsf(f[k]) > cfg->clipping_value) continue;
        const float label = k == 0 ? 1.0f : 0.0f;
        const float g = (label - (cfg->fast_math ? 1.0f / (1.0f + expf(-f[k])) : orc_sigmoid(f[k]))) * lr;
        out->loss += softplus(k == 0 ? -(double)f[k] : (double)f[k]);
        float *row = t1 + (uint64_t)target[k] * stride;
        if (cfg->fast_math) {
#pragma omp simd
            for (uint32_t e = 0; e < stride; ++e) {
                const float old = row[e];
                acc[e] += g * old;
                row[e] = old + g * h[e];
            }
            continue;
        }
        for (uint32_t e = 0; e < stride; ++e) {
            const float old = row[e];
            acc[e] = fmaf(g, old, acc[e]);
            row[e] = fmaf(g, h[e], old);
        }
    }
}

/*
 * SkipGram with shared negatives (cfg->shared_negatives; north_star's opt-in "shared-negative
 * batching"): the draw site is the CENTRE, not the pair.  One set of K negatives (the CBOW site
 * key, (i << 16) | 0xFFFF) serves the m pairs of centre i, and all m + K targets are scored
 * against the same h = T0[c] before anything is updated (batch semantics):
 *   context at window position q:  g_q = (1 - sigmoid(f_q)) lr;  T1[o_q] += fl(g_q h)   (mul, then
 *       add: the product is what the GPU hands to the L2 atomic adder); a token at k positions
 *       receives its k additions one after the other;
 *   negative s (valid: not the centre, not a repeat, not a context token of this window):
 *       g_s = ((0 - sigmoid(f_s)) lr) m  -- it stands for the m pairs that would each have drawn
 *       their own; T1[n_s] = fma(g_s, h, T1[n_s]);
 *   T0[c] = h + (sum_q g_q T1[o_q] + sum_s g_s T1[n_s]), rows as they were before the update,
 *       each sum accumulated with fma in the order written.
 * Rows moved per pair fall from K + 1 (+ the centre's share) to (K + 2) / m: DESIGN.md section 5.
 */
static void train_centre_shared(const orc_sgns_cfg *cfg, float lr, float inv_scale, uint32_t c,
                                const uint32_t *ctx, uint32_t m, uint64_t seed, uint64_t wid,
                                uint32_t i, uint64_t n, const uint32_t *thr, const uint32_t *alias,
                                float *t0, float *t1, float *h, float *acc, train_acc *out) {
    const uint32_t stride = cfg->row_stride, K = cfg->negatives;
    uint32_t neg[MAX_TARGETS];
    int valid[MAX_TARGETS];
    float fp[2 * 64], fn[MAX_TARGETS], gp[2 * 64];
    float *crow = t0 + (uint64_t)c * stride;
    memcpy(h, crow, stride * sizeof(float));
    draw_negatives(cfg, seed, wid, (i << 16) | 0xFFFFu, n, thr, alias, c, c, neg, valid);
    for (uint32_t k = 0; k < K; ++k)
        for (uint32_t q = 0; q < m && valid[k]; ++q) valid[k] = neg[k] != ctx[q];
    /* cfg->fast_math: vectorisable dot and libm exp, for timing the CPU baseline only (apply_targets) */
    const int fast = cfg->fast_math != 0;
    for (uint32_t q = 0; q < m; ++q) {
        fp[q] = fast ? fast_dot(h, t1 + (uint64_t)ctx[q] * stride, stride)
                     : orc_dot(h, t1 + (uint64_t)ctx[q] * stride, stride);
        if (cfg->scale_by_sqrt_dim) fp[q] = fp[q] * inv_scale;
    }
    for (uint32_t k = 0; k < K; ++k) {
        if (!valid[k]) continue;
        fn[k] = fast ? fast_dot(h, t1 + (uint64_t)neg[k] * stride, stride)
                     : orc_dot(h, t1 + (uint64_t)neg[k] * stride, stride);
        if (cfg->scale_by_sqrt_dim) fn[k] = fn[k] * inv_scale;
    }
    const float fm = (float)m;
    float *acc_p = acc, *acc_n = acc + stride;
    memset(acc, 0, 2 * stride * sizeof(float));
    for (uint32_t k = 0; k < K; ++k) {
        if (!valid[k]) continue;
        ++out->targets;
        if (fabsf(fn[k]) > cfg->clipping_value) continue;
        const float g = ((0.0f - (fast ? 1.0f / (1.0f + expf(-fn[k])) : orc_sigmoid(fn[k]))) * lr) * fm;
        out->loss += (double)m * softplus((double)fn[k]);
        float *row = t1 + (uint64_t)neg[k] * stride;
        for (uint32_t e = 0; e < stride; ++e) {
            const float old = row[e];
            acc_n[e] = fmaf(g, old, acc_n[e]);
            row[e] = fmaf(g, h[e], old);
        }
    }
    for (uint32_t q = 0; q < m; ++q) { /* every sum over the rows as they were */
        ++out->targets;
        gp[q] = 0.0f;
        if (fabsf(fp[q]) > cfg->clipping_value) continue;
        gp[q] = (1.0f - (fast ? 1.0f / (1.0f + expf(-fp[q])) : orc_sigmoid(fp[q]))) * lr;
        out->loss += softplus(-(double)fp[q]);
        const float *row = t1 + (uint64_t)ctx[q] * stride;
        for (uint32_t e = 0; e < stride; ++e) acc_p[e] = fmaf(gp[q], row[e], acc_p[e]);
    }
    for (uint32_t q = 0; q < m; ++q) {
        if (fabsf(fp[q]) > cfg->clipping_value) continue;
        float *row = t1 + (uint64_t)ctx[q] * stride;
        for (uint32_t e = 0; e < stride; ++e) row[e] = row[e] + gp[q] * h[e];
    }
    for (uint32_t e = 0; e < stride; ++e) crow[e] = h[e] + (acc_p[e] + acc_n[e]);
    out->pairs += m;
}

static void train_one_walk(const orc_sgns_cfg *cfg, const uint32_t *walk, uint64_t wid,
                           uint64_t seed, uint64_t n, const int64_t *indptr, const uint32_t *thr,
                           const uint32_t *alias, float *t0, float *t1, float *h, float *acc,
                           train_acc *out) {
    const uint32_t L = cfg->walk_length, w = cfg->window_size, K = cfg->negatives;
    const uint32_t stride = cfg->row_stride;
    const float inv_scale = 1.0f / sqrtf((float)cfg->embedding_size);
    uint32_t target[MAX_TARGETS];
    int valid[MAX_TARGETS];
    for (uint32_t i = 0; i < L; ++i) {
        const uint32_t c = walk[i];
        if (c == ORC_PAD_TOKEN) break;
        /* stochastic_downsample_by_degree (.../node2vec_skipgram.py:97-98): the centre is skipped
         * with probability deg(c) / (max degree + 1); it still serves as context of its neighbours */
        if (cfg->downsample_bound) {
            uint32_t rnd[4];
            orc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)wid,
                              (uint32_t)(wid >> 32), i, ORC_TAG_SKIP << 24, rnd);
            if (orc_mulhi(rnd[0], cfg->downsample_bound) < (uint64_t)(indptr[c + 1] - indptr[c]))
                continue;
        }
        float lr = cfg->learning_rate;
        if (cfg->normalize_learning_rate_by_degree) { /* a sink (directed graphs) counts as degree 1 */
            const uint64_t deg = (uint64_t)(indptr[c + 1] - indptr[c]);
            lr = lr / (float)(deg ? deg : 1);
        }
        const uint32_t lo = i > w ? i - w : 0;
        const uint32_t hi = i + w < L - 1 ? i + w : L - 1;
        if (cfg->model == 0 && cfg->shared_negatives) {
            uint32_t ctx[2 * 64];
            uint32_t m = 0;
            for (uint32_t j = lo; j <= hi; ++j) {
                const uint32_t o = walk[j];
                if (j == i || o == ORC_PAD_TOKEN || o == c) continue;
                ctx[m++] = o;
            }
            if (m) train_centre_shared(cfg, lr, inv_scale, c, ctx, m, seed, wid, i, n, thr, alias, t0, t1, h, acc, out);
        } else if (cfg->model == 0) {
            float *crow = t0 + (uint64_t)c * stride;
            memcpy(h, crow, stride * sizeof(float));
            for (uint32_t j = lo; j <= hi; ++j) {
                const uint32_t o = walk[j];
                if (j == i || o == ORC_PAD_TOKEN || o == c) continue;
                target[0] = o;
                valid[0] = 1;
                draw_negatives(cfg, seed, wid, (i << 16) | j, n, thr, alias, c, o, target + 1,
                               valid + 1);
                memset(acc, 0, stride * sizeof(float));
                apply_targets(cfg, lr, inv_scale, h, t1, target, valid, K + 1, acc, out);
                for (uint32_t e = 0; e < stride; ++e) h[e] = h[e] + acc[e];
                ++out->pairs;
            }
            memcpy(crow, h, stride * sizeof(float));
        } else {
            uint32_t ctx[2 * 64];
            uint32_t m = 0;
            for (uint32_t j = lo; j <= hi; ++j) {
                const uint32_t o = walk[j];
                if (j == i || o == ORC_PAD_TOKEN || o == c) continue;
                ctx[m++] = o;
            }
            if (m == 0) continue;
            memcpy(h, t0 + (uint64_t)ctx[0] * stride, stride * sizeof(float));
            for (uint32_t q = 1; q < m; ++q) {
                const float *row = t0 + (uint64_t)ctx[q] * stride;
                for (uint32_t e = 0; e < stride; ++e) h[e] = h[e] + row[e];
            }
            const float fm = (float)m;
            for (uint32_t e = 0; e < stride; ++e) h[e] = h[e] / fm;
            target[0] = c;
            valid[0] = 1;
            draw_negatives(cfg, seed, wid, (i << 16) | 0xFFFFu, n, thr, alias, c, c, target + 1,
                           valid + 1);
            memset(acc, 0, stride * sizeof(float));
            apply_targets(cfg, lr, inv_scale, h, t1, target, valid, K + 1, acc, out);
            for (uint32_t q = 0; q < m; ++q) {
                float *row = t0 + (uint64_t)ctx[q] * stride;
                for (uint32_t e = 0; e < stride; ++e) row[e] = row[e] + acc[e];
            }
            out->pairs += m;
        }
    }
}

int orc_train(const orc_sgns_cfg *cfg, const uint32_t *walks, uint64_t n_walks,
              uint64_t first_walk, uint64_t walk_id_stride, uint64_t seed, uint64_t n,
              const int64_t *indptr, const uint32_t *thr, const uint32_t *alias, float *t0,
              float *t1, double *loss_sum, uint64_t *pairs, uint64_t *targets) {
    if (!cfg || !walks || !t0 || !t1) return -1;
    if (cfg->negatives + 1 > MAX_TARGETS || cfg->window_size > 64 || cfg->walk_length > 65535 ||
        (cfg->row_stride & 3) || cfg->row_stride < cfg->embedding_size || n > 0xFFFFFFFFull)
        return -1;
    if (cfg->use_alias && (!thr || !alias)) return -1;
    if (cfg->shared_negatives && cfg->model != 0) return -1;
    if ((cfg->normalize_learning_rate_by_degree || cfg->downsample_bound) && !indptr) return -1;
    train_acc total = {0.0, 0, 0};
    int failed = 0;
#pragma omp parallel num_threads(g_threads) if (g_threads > 1)
    {
        float *h = (float *)malloc(cfg->row_stride * sizeof(float));
        float *acc = (float *)malloc(2 * cfg->row_stride * sizeof(float)); /* two sums in the shared mode */
        train_acc local = {0.0, 0, 0};
        if (!h || !acc) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 64)
            for (uint64_t i = 0; i < n_walks; ++i)
                train_one_walk(cfg, walks + i * (uint64_t)cfg->walk_length,
                               first_walk + i * walk_id_stride, seed, n, indptr, thr, alias, t0,
                               t1, h, acc, &local);
        }
#pragma omp critical
        {
            total.loss += local.loss;
            total.pairs += local.pairs;
            total.targets += local.targets;
        }
        free(h);
        free(acc);
    }
    if (failed) return -2;
    if (loss_sum) *loss_sum += total.loss;
    if (pairs) *pairs += total.pairs;
    if (targets) *targets += total.targets;
    return 0;
}
