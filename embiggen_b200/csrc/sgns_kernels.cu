// K4 / K5: fused window + negative-sampling SkipGram and CBOW SGD kernels for sm_100a.
//
// Replaces the training half of `ensmallen.models.SkipGram/CBOW.fit_transform`
// (/root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py:99; kwargs
// .../node2vec_skipgram.py:37-119).  One warp per walk; a warp slides the window over its
// walk, draws the negatives of a pair from the alias table (one lane per negative), gathers
// the K+1 target rows of the contextual table with 128-bit loads issued back to back,
// reduces the dot products with warp shuffles, and scatters the updated rows back
// (Hogwild: plain vector stores, no locks).  The bound is HBM: 2 * 4D bytes per target row.
//
// Floating point follows the normative spec: explicit round-to-nearest
// intrinsics, a fixed reduction order and a polynomial sigmoid, so a single-warp launch
// (cfg.deterministic) reproduces the CPU oracle's tables bit for bit.
#include "common.cuh"

namespace b2e {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float exp_det(float y) {
    y = y > 80.0f ? 80.0f : y;
    y = y < -80.0f ? -80.0f : y;
    const float k = rintf(__fmul_rn(y, 1.44269504088896341f));
    float r = __fmaf_rn(k, -0.693145751953125f, y);
    r = __fmaf_rn(k, -1.42860682030941723212e-6f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    p = __fmaf_rn(p, __fmul_rn(r, r), r);
    p = __fadd_rn(p, 1.0f);
    return __fmul_rn(p, __int_as_float(((int)k + 127) << 23));
}

__device__ __forceinline__ float sigmoid_det(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_det(-x)));
}

__device__ __forceinline__ float softplus(float z) { return z > 15.0f ? z : log1pf(expf(z)); }

template <int CH>
__device__ __forceinline__ void load_row(const float *row, uint32_t chunks, uint32_t lane,
                                         float4 (&r)[CH]) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        const uint32_t c = lane + 32u * ch;
        r[ch] = c < chunks ? *reinterpret_cast<const float4 *>(row + 4u * c)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int CH>
__device__ __forceinline__ void store_row(float *row, uint32_t chunks, uint32_t lane,
                                          const float4 (&r)[CH]) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        const uint32_t c = lane + 32u * ch;
        if (c < chunks) *reinterpret_cast<float4 *>(row + 4u * c) = r[ch];
    }
}

// lane l owns float4 chunks l, l+32, ...; xor-butterfly 16, 8, 4, 2, 1
template <int CH>
__device__ __forceinline__ float warp_dot(const float4 (&a)[CH], const float4 (&b)[CH]) {
    float p = 0.0f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        p = __fmaf_rn(a[ch].x, b[ch].x, p);
        p = __fmaf_rn(a[ch].y, b[ch].y, p);
        p = __fmaf_rn(a[ch].z, b[ch].z, p);
        p = __fmaf_rn(a[ch].w, b[ch].w, p);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(FULL, p, off));
    return p;
}

// lane k < K draws negative k of the site; returns the validity of that lane's draw
__device__ __forceinline__ bool draw_negative(const TrainParams &p, uint32_t wid_lo,
                                              uint32_t wid_hi, uint32_t site, uint32_t lane,
                                              uint32_t centre, uint32_t context, uint32_t &neg) {
    bool valid = false;
    neg = PAD;
    if (lane < p.negatives) {
        const uint4 r = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, site,
                                      (TAG_NEG << 24) | lane);
        const uint32_t idx = __umulhi(r.x, p.n);
        neg = idx;
        if (p.use_alias) {
            const uint2 e = __ldg(p.alias + idx);
            neg = r.y < e.x ? idx : e.y;
        }
        valid = neg != centre && neg != context;
    }
    // a draw equal to an earlier draw of the same site is dropped
    for (uint32_t q = 0; q + 1 < p.negatives; ++q) {
        const uint32_t other = __shfl_sync(FULL, neg, q);
        if (lane > q && other == neg) valid = false;
    }
    return valid;
}

// Score h against slot 0 (= positive) and the valid negatives, in batches of NT rows whose
// loads are all issued before the first use; acc += sum g * row (pre-update rows).
template <int CH, int NT>
__device__ __forceinline__ void apply_targets(const TrainParams &p, uint32_t chunks, uint32_t lane,
                                              const float4 (&h)[CH], float lr, uint32_t positive,
                                              uint32_t neg, uint32_t vmask, float4 (&acc)[CH],
                                              float &loss_acc, unsigned long long &n_targets) {
    const uint32_t K = p.negatives;
    for (uint32_t b = 0; b <= K; b += NT) {
        float4 rows[NT][CH];
        uint32_t ids[NT];
#pragma unroll
        for (int s = 0; s < NT; ++s) {
            const uint32_t slot = b + s;
            const uint32_t id = slot == 0 ? positive : __shfl_sync(FULL, neg, (slot - 1) & 31u);
            const bool on = slot <= K && ((vmask >> slot) & 1u);
            ids[s] = on ? id : PAD;
            if (on) load_row<CH>(p.t1 + (uint64_t)id * p.row_stride, chunks, lane, rows[s]);
        }
        float my_f = 0.0f;
        bool my_on = false;
#pragma unroll
        for (int s = 0; s < NT; ++s) {
            if (ids[s] != PAD) {
                float f = warp_dot<CH>(h, rows[s]);
                if (p.scale_dot) f = __fmul_rn(f, p.inv_scale);
                if (lane == (uint32_t)s) { my_f = f; my_on = true; }
            }
        }
        float g_mine = 0.0f;
        bool apply = false;
        if (my_on && !(fabsf(my_f) > p.clip)) {
            const bool is_positive = (b + lane) == 0;
            g_mine = __fmul_rn(__fsub_rn(is_positive ? 1.0f : 0.0f, sigmoid_det(my_f)), lr);
            loss_acc += softplus(is_positive ? -my_f : my_f);
            apply = true;
        }
        const uint32_t amask = __ballot_sync(FULL, apply);
        n_targets += __popc(__ballot_sync(FULL, my_on));
#pragma unroll
        for (int s = 0; s < NT; ++s) {
            if ((amask >> s) & 1u) {
                const float g = __shfl_sync(FULL, g_mine, s);
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) {
                    const float4 old = rows[s][ch];
                    acc[ch].x = __fmaf_rn(g, old.x, acc[ch].x);
                    acc[ch].y = __fmaf_rn(g, old.y, acc[ch].y);
                    acc[ch].z = __fmaf_rn(g, old.z, acc[ch].z);
                    acc[ch].w = __fmaf_rn(g, old.w, acc[ch].w);
                    rows[s][ch].x = __fmaf_rn(g, h[ch].x, old.x);
                    rows[s][ch].y = __fmaf_rn(g, h[ch].y, old.y);
                    rows[s][ch].z = __fmaf_rn(g, h[ch].z, old.z);
                    rows[s][ch].w = __fmaf_rn(g, h[ch].w, old.w);
                }
                store_row<CH>(p.t1 + (uint64_t)ids[s] * p.row_stride, chunks, lane, rows[s]);
            }
        }
    }
}

template <int CH>
__device__ __forceinline__ void add_rows(float4 (&a)[CH], const float4 (&b)[CH]) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        a[ch].x = __fadd_rn(a[ch].x, b[ch].x);
        a[ch].y = __fadd_rn(a[ch].y, b[ch].y);
        a[ch].z = __fadd_rn(a[ch].z, b[ch].z);
        a[ch].w = __fadd_rn(a[ch].w, b[ch].w);
    }
}

template <int CH>
__device__ __forceinline__ void zero_rows(float4 (&a)[CH]) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) a[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float centre_lr(const TrainParams &p, uint32_t centre) {
    if (!p.normalize_lr) return p.lr;
    const uint32_t deg = (uint32_t)(__ldg(p.indptr + centre + 1) - __ldg(p.indptr + centre));
    return __fdiv_rn(p.lr, (float)deg);
}

template <int MODEL, int CH, int NT>
__global__ void __launch_bounds__(256) train_kernel(const TrainParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t chunks = p.row_stride >> 2;
    const uint32_t L = p.walk_length, W = p.window;
    float loss_acc = 0.0f;
    unsigned long long n_pairs = 0, n_targets = 0;

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(&p.counters->work_counter, 1ull);
        w = __shfl_sync(FULL, w, 0);
        if (w >= p.n_walks) break;
        const uint64_t wid = p.first_walk + w * p.walk_id_stride;
        const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
        const uint32_t *walk = p.walks + w * (uint64_t)L;

        for (uint32_t i = 0; i < L; ++i) {
            const uint32_t c = __ldg(walk + i);
            if (c == PAD) break;
            const float lr = centre_lr(p, c);
            const uint32_t lo = i > W ? i - W : 0u;
            const uint32_t hi = i + W < L - 1 ? i + W : L - 1;
            float4 h[CH], acc[CH];
            if (MODEL == B2E_SKIPGRAM) {
                float *crow = p.t0 + (uint64_t)c * p.row_stride;
                load_row<CH>(crow, chunks, lane, h);
                for (uint32_t j = lo; j <= hi; ++j) {
                    if (j == i) continue;
                    const uint32_t o = __ldg(walk + j);
                    if (o == PAD || o == c) continue;
                    uint32_t neg;
                    const bool valid = draw_negative(p, wid_lo, wid_hi, (i << 16) | j, lane, c, o, neg);
                    const uint32_t vmask = (__ballot_sync(FULL, valid) << 1) | 1u;
                    zero_rows<CH>(acc);
                    apply_targets<CH, NT>(p, chunks, lane, h, lr, o, neg, vmask, acc, loss_acc,
                                          n_targets);
                    add_rows<CH>(h, acc);
                    ++n_pairs;
                }
                store_row<CH>(crow, chunks, lane, h);
            } else {
                uint32_t m = 0;
                float4 row[CH];
                for (uint32_t j = lo; j <= hi; ++j) {
                    if (j == i) continue;
                    const uint32_t o = __ldg(walk + j);
                    if (o == PAD || o == c) continue;
                    if (m == 0) {
                        load_row<CH>(p.t0 + (uint64_t)o * p.row_stride, chunks, lane, h);
                    } else {
                        load_row<CH>(p.t0 + (uint64_t)o * p.row_stride, chunks, lane, row);
                        add_rows<CH>(h, row);
                    }
                    ++m;
                }
                if (m == 0) continue;
                const float fm = (float)m;
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) {
                    h[ch].x = __fdiv_rn(h[ch].x, fm);
                    h[ch].y = __fdiv_rn(h[ch].y, fm);
                    h[ch].z = __fdiv_rn(h[ch].z, fm);
                    h[ch].w = __fdiv_rn(h[ch].w, fm);
                }
                uint32_t neg;
                const bool valid = draw_negative(p, wid_lo, wid_hi, (i << 16) | 0xFFFFu, lane, c, c, neg);
                const uint32_t vmask = (__ballot_sync(FULL, valid) << 1) | 1u;
                zero_rows<CH>(acc);
                apply_targets<CH, NT>(p, chunks, lane, h, lr, c, neg, vmask, acc, loss_acc,
                                      n_targets);
                for (uint32_t j = lo; j <= hi; ++j) {
                    if (j == i) continue;
                    const uint32_t o = __ldg(walk + j);
                    if (o == PAD || o == c) continue;
                    float *orow = p.t0 + (uint64_t)o * p.row_stride;
                    load_row<CH>(orow, chunks, lane, row);
                    add_rows<CH>(row, acc);
                    store_row<CH>(orow, chunks, lane, row);
                }
                n_pairs += m;
            }
        }
    }
    double loss = (double)loss_acc;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) loss += __shfl_xor_sync(FULL, loss, off);
    if (lane == 0) {
        atomicAdd(&p.counters->pairs, n_pairs);
        atomicAdd(&p.counters->targets, n_targets);
        atomicAdd(&p.counters->loss_sum, loss);
    }
}

template <int MODEL, int CH, int NT>
static cudaError_t launch_one(const TrainParams &p, bool deterministic, int sm_count,
                              cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(&p.counters->work_counter, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    if (deterministic) {
        train_kernel<MODEL, CH, NT><<<1, 32, 0, stream>>>(p);
        return cudaGetLastError();
    }
    const int block = 256;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_kernel<MODEL, CH, NT>, block, 0);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    const uint64_t needed = (p.n_walks + 7) / 8;
    if (grid > needed) grid = needed;
    train_kernel<MODEL, CH, NT><<<(unsigned)grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_model(const TrainParams &p, bool deterministic, int sm_count,
                                cudaStream_t stream) {
    const uint32_t chunks = p.row_stride >> 2;
    if (chunks <= 32) {
        if (p.negatives + 1 <= 6) return launch_one<MODEL, 1, 6>(p, deterministic, sm_count, stream);
        return launch_one<MODEL, 1, 11>(p, deterministic, sm_count, stream);
    }
    if (chunks <= 64) return launch_one<MODEL, 2, 6>(p, deterministic, sm_count, stream);
    if (chunks <= 128) return launch_one<MODEL, 4, 3>(p, deterministic, sm_count, stream);
    return cudaErrorInvalidValue;
}

cudaError_t launch_train(const TrainParams &p, uint32_t model, bool deterministic, int sm_count,
                         cudaStream_t stream) {
    if (p.n_walks == 0) return cudaSuccess;
    if (model == B2E_SKIPGRAM) return launch_model<B2E_SKIPGRAM>(p, deterministic, sm_count, stream);
    return launch_model<B2E_CBOW>(p, deterministic, sm_count, stream);
}

// ---- table initialisation: one thread per float4 chunk ----
__global__ void __launch_bounds__(256) init_kernel(float *t0, float *t1, uint64_t n,
                                                   uint32_t embedding_size, uint32_t row_stride,
                                                   uint32_t seed_lo, uint32_t seed_hi) {
    const uint32_t chunks = row_stride >> 2;
    const uint64_t total = n * chunks * 2;
    const float dim = (float)embedding_size;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t per_table = n * chunks;
        const uint32_t table = idx >= per_table;
        const uint64_t local = table ? idx - per_table : idx;
        const uint64_t row = local / chunks;
        const uint32_t ch = (uint32_t)(local - row * chunks);
        const uint4 r = philox4x32_10(seed_lo, seed_hi, (uint32_t)row, (uint32_t)(row >> 32), ch,
                                      (table ? TAG_INIT1 : TAG_INIT0) << 24);
        const uint32_t words[4] = {r.x, r.y, r.z, r.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float u01 = __fmul_rn((float)(words[e] >> 8), 5.9604644775390625e-8f);
            v[e] = (4u * ch + e) < embedding_size ? __fdiv_rn(__fsub_rn(u01, 0.5f), dim) : 0.0f;
        }
        float *dst = (table ? t1 : t0) + row * row_stride + 4u * ch;
        *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

cudaError_t launch_init_tables(float *t0, float *t1, uint64_t n, uint32_t embedding_size,
                               uint32_t row_stride, uint64_t seed, cudaStream_t stream) {
    const uint64_t total = n * (row_stride >> 2) * 2;
    uint64_t grid = (total + 255) / 256;
    if (grid > 148ull * 32) grid = 148ull * 32;
    if (grid == 0) return cudaSuccess;
    init_kernel<<<(unsigned)grid, 256, 0, stream>>>(t0, t1, n, embedding_size, row_stride,
                                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    return cudaGetLastError();
}

}  // namespace b2e
