// Device helpers shared by the SGD kernels (sgns_kernels.cu: register-staged generic path,
// sgns_pipe.cu: shared-memory pipelined path).  Normative floating point: DESIGN.md.
#pragma once
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

// ---- register-staged rows: lane l owns float4 chunks l, l + 32, ... of a row ----
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
    // a sink of a directed graph is the last real token of its walk and still serves as a centre
    const uint32_t deg = (uint32_t)(__ldg(p.indptr + centre + 1) - __ldg(p.indptr + centre));
    return __fdiv_rn(p.lr, (float)(deg ? deg : 1u));
}

// bit i of the mask staged behind a shared-memory walk of L tokens (rounded up to 32)
__device__ __forceinline__ bool staged_skip(const uint32_t *walk, uint32_t L, uint32_t i) {
    return (walk[((L + 31u) & ~31u) + (i >> 5)] >> (i & 31u)) & 1u;
}

// stochastic_downsample_by_degree: the centre at position i is skipped with probability
// deg(c) / (max degree + 1); one Philox block per centre, the same on every lane.
__device__ __forceinline__ bool skip_centre(const TrainParams &p, uint32_t wid_lo, uint32_t wid_hi,
                                            uint32_t i, uint32_t c) {
    if (!p.downsample) return false;
    const uint4 r = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, i, TAG_SKIP << 24);
    const uint32_t deg = (uint32_t)(__ldg(p.indptr + c + 1) - __ldg(p.indptr + c));
    return __umulhi(r.x, p.downsample) < deg;
}

// ---- SkipGram: the draw sites of a walk are its (centre, context) pairs, in oracle order ----
struct PairCursor {
    uint32_t i, j;  // centre and context positions; start with i = 0xFFFFFFFF
    uint32_t c, o;  // their tokens
    uint32_t hi;    // last window position of centre i
};

// advance to the next pair of the walk; false when the walk is exhausted.  STAGED: `walk` is a
// shared-memory copy of the walk (plain loads) instead of global memory (read-only path).  With
// stochastic_downsample_by_degree the pipelined kernels (STAGED) read the skip decision from a
// bit mask staged behind the walk (sgns_pipe.cu: stage_skip_mask); the generic kernel evaluates
// it in place.
template <bool STAGED = false>
__device__ __forceinline__ bool next_pair(const TrainParams &p, uint32_t wid_lo, uint32_t wid_hi,
                                          const uint32_t *__restrict__ walk, uint32_t L, uint32_t W,
                                          PairCursor &s) {
    for (;;) {
        if (s.i == 0xFFFFFFFFu || s.j >= s.hi) {
            const uint32_t i = s.i + 1u;  // wraps 0xFFFFFFFF -> 0
            if (i >= L) return false;
            const uint32_t c = STAGED ? walk[i] : __ldg(walk + i);
            if (c == PAD) return false;
            s.i = i;
            if (STAGED ? (p.downsample && staged_skip(walk, L, i)) : skip_centre(p, wid_lo, wid_hi, i, c)) {
                s.j = s.hi = 0;
                continue;
            }
            s.c = c;
            s.hi = i + W < L - 1 ? i + W : L - 1;
            s.j = i > W ? i - W : 0u;
        } else {
            ++s.j;
        }
        if (s.j == s.i) continue;
        s.o = STAGED ? walk[s.j] : __ldg(walk + s.j);
        if (s.o == PAD || s.o == s.c) continue;
        return true;
    }
}

// ---- CBOW: the draw sites of a walk are its centres ----
// next centre position >= i with at least one valid context; returns L when exhausted
template <bool STAGED = false>
__device__ __forceinline__ uint32_t next_centre(const TrainParams &p, uint32_t wid_lo, uint32_t wid_hi,
                                                const uint32_t *__restrict__ walk, uint32_t L,
                                                uint32_t W, uint32_t i, uint32_t &c) {
    for (; i < L; ++i) {
        c = STAGED ? walk[i] : __ldg(walk + i);
        if (c == PAD) return L;
        if (STAGED ? (p.downsample && staged_skip(walk, L, i)) : skip_centre(p, wid_lo, wid_hi, i, c)) continue;
        const uint32_t lo = i > W ? i - W : 0u;
        const uint32_t hi = i + W < L - 1 ? i + W : L - 1;
        for (uint32_t j = lo; j <= hi; ++j) {
            if (j == i) continue;
            const uint32_t o = STAGED ? walk[j] : __ldg(walk + j);
            if (o != PAD && o != c) return i;
        }
    }
    return L;
}

}  // namespace b2e
