// K2: DeepWalk / node2vec walk kernels for sm_100a.
//
// Replaces the walk half of `ensmallen.models.SkipGram/CBOW.fit_transform`
// (/root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py:99; walk kwargs
// .../node2vec_skipgram.py:51-81).  One walk per thread: a walk is a chain of dependent
// random 32 B-sector gathers (offsets -> neighbour), so throughput comes from the number of
// independent chains in flight per SM, not from wide loads.  Second order uses KnightKing
// rejection sampling against a uniform neighbour proposal with an integer accept test, so the
// result is bit-identical to the CPU oracle driven by the same Philox stream.
//
// Tokens are buffered four at a time in registers and written with one 16 B store: two
// consecutive stores fill a 32 B sector while the line is still resident in L2.
#include "common.cuh"

namespace b2e {

__device__ __forceinline__ bool row_contains(const uint32_t *__restrict__ row, uint32_t len,
                                             uint32_t key) {
    uint32_t lo = 0, hi = len;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(row + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo < len && __ldg(row + lo) == key;
}

// index of a proposal inside a row: uniform, or proportional to the edge weights through the
// per-edge table built at load time (first entry whose cdf exceeds the random word)
template <bool WEIGHTED>
__device__ __forceinline__ uint32_t propose(const uint32_t *__restrict__ cdf, int64_t off, uint32_t deg,
                                            uint32_t r) {
    if (!WEIGHTED) return __umulhi(r, deg);
    const uint32_t *row = cdf + off;
    uint32_t lo = 0, hi = deg;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(row + mid) > r) hi = mid; else lo = mid + 1;
    }
    return lo < deg ? lo : deg - 1;
}

template <bool SECOND, bool VEC, bool WEIGHTED>
__global__ void __launch_bounds__(256) walk_kernel(const WalkParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_steps = 0, n_trials = 0, n_searches = 0;
    if (i < p.n_walks) {
        const uint64_t wid = p.first_walk + i * p.walk_id_stride;
        const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
        uint32_t *out = p.out + i * (uint64_t)p.walk_length;
        const unsigned long long thr_lo = min(p.thr_common, p.thr_explore);
        const unsigned long long thr_hi = max(p.thr_common, p.thr_explore);

        uint32_t cur = __ldg(p.sources + (wid % p.n_src));
        int64_t prev_off = 0;
        uint32_t prev = PAD, prev_deg = 0;
        bool alive = true;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        uint32_t tok[4];
        const uint32_t L = p.walk_length;
        for (uint32_t base = 0; base < L; base += 4) {
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t t = base + u;
                if (t == 0) { tok[0] = cur; continue; }
                if (t >= L) { tok[u] = PAD; continue; }
                uint32_t next = PAD;
                if (alive) {
                    const int64_t off = __ldg(p.indptr + cur);
                    const uint32_t deg = (uint32_t)(__ldg(p.indptr + cur + 1) - off);
                    if (deg == 0) {
                        alive = false;
                    } else {
                        if (!SECOND || t == 1) {
                            const uint32_t s = t - 1;
                            if ((s & 3u) == 0 || (SECOND && t == 1))
                                rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, s >> 2,
                                                    TAG_WALK1 << 24);
                            const uint32_t r = (s & 3u) == 0 ? rnd.x : (s & 3u) == 1 ? rnd.y
                                             : (s & 3u) == 2 ? rnd.z : rnd.w;
                            next = __ldg(p.indices + off + propose<WEIGHTED>(p.cdf, off, deg, r));
                        } else {
                            const uint32_t *prow = p.indices + prev_off;
                            uint32_t trial = 0;
                            for (;;) {
                                if ((trial & 1u) == 0)
                                    rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, t - 1,
                                                        (TAG_WALK2 << 24) | (trial >> 1));
                                const uint32_t r0 = (trial & 1u) ? rnd.z : rnd.x;
                                const unsigned long long r1 = (trial & 1u) ? rnd.w : rnd.y;
                                next = __ldg(p.indices + off + propose<WEIGHTED>(p.cdf, off, deg, r0));
                                ++n_trials;
                                bool accept;
                                if (next == prev) {
                                    accept = r1 < p.thr_return;
                                } else if (r1 < thr_lo) {
                                    accept = true;   // every non-return class accepts
                                } else if (r1 >= thr_hi) {
                                    accept = false;  // every non-return class rejects
                                } else {
                                    ++n_searches;
                                    const bool common = row_contains(prow, prev_deg, next);
                                    accept = r1 < (common ? p.thr_common : p.thr_explore);
                                }
                                if (accept) break;
                                ++trial;
                                if (trial >= MAX_TRIALS) break;
                            }
                        }
                        ++n_steps;
                        prev = cur;
                        prev_off = off;
                        prev_deg = deg;
                        cur = next;
                    }
                }
                tok[u] = next;
            }
            if (VEC) {
                *reinterpret_cast<uint4 *>(out + base) = make_uint4(tok[0], tok[1], tok[2], tok[3]);
            } else {
#pragma unroll
                for (uint32_t u = 0; u < 4; ++u)
                    if (base + u < L) out[base + u] = tok[u];
            }
        }
    }
    // one atomic per warp and counter
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        n_steps += __shfl_xor_sync(0xffffffffu, n_steps, off);
        n_trials += __shfl_xor_sync(0xffffffffu, n_trials, off);
        n_searches += __shfl_xor_sync(0xffffffffu, n_searches, off);
    }
    if ((threadIdx.x & 31) == 0 && p.counters) {
        atomicAdd(&p.counters->walk_steps, n_steps);
        if (SECOND) {
            atomicAdd(&p.counters->walk_trials, n_trials);
            atomicAdd(&p.counters->walk_searches, n_searches);
        }
    }
}

cudaError_t launch_walks(const WalkParams &p, bool second_order, cudaStream_t stream) {
    if (p.n_walks == 0) return cudaSuccess;
    const unsigned block = 256;
    const unsigned grid = (unsigned)((p.n_walks + block - 1) / block);
    const bool vec = (p.walk_length % 4u) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16u) == 0;
    const bool weighted = p.cdf != nullptr;
#define B2E_LAUNCH_WALK(S, V, W) walk_kernel<S, V, W><<<grid, block, 0, stream>>>(p)
    if (second_order) {
        if (vec) { if (weighted) B2E_LAUNCH_WALK(true, true, true); else B2E_LAUNCH_WALK(true, true, false); }
        else { if (weighted) B2E_LAUNCH_WALK(true, false, true); else B2E_LAUNCH_WALK(true, false, false); }
    } else {
        if (vec) { if (weighted) B2E_LAUNCH_WALK(false, true, true); else B2E_LAUNCH_WALK(false, true, false); }
        else { if (weighted) B2E_LAUNCH_WALK(false, false, true); else B2E_LAUNCH_WALK(false, false, false); }
    }
#undef B2E_LAUNCH_WALK
    return cudaGetLastError();
}

}  // namespace b2e
