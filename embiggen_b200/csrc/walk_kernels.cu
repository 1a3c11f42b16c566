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
#include <algorithm>

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
// row's Vose alias table built at load time -- O(1): one 8-byte gather; the high word of r * deg
// picks the slot, its low word is the coin (oracle/walks.c: propose)
template <bool WEIGHTED>
__device__ __forceinline__ uint32_t propose(const uint2 *__restrict__ table, int64_t off, uint32_t deg,
                                            uint32_t r) {
    const unsigned long long u = (unsigned long long)r * deg;
    const uint32_t i = (uint32_t)(u >> 32);
    if (!WEIGHTED) return i;
    const uint2 e = __ldg(table + off + i);
    return (uint32_t)u < e.x ? i : e.y;
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
                            next = __ldg(p.indices + off + propose<WEIGHTED>(p.edge_alias, off, deg, r));
                        } else {
                            const uint32_t *prow = p.indices + prev_off;
                            uint32_t trial = 0;
                            for (;;) {
                                if ((trial & 1u) == 0)
                                    rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, t - 1,
                                                        (TAG_WALK2 << 24) | (trial >> 1));
                                const uint32_t r0 = (trial & 1u) ? rnd.z : rnd.x;
                                const unsigned long long r1 = (trial & 1u) ? rnd.w : rnd.y;
                                next = __ldg(p.indices + off + propose<WEIGHTED>(p.edge_alias, off, deg, r0));
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


// ---- second-order walks as a per-lane state machine ----
//
// In the plain kernel a warp advances at the pace of its slowest lane: every step waits for
// the lane with the most rejected trials, every trial for the lane with the longest adjacency
// search.  Here every lane carries its own walk through the states below and performs exactly
// ONE dependent gather per loop iteration, so no lane ever idles on a neighbour's retry and the
// loads of all 32 lanes are issued by the same instruction (maximum memory-level parallelism).
// Finished lanes fetch the next walk (grid-stride), so warps stay full until the chunk ends.
//
//   SRC   : start node of the walk                      (sources[wid mod n_src])
//   ROW   : row bounds of the current node              (indptr[cur], indptr[cur + 1])
//   TRIAL : one proposal                                (indices[off + idx]) -> accept / reject /
//           needs an adjacency check
//   XROW  : (undirected graphs) row bounds of the proposal, to search the shorter of
//           N(prev) and N(x): x in N(prev) <=> prev in N(x); reused if x is accepted
//   SEARCH: one bisection step of the adjacency check
//
// Decisions are the oracle's (same Philox words, same integer thresholds), only their schedule
// differs, so the walks stay bit-identical.
enum WalkState : uint32_t { W_SRC = 0, W_ROW = 1, W_TRIAL = 2, W_XROW = 3, W_SEARCH = 4, W_DONE = 5 };

template <bool UNDIRECTED, bool VEC>
__global__ void __launch_bounds__(256) walk_sm_kernel(const WalkParams p) {
    const uint64_t threads = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = p.walk_length;
    const unsigned long long thr_lo = min(p.thr_common, p.thr_explore);
    const unsigned long long thr_hi = max(p.thr_common, p.thr_explore);
    unsigned long long n_steps = 0, n_trials = 0, n_searches = 0;

    uint32_t state = i < p.n_walks ? W_SRC : W_DONE;
    uint64_t wid = 0;
    uint32_t t = 0, cur = PAD, prev = PAD, deg = 0, pdeg = 0, trial = 0;
    int64_t off = 0, poff = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    uint32_t x = PAD;                  // proposal under examination
    unsigned long long r1 = 0;         // its accept word
    int64_t xoff = 0;                  // its row, when XROW fetched it
    uint32_t xdeg = 0;
    bool have_xrow = false;
    int64_t sbase = 0;                 // bisection: row base, bounds, key
    uint32_t slo = 0, shi = 0, skey = 0;
    uint32_t tok[4] = {PAD, PAD, PAD, PAD};
    uint32_t *out = nullptr;

    // append one token to the walk; groups of four leave as one 16-byte store
    auto emit = [&](uint32_t token) {
        const uint32_t slot = t & 3u;  // explicit selects keep the group in registers
        if (slot == 0) tok[0] = token; else if (slot == 1) tok[1] = token;
        else if (slot == 2) tok[2] = token; else tok[3] = token;
        ++t;
        const uint32_t filled = t & 3u;
        if (filled == 0) {
            if (VEC) {
                *reinterpret_cast<uint4 *>(out + t - 4u) = make_uint4(tok[0], tok[1], tok[2], tok[3]);
            } else {
                out[t - 4u] = tok[0]; out[t - 3u] = tok[1]; out[t - 2u] = tok[2]; out[t - 1u] = tok[3];
            }
        } else if (t == L) {
            out[t - filled] = tok[0];
            if (filled > 1) out[t - filled + 1u] = tok[1];
            if (filled > 2) out[t - filled + 2u] = tok[2];
        }
    };

    while (__any_sync(0xffffffffu, state != W_DONE)) {
        // ---- phase A: the address of this iteration's gather ----
        const void *addr = p.sources;
        bool wide = false;  // two 8-byte entries of indptr vs one 4-byte token
        switch (state) {
            case W_SRC:
                wid = p.first_walk + i * p.walk_id_stride;
                out = p.out + i * (uint64_t)L;
                addr = p.sources + (wid % p.n_src);
                break;
            case W_ROW:
                addr = p.indptr + cur;
                wide = true;
                break;
            case W_XROW:
                addr = p.indptr + x;
                wide = true;
                break;
            case W_TRIAL: {
                const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
                uint32_t r0;
                if (t == 1) {
                    rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, 0u, TAG_WALK1 << 24);
                    r0 = rnd.x;
                } else {
                    if ((trial & 1u) == 0)
                        rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, t - 1,
                                            (TAG_WALK2 << 24) | (trial >> 1));
                    r0 = (trial & 1u) ? rnd.z : rnd.x;
                    r1 = (trial & 1u) ? rnd.w : rnd.y;
                    ++n_trials;
                }
                addr = p.indices + off + __umulhi(r0, deg);
                break;
            }
            case W_SEARCH:
                addr = p.indices + sbase + (slo + ((shi - slo) >> 1));
                break;
            default:
                break;
        }
        // ---- phase B: one gather per lane, issued by the whole warp at once ----
        uint32_t word = 0;
        long long first = 0, second = 0;
        if (state != W_DONE) {
            if (wide) {
                first = __ldg(reinterpret_cast<const long long *>(addr));
                second = __ldg(reinterpret_cast<const long long *>(addr) + 1);
            } else {
                word = __ldg(reinterpret_cast<const uint32_t *>(addr));
            }
        }
        // ---- phase C: consume it ----
        bool accept = false, reject = false, new_row = false;
        switch (state) {
            case W_SRC:
                cur = word;
                prev = PAD;
                t = 0;
                emit(cur);
                have_xrow = false;
                state = L > 1 ? W_ROW : W_DONE;
                break;
            case W_ROW:
                off = first;
                deg = (uint32_t)(second - first);
                new_row = true;
                break;
            case W_TRIAL:
                x = word;
                if (t == 1) {
                    accept = true;
                } else if (x == prev) {
                    if (r1 < p.thr_return) accept = true; else reject = true;
                } else if (r1 < thr_lo) {
                    accept = true;
                } else if (r1 >= thr_hi) {
                    reject = true;
                } else {
                    ++n_searches;
                    if (UNDIRECTED) {
                        state = W_XROW;
                    } else {
                        sbase = poff; slo = 0; shi = pdeg; skey = x;
                        state = W_SEARCH;
                    }
                }
                break;
            case W_XROW:
                xoff = first;
                xdeg = (uint32_t)(second - first);
                have_xrow = true;
                // x in N(prev) <=> prev in N(x) on an undirected graph: bisect the shorter row
                if (xdeg < pdeg) { sbase = xoff; slo = 0; shi = xdeg; skey = prev; }
                else { sbase = poff; slo = 0; shi = pdeg; skey = x; }
                state = W_SEARCH;
                break;
            case W_SEARCH: {
                const uint32_t mid = slo + ((shi - slo) >> 1);
                bool decided = false, common = false;
                if (word == skey) { decided = true; common = true; }
                else if (word < skey) slo = mid + 1;
                else shi = mid;
                if (!decided && slo >= shi) decided = true;
                if (decided) {
                    if (r1 < (common ? p.thr_common : p.thr_explore)) accept = true; else reject = true;
                }
                break;
            }
            default:
                break;
        }
        if (reject) {
            ++trial;
            have_xrow = false;
            if (trial >= MAX_TRIALS) accept = true; else state = W_TRIAL;
        }
        if (accept) {
            ++n_steps;
            prev = cur; poff = off; pdeg = deg;
            cur = x;
            emit(x);
            if (t >= L) {
                state = W_DONE;
            } else if (have_xrow) {
                off = xoff; deg = xdeg;
                new_row = true;
            } else {
                state = W_ROW;
            }
            have_xrow = false;
        }
        if (new_row) {  // the row of the current node is known: walk on, or pad after a dead end
            trial = 0;
            state = W_TRIAL;
            if (deg == 0) {
                while (t < L) emit(PAD);
                state = W_DONE;
            }
        }
        if (state == W_DONE && i < p.n_walks) {  // this lane's walk is complete: fetch the next one
            i += threads;
            if (i < p.n_walks) state = W_SRC; else i = p.n_walks;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        n_steps += __shfl_xor_sync(0xffffffffu, n_steps, o);
        n_trials += __shfl_xor_sync(0xffffffffu, n_trials, o);
        n_searches += __shfl_xor_sync(0xffffffffu, n_searches, o);
    }
    if ((threadIdx.x & 31) == 0 && p.counters) {
        atomicAdd(&p.counters->walk_steps, n_steps);
        atomicAdd(&p.counters->walk_trials, n_trials);
        atomicAdd(&p.counters->walk_searches, n_searches);
    }
}


// Typed walks (change_node_type_weight / change_edge_type_weight, .../node2vec_skipgram.py:72-77):
// every transition is a trial loop with ONE Philox block per trial (tag 7): x proposal, z
// node-type test, w edge-type test, y p/q test.  Independent words => the acceptance probability
// is the product of the three ratios.  Cheap tests first; see oracle/walks.c:walks_general.
// (normalize_by_degree needs no kernel: it is folded into the proposal table at load.)
template <bool VEC, bool WEIGHTED>
__global__ void __launch_bounds__(256) walk_general_kernel(const WalkParams p) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_steps = 0, n_trials = 0, n_searches = 0;
    if (i < p.n_walks) {
        const uint64_t wid = p.first_walk + i * p.walk_id_stride;
        const uint32_t wid_lo = (uint32_t)wid, wid_hi = (uint32_t)(wid >> 32);
        uint32_t *out = p.out + i * (uint64_t)p.walk_length;
        const unsigned long long thr_lo = min(p.thr_common, p.thr_explore);
        const unsigned long long thr_hi = max(p.thr_common, p.thr_explore);
        const bool use_nt = p.node_types != nullptr && p.q_node[0] != p.q_node[1];
        const bool use_et = p.edge_types != nullptr && p.q_edge[0] != p.q_edge[1];
        uint32_t cur = __ldg(p.sources + (wid % p.n_src));
        int64_t prev_off = 0;
        uint32_t prev = PAD, prev_deg = 0, prev_etype = 0;
        bool alive = true;
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
                        const uint32_t cur_type = use_nt ? __ldg(p.node_types + cur) : 0u;
                        uint32_t trial = 0;
                        int64_t e = off;
                        for (;;) {
                            const uint4 rnd = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi, t - 1,
                                                            (TAG_WALK3 << 24) | trial);
                            e = off + propose<WEIGHTED>(p.edge_alias, off, deg, rnd.x);
                            next = __ldg(p.indices + e);
                            ++n_trials;
                            bool accept = true;
                            if (use_nt)
                                accept = rnd.z < p.q_node[__ldg(p.node_types + next) != cur_type ? 1 : 0];
                            if (accept && use_et && t > 1)
                                accept = rnd.w < p.q_edge[__ldg(p.edge_types + e) != prev_etype ? 1 : 0];
                            if (accept && t > 1) {
                                const unsigned long long lhs = rnd.y;
                                if (next == prev) {
                                    accept = lhs < p.thr_return;
                                } else if (lhs < thr_lo) {
                                    accept = true;
                                } else if (lhs >= thr_hi) {
                                    accept = false;
                                } else {
                                    ++n_searches;
                                    const bool common = row_contains(p.indices + prev_off, prev_deg, next);
                                    accept = lhs < (common ? p.thr_common : p.thr_explore);
                                }
                            }
                            if (accept) break;
                            ++trial;
                            if (trial >= MAX_TRIALS) break;
                        }
                        ++n_steps;
                        if (p.edge_types) prev_etype = __ldg(p.edge_types + e);
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
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        n_steps += __shfl_xor_sync(0xffffffffu, n_steps, off);
        n_trials += __shfl_xor_sync(0xffffffffu, n_trials, off);
        n_searches += __shfl_xor_sync(0xffffffffu, n_searches, off);
    }
    if ((threadIdx.x & 31) == 0 && p.counters) {
        atomicAdd(&p.counters->walk_steps, n_steps);
        atomicAdd(&p.counters->walk_trials, n_trials);
        atomicAdd(&p.counters->walk_searches, n_searches);
    }
}

// Walklets (.../walklets.py:7-149): scale k keeps the pairs exactly k hops apart, i.e. adjacent
// tokens of the k sub-walks made of every k-th token.  out[r][w][m] = raw[w][r + m k] (PAD past
// the end), sub-walk length ceil(L / k); the SGD kernels then run unchanged on the sub-walks.
__global__ void __launch_bounds__(256) walklet_split_kernel(const uint32_t *__restrict__ raw, uint64_t n_walks,
                                                            uint32_t L, uint32_t k, uint32_t Ls,
                                                            uint32_t *__restrict__ out) {
    const uint64_t total = n_walks * k * Ls;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t m = (uint32_t)(idx % Ls);
        const uint64_t rest = idx / Ls;
        const uint64_t w = rest % n_walks;
        const uint32_t r = (uint32_t)(rest / n_walks);
        const uint32_t t = r + m * k;
        out[idx] = t < L ? __ldg(raw + w * L + t) : PAD;
    }
}

cudaError_t launch_walklet_split(const uint32_t *raw, uint64_t n_walks, uint32_t walk_length, uint32_t scale,
                                 uint32_t *out, cudaStream_t stream) {
    const uint32_t Ls = (walk_length + scale - 1) / scale;
    const uint64_t total = n_walks * scale * Ls;
    if (total == 0) return cudaSuccess;
    const uint64_t grid = std::min<uint64_t>((total + 255) / 256, 148ull * 16);
    walklet_split_kernel<<<(unsigned)grid, 256, 0, stream>>>(raw, n_walks, walk_length, scale, Ls, out);
    return cudaGetLastError();
}

// One thread per directed edge (u, v): is u in the row of v?  Clears *symmetric otherwise.
__global__ void __launch_bounds__(256) symmetry_kernel(const int64_t *__restrict__ indptr,
                                                       const uint32_t *__restrict__ indices, uint64_t n,
                                                       uint64_t nnz, int *symmetric) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz || *symmetric == 0) return;
    uint64_t lo = 0, hi = n;  // row of edge e: last u with indptr[u] <= e
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo + 1) >> 1);
        if ((uint64_t)__ldg(indptr + mid) <= e) lo = mid; else hi = mid - 1;
    }
    const uint32_t u = (uint32_t)lo, v = __ldg(indices + e);
    const int64_t begin = __ldg(indptr + v);
    const uint32_t len = (uint32_t)(__ldg(indptr + v + 1) - begin);
    if (!row_contains(indices + begin, len, u)) *symmetric = 0;
}

cudaError_t launch_symmetry_check(const int64_t *indptr, const uint32_t *indices, uint64_t n,
                                  uint64_t nnz, int *d_flag, cudaStream_t stream) {
    if (nnz == 0) return cudaSuccess;
    symmetry_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(indptr, indices, n, nnz, d_flag);
    return cudaGetLastError();
}

cudaError_t launch_walks(const WalkParams &p, bool second_order, cudaStream_t stream) {
    if (p.n_walks == 0) return cudaSuccess;
    const unsigned block = 256;
    const unsigned grid = (unsigned)((p.n_walks + block - 1) / block);
    const bool vec = (p.walk_length % 4u) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16u) == 0;
    const bool weighted = p.edge_alias != nullptr;
    const bool typed = (p.node_types && p.q_node[0] != p.q_node[1]) ||
                       (p.edge_types && p.q_edge[0] != p.q_edge[1]);
    if (typed) {  // typed walks: one trial loop per transition
        if (vec) { if (weighted) walk_general_kernel<true, true><<<grid, block, 0, stream>>>(p);
                   else walk_general_kernel<true, false><<<grid, block, 0, stream>>>(p); }
        else { if (weighted) walk_general_kernel<false, true><<<grid, block, 0, stream>>>(p);
               else walk_general_kernel<false, false><<<grid, block, 0, stream>>>(p); }
        return cudaGetLastError();
    }
    if (second_order && !weighted && p.state_machine) {
        // persistent grid: lanes fetch walks grid-stride, so size it to the machine, not the chunk
        int per_sm = 0;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, p.undirected ? (vec ? walk_sm_kernel<true, true> : walk_sm_kernel<true, false>)
                                  : (vec ? walk_sm_kernel<false, true> : walk_sm_kernel<false, false>),
            (int)block, 0);
        if (err != cudaSuccess) return err;
        unsigned resident = (unsigned)std::max(1, per_sm) * (unsigned)p.sm_count;
        const unsigned sm_grid = std::min(grid, resident);
        if (p.undirected) {
            if (vec) walk_sm_kernel<true, true><<<sm_grid, block, 0, stream>>>(p);
            else walk_sm_kernel<true, false><<<sm_grid, block, 0, stream>>>(p);
        } else {
            if (vec) walk_sm_kernel<false, true><<<sm_grid, block, 0, stream>>>(p);
            else walk_sm_kernel<false, false><<<sm_grid, block, 0, stream>>>(p);
        }
        return cudaGetLastError();
    }
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
