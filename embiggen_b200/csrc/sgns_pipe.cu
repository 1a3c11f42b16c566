// K4 (production path): SkipGram negative-sampling SGD as a per-warp asynchronous pipeline
// staged through shared memory (sm_100a).
//
// Replaces the training half of `ensmallen.models.SkipGram.fit_transform`
// (/root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py:99; kwargs
// .../node2vec_skipgram.py:37-119).  One warp per walk.  The kernel is HBM-latency/-bandwidth
// bound: each (centre, context) pair gathers and scatters K+1 rows of the contextual table, K
// of them at random.  Every gather is therefore an asynchronous global->shared copy
// (cp.async / LDGSTS, no register target, L2-only) issued one pair ahead into a two-stage
// per-warp ring, and the alias-table gathers of the negative draws run two pairs ahead:
//
//   iteration p :  wait(all copies committed in iteration p-1)
//                  resolve the ids of pair p+1 from its alias entries (shared memory)
//                  issue the row copies of pair p+1  -> stage (p+1)&1        } one commit
//                  Philox draw of pair p+2, issue its alias-entry copies     } group
//                  train pair p out of stage p&1: dots (LDS.128 + transposed warp reduction),
//                  sigmoid, axpy, rows scattered to global with 128-bit stores (Hogwild)
//
// A row of pair p+1 that pair p is about to update (repeated context, a negative equal to a
// neighbour) would be copied stale; such pairs are detected (one MATCH over the two id sets)
// and their copies are issued after pair p's stores instead.  Lane l copies, reads and stores
// chunk l of every row, so no cross-lane shared-memory hazard exists and the deterministic
// single-warp launch reproduces the CPU oracle bit for bit: the additions of the transposed
// reduction are the same additions, in the same order, as the oracle's xor-butterfly.
#include "sgns_device.cuh"

namespace b2e {

constexpr int PIPE_SLOTS = 16;  // K + 1 targets at most

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(const float *smem) {
    return *reinterpret_cast<const float4 *>(smem);
}
__device__ __forceinline__ void stg128(float *gmem, const float4 &v) {
    *reinterpret_cast<float4 *>(gmem) = v;
}

// Sum 16 per-lane partials over the 32 lanes.  Level `off` pairs lane l with l^off exactly like
// an xor-butterfly, but each lane keeps only half of the values it holds, so 16 shuffles do the
// work of 80.  Afterwards lane l holds the total of value (l >> 1) & 15.
__device__ __forceinline__ float reduce16(float (&v)[16], uint32_t lane) {
#pragma unroll
    for (int n = 8, off = 16; n >= 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < n; ++k) {
            const float send = upper ? v[k] : v[k + n];
            const float keep = upper ? v[k + n] : v[k];
            v[k] = __fadd_rn(keep, __shfl_xor_sync(FULL, send, off));
        }
    }
    return __fadd_rn(v[0], __shfl_xor_sync(FULL, v[0], 1));
}

struct PipeSmem {  // per-warp carve-up of dynamic shared memory, two stages of each
    float *rows_base;      // [2][slots + 1][row_stride]: target rows, then the centre row of T0
    uint2 *alias_base;     // [2][32] alias entries of a draw in flight
    uint32_t *ids_base;    // [2][PIPE_SLOTS] target id per slot (sentinel when the slot is off)
    uint32_t stage_floats;
    __device__ __forceinline__ float *rows(uint32_t stage) const { return rows_base + stage * stage_floats; }
    __device__ __forceinline__ uint2 *alias(uint32_t a) const { return alias_base + a * 32u; }
    __device__ __forceinline__ uint32_t *ids(uint32_t stage) const { return ids_base + stage * PIPE_SLOTS; }
};

// ids of the targets of a pair, one per lane: lane 0 = context, lane k+1 = negative k;
// lanes whose slot is off get a sentinel that can never equal a node id or another sentinel
__device__ __forceinline__ uint32_t slot_ids(uint32_t lane, uint32_t context, uint32_t neg,
                                             uint32_t vmask) {
    const uint32_t shifted = __shfl_up_sync(FULL, neg, 1);
    const uint32_t id = lane == 0 ? context : shifted;
    return ((vmask >> lane) & 1u) ? id : (0xFFFFFF00u | lane);
}

__device__ __forceinline__ void issue_rows(const TrainParams &p, const PipeSmem &sm, uint32_t stage,
                                           uint32_t lane, uint32_t chunks, uint32_t my_id,
                                           uint32_t vmask, uint32_t centre_or_pad) {
    if (lane < PIPE_SLOTS) sm.ids(stage)[lane] = my_id;
    __syncwarp();  // ids are read back by every lane when the pair is trained
    float *dst = sm.rows(stage) + 4u * lane;
    const uint32_t K = p.negatives;
#pragma unroll
    for (int s = 0; s < PIPE_SLOTS; ++s) {
        if (s <= (int)K && ((vmask >> s) & 1u)) {
            const uint32_t id = __shfl_sync(FULL, my_id, s);
            if (lane < chunks)
                cp_async16(dst + (uint32_t)s * p.row_stride,
                           p.t1 + (uint64_t)id * p.row_stride + 4u * lane);
        }
    }
    if (centre_or_pad != PAD && lane < chunks)
        cp_async16(dst + (K + 1u) * p.row_stride,
                   p.t0 + (uint64_t)centre_or_pad * p.row_stride + 4u * lane);
}

__global__ void __launch_bounds__(128, 5) skipgram_pipe_kernel(const TrainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t chunks = p.row_stride >> 2;
    const uint32_t K = p.negatives, L = p.walk_length, W = p.window;
    const uint32_t stage_floats = (K + 2u) * p.row_stride;
    const uint32_t warp_bytes = 2u * stage_floats * 4u + 2u * 32u * 8u + 2u * PIPE_SLOTS * 4u;
    PipeSmem sm;
    {
        unsigned char *base = smem_raw + warp * warp_bytes;
        sm.stage_floats = stage_floats;
        sm.rows_base = reinterpret_cast<float *>(base);
        sm.alias_base = reinterpret_cast<uint2 *>(sm.rows_base + 2u * stage_floats);
        sm.ids_base = reinterpret_cast<uint32_t *>(sm.alias_base + 64);
    }
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

        // draw of a site: proposal + second word stay in registers, the alias entry lands in
        // shared memory (slot `a`) through cp.async
        auto draw = [&](const PairCursor &s, uint32_t a, uint32_t &idx, uint32_t &ry) {
            idx = PAD;
            ry = 0;
            if (lane < K) {
                const uint4 r = philox4x32_10(p.seed_lo, p.seed_hi, wid_lo, wid_hi,
                                              (s.i << 16) | s.j, (TAG_NEG << 24) | lane);
                idx = __umulhi(r.x, p.n);
                ry = r.y;
                if (p.use_alias) cp_async8(sm.alias(a) + lane, p.alias + idx);
            }
        };
        auto resolve = [&](const PairCursor &s, uint32_t a, uint32_t idx, uint32_t ry,
                           uint32_t &neg) -> uint32_t {
            neg = idx;
            if (p.use_alias && lane < K) {
                const uint2 e = sm.alias(a)[lane];
                neg = ry < e.x ? idx : e.y;
            }
            const uint32_t same = __match_any_sync(FULL, neg);
            const bool valid = lane < K && neg != s.c && neg != s.o &&
                               (same & ((1u << lane) - 1u)) == 0u;
            return (__ballot_sync(FULL, valid) << 1) | 1u;
        };

        PairCursor scan;
        scan.i = 0xFFFFFFFFu; scan.j = 0; scan.c = PAD; scan.o = PAD; scan.hi = 0;
        bool ok_cur = next_pair(walk, L, W, scan);
        if (!ok_cur) continue;
        PairCursor cur = scan;
        uint32_t stage = 0, slot_a = 0;

        // prologue: pair 0 synchronously, draw of pair 1 in flight
        uint32_t idx_n, ry_n, neg_cur, vmask_cur, ids_cur;
        draw(cur, slot_a, idx_n, ry_n);
        cp_async_commit();
        cp_async_wait_all();
        vmask_cur = resolve(cur, slot_a, idx_n, ry_n, neg_cur);
        ids_cur = slot_ids(lane, cur.o, neg_cur, vmask_cur);
        issue_rows(p, sm, stage, lane, chunks, ids_cur, vmask_cur, cur.c);
        bool ok_nxt = next_pair(walk, L, W, scan);
        PairCursor nxt = scan;
        slot_a ^= 1u;
        if (ok_nxt) draw(nxt, slot_a, idx_n, ry_n);
        cp_async_commit();

        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t loaded = 0xFFFFFFFFu;
        float lr = p.lr;
        while (ok_cur) {
            cp_async_wait_all();  // rows of `cur` (this stage) and alias entries of `nxt` are here
            // ---- pair p+1: resolve ids, copy its rows unless pair p is about to update one ----
            uint32_t neg_nxt = PAD, vmask_nxt = 0, ids_nxt = 0xFFFFFF00u | lane;
            bool deferred = false;
            if (ok_nxt) {
                vmask_nxt = resolve(nxt, slot_a, idx_n, ry_n, neg_nxt);
                ids_nxt = slot_ids(lane, nxt.o, neg_nxt, vmask_nxt);
                uint32_t moved = __shfl_sync(FULL, ids_nxt, lane & 15u);
                if (moved >= 0xFFFFFF00u) moved |= 16u;  // keep the two sentinel families apart
                const uint32_t both = lane < 16u ? ids_cur : moved;
                const uint32_t same = __match_any_sync(FULL, both);
                deferred = __ballot_sync(FULL, lane < 16u && (same >> 16) != 0u) != 0u ||
                           (nxt.i != cur.i && nxt.c == cur.c);
                if (!deferred)
                    issue_rows(p, sm, stage ^ 1u, lane, chunks, ids_nxt, vmask_nxt,
                               nxt.i != cur.i ? nxt.c : PAD);
            }
            // ---- pair p+2: start its draw ----
            const bool ok_far = ok_nxt && next_pair(walk, L, W, scan);
            const PairCursor far = scan;
            uint32_t idx_f = PAD, ry_f = 0;
            if (ok_far) draw(far, slot_a ^ 1u, idx_f, ry_f);
            cp_async_commit();

            // ---- pair p: train out of shared memory ----
            const float *rows = sm.rows(stage) + 4u * lane;
            const bool active = lane < chunks;
            if (loaded != cur.i) {
                h = active ? lds128(rows + (K + 1u) * p.row_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
                lr = centre_lr(p, cur.c);
                loaded = cur.i;
            }
            float part[16];
#pragma unroll
            for (int s = 0; s < PIPE_SLOTS; ++s) {
                float d = 0.0f;
                if (s <= (int)K && ((vmask_cur >> s) & 1u) && active) {
                    const float4 r = lds128(rows + (uint32_t)s * p.row_stride);
                    d = __fmaf_rn(h.x, r.x, d);
                    d = __fmaf_rn(h.y, r.y, d);
                    d = __fmaf_rn(h.z, r.z, d);
                    d = __fmaf_rn(h.w, r.w, d);
                }
                part[s] = d;
            }
            float f = reduce16(part, lane);  // lane l: score of slot (l >> 1) & 15
            if (p.scale_dot) f = __fmul_rn(f, p.inv_scale);
            const uint32_t my_slot = (lane >> 1) & 15u;
            const bool my_on = (vmask_cur >> my_slot) & 1u;  // vmask has no bit above K
            float g_mine = 0.0f;
            bool apply = false;
            if (my_on && !(fabsf(f) > p.clip)) {
                const float e = exp_det(-f);
                const float sigmoid = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
                g_mine = __fmul_rn(__fsub_rn(my_slot == 0 ? 1.0f : 0.0f, sigmoid), lr);
                // -log sigmoid(f) = log(1 + e^-f);  -log sigmoid(-f) = log(1 + e^-f) + f
                if ((lane & 1u) == 0) loss_acc += __logf(1.0f + e) + (my_slot == 0 ? 0.0f : f);
                apply = true;
            }
            const uint32_t amask = __ballot_sync(FULL, apply);  // bit 2s: slot s is applied
            n_targets += __popc(vmask_cur);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < PIPE_SLOTS; ++s) {
                if (s <= (int)K && ((amask >> (2 * s)) & 1u)) {
                    const float g = __shfl_sync(FULL, g_mine, 2 * s);
                    const uint32_t id = sm.ids(stage)[s];
                    if (active) {
                        float4 r = lds128(rows + (uint32_t)s * p.row_stride);
                        acc.x = __fmaf_rn(g, r.x, acc.x);
                        acc.y = __fmaf_rn(g, r.y, acc.y);
                        acc.z = __fmaf_rn(g, r.z, acc.z);
                        acc.w = __fmaf_rn(g, r.w, acc.w);
                        r.x = __fmaf_rn(g, h.x, r.x);
                        r.y = __fmaf_rn(g, h.y, r.y);
                        r.z = __fmaf_rn(g, h.z, r.z);
                        r.w = __fmaf_rn(g, h.w, r.w);
                        stg128(p.t1 + (uint64_t)id * p.row_stride + 4u * lane, r);
                    }
                }
            }
            h.x = __fadd_rn(h.x, acc.x);
            h.y = __fadd_rn(h.y, acc.y);
            h.z = __fadd_rn(h.z, acc.z);
            h.w = __fadd_rn(h.w, acc.w);
            ++n_pairs;
            if ((!ok_nxt || nxt.i != cur.i) && active)
                stg128(p.t0 + (uint64_t)cur.c * p.row_stride + 4u * lane, h);

            if (deferred) {  // its rows overlap the rows just stored: copy them now
                issue_rows(p, sm, stage ^ 1u, lane, chunks, ids_nxt, vmask_nxt,
                           nxt.i != cur.i ? nxt.c : PAD);
                cp_async_commit();
            }
            cur = nxt; neg_cur = neg_nxt; vmask_cur = vmask_nxt; ids_cur = ids_nxt; ok_cur = ok_nxt;
            nxt = far; idx_n = idx_f; ry_n = ry_f; ok_nxt = ok_far;
            stage ^= 1u;
            slot_a ^= 1u;
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

bool pipe_supported(const TrainParams &p, uint32_t model) {
    return model == B2E_SKIPGRAM && p.row_stride <= 128u && p.negatives + 1u <= PIPE_SLOTS;
}

cudaError_t launch_skipgram_pipe(const TrainParams &p, bool deterministic, int sm_count,
                                 cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(&p.counters->work_counter, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    const uint32_t stage_floats = (p.negatives + 2u) * p.row_stride;
    const size_t warp_bytes = 2u * stage_floats * 4u + 2u * 32u * 8u + 2u * PIPE_SLOTS * 4u;
    const int warps = deterministic ? 1 : 4;
    const size_t smem = warp_bytes * warps;
    static bool configured = false;
    if (!configured) {
        err = cudaFuncSetAttribute(skipgram_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    if (deterministic) {
        skipgram_pipe_kernel<<<1, 32, smem, stream>>>(p);
        return cudaGetLastError();
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skipgram_pipe_kernel, 128, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * per_sm;  // persistent: every CTA resident, walks fetched
    const uint64_t needed = (p.n_walks + warps - 1) / warps;
    if (grid > needed) grid = needed;
    skipgram_pipe_kernel<<<(unsigned)grid, 128, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace b2e
