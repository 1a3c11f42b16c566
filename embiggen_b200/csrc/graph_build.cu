// Graph ingest on the GPU: edge list -> sorted, de-duplicated,
// self-loop-free CSR, and the seeded synthetic generators of the BASELINE.json shapes.
//
// The reference receives a ready `ensmallen.Graph`; the only in-tree idiom for building one is
// the GraphBuilder loop of /root/reference/embiggen/utils/networkx_utils.py:79-113, and the CSR
// it hands over is described at .../embedders/pecanpy_embedders/node2vec.py:139-163 (rows
// sorted ascending).  This file produces exactly that layout.  Sorting / compaction are
// library calls (CUB ships with the CUDA toolkit); this is the step before the hot path, not
// the hot path.
//
// Synthetic graphs are "the first m distinct undirected edges, in draw order, of the Philox
// stream (seed, draw index)" -- the same definition embiggen_b200/graph.py implements in
// numpy, so both produce the same graph (tests/test_gpu_graph_build.py).
#include <cub/cub.cuh>

#include <algorithm>
#include <string>

#include "common.cuh"

namespace b2e {

constexpr uint32_t TAG_ER = 0x10u;    // graph.py: erdos_renyi
constexpr uint32_t TAG_RMAT = 0x11u;  // graph.py: rmat
constexpr unsigned long long INVALID_KEY = ~0ull;

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    ~DeviceBuffer() { cudaFree(ptr); }
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() { return static_cast<T *>(ptr); }
};

// candidate undirected edge of draw `idx`: key = (min << 32) | max, INVALID_KEY if rejected
__global__ void __launch_bounds__(256) draw_edges_kernel(int kind, uint64_t n, uint32_t scale,
                                                         uint32_t seed_lo, uint32_t seed_hi,
                                                         unsigned long long t_a, unsigned long long t_ab,
                                                         unsigned long long t_abc, uint64_t first_idx,
                                                         uint64_t count, unsigned long long *keys,
                                                         unsigned long long *ids) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t idx = first_idx + k;
    const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
    uint64_t u = 0, v = 0;
    if (kind == 0) {
        const uint4 r = philox4x32_10(seed_lo, seed_hi, lo, hi, 0u, TAG_ER << 24);
        u = __umulhi(r.x, (uint32_t)n);
        v = __umulhi(r.y, (uint32_t)n);
    } else {
        for (uint32_t block = 0; block * 4u < scale; ++block) {
            const uint4 r = philox4x32_10(seed_lo, seed_hi, lo, hi, block, TAG_RMAT << 24);
            const uint32_t words[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (uint32_t level = 0; level < 4u; ++level) {
                if (block * 4u + level >= scale) break;
                const unsigned long long w = words[level];
                const uint64_t bit_u = w >= t_ab;
                const uint64_t bit_v = ((w >= t_a) && (w < t_ab)) || (w >= t_abc);
                u = (u << 1) | bit_u;
                v = (v << 1) | bit_v;
            }
        }
    }
    const bool ok = u != v && u < n && v < n;
    const uint64_t a = u < v ? u : v, b = u < v ? v : u;
    keys[k] = ok ? ((a << 32) | b) : INVALID_KEY;
    ids[k] = idx;
}

// (src, dst) host-style edge list -> directed keys (src << 32) | dst, self-loops dropped
__global__ void __launch_bounds__(256) edge_keys_kernel(const uint32_t *src, const uint32_t *dst,
                                                        uint64_t count, uint64_t n, int symmetrise,
                                                        unsigned long long *keys, int *out_of_range) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t s = src[k], d = dst[k];
    if (s >= n || d >= n) *out_of_range = 1;
    const bool ok = s != d && s < n && d < n;
    keys[k] = ok ? ((s << 32) | d) : INVALID_KEY;
    if (symmetrise) keys[count + k] = ok ? ((d << 32) | s) : INVALID_KEY;
}

__global__ void __launch_bounds__(256) mirror_keys_kernel(const unsigned long long *undirected,
                                                          uint64_t count, unsigned long long *directed) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const unsigned long long key = undirected[k];
    directed[k] = key;
    directed[count + k] = (key << 32) | (key >> 32);
}

// sorted unique directed keys -> indices (low word) and indptr (lower bound of every row start)
__global__ void __launch_bounds__(256) csr_from_keys_kernel(const unsigned long long *keys, uint64_t nnz,
                                                            uint64_t n, uint32_t *indices,
                                                            long long *indptr) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) indices[k] = (uint32_t)keys[k];
    if (k <= n) {
        const unsigned long long target = (unsigned long long)k << 32;
        uint64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        indptr[k] = (long long)lo;
    }
}

static inline unsigned blocks_for(uint64_t count) { return (unsigned)((count + 255) / 256); }

#define GB_TRY(expr)                                                          \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            error = std::string(#expr) + ": " + cudaGetErrorString(_e);       \
            return _e;                                                        \
        }                                                                     \
    } while (0)

// stable radix sort of (keys, values) by key; results land in keys_out / values_out
static cudaError_t sort_pairs(DeviceBuffer &temp, const unsigned long long *keys_in,
                              unsigned long long *keys_out, const unsigned long long *values_in,
                              unsigned long long *values_out, uint64_t count, std::string &error) {
    size_t bytes = 0;
    GB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, values_in, values_out, count));
    GB_TRY(temp.reserve(bytes));
    GB_TRY(cub::DeviceRadixSort::SortPairs(temp.ptr, bytes, keys_in, keys_out, values_in, values_out, count));
    return cudaSuccess;
}

static cudaError_t sort_keys(DeviceBuffer &temp, const unsigned long long *keys_in,
                             unsigned long long *keys_out, uint64_t count, std::string &error) {
    size_t bytes = 0;
    GB_TRY(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys_in, keys_out, count));
    GB_TRY(temp.reserve(bytes));
    GB_TRY(cub::DeviceRadixSort::SortKeys(temp.ptr, bytes, keys_in, keys_out, count));
    return cudaSuccess;
}

// directed keys (any order) -> CSR in device buffers.  `scratch` is the alternate buffer of the
// radix sort (no third copy).  With `dedup` the keys may repeat and contain INVALID_KEY.
static cudaError_t keys_to_csr(DeviceBuffer &temp, unsigned long long *keys, unsigned long long *scratch,
                               uint64_t count, uint64_t n, bool dedup, DeviceBuffer &d_indptr,
                               DeviceBuffer &d_indices, uint64_t *nnz_out, std::string &error) {
    cub::DoubleBuffer<unsigned long long> buffers(keys, scratch);
    size_t bytes = 0;
    GB_TRY(cub::DeviceRadixSort::SortKeys(nullptr, bytes, buffers, count));
    GB_TRY(temp.reserve(bytes));
    GB_TRY(cub::DeviceRadixSort::SortKeys(temp.ptr, bytes, buffers, count));
    unsigned long long *sorted = buffers.Current(), *other = buffers.Alternate();
    uint64_t nnz = count;
    if (dedup) {
        DeviceBuffer selected;
        GB_TRY(selected.reserve(sizeof(unsigned long long)));
        GB_TRY(cub::DeviceSelect::Unique(nullptr, bytes, sorted, other, selected.as<unsigned long long>(), count));
        GB_TRY(temp.reserve(bytes));
        GB_TRY(cub::DeviceSelect::Unique(temp.ptr, bytes, sorted, other, selected.as<unsigned long long>(), count));
        unsigned long long unique = 0, last = 0;
        GB_TRY(cudaMemcpy(&unique, selected.ptr, sizeof(unique), cudaMemcpyDeviceToHost));
        if (unique) GB_TRY(cudaMemcpy(&last, other + unique - 1, sizeof(last), cudaMemcpyDeviceToHost));
        nnz = unique - ((unique && last == INVALID_KEY) ? 1 : 0);
        sorted = other;
    }
    *nnz_out = nnz;
    GB_TRY(d_indices.reserve(std::max<uint64_t>(nnz, 1) * sizeof(uint32_t)));
    GB_TRY(d_indptr.reserve((n + 1) * sizeof(long long)));
    csr_from_keys_kernel<<<blocks_for(std::max<uint64_t>(nnz, n + 1)), 256>>>(
        sorted, nnz, n, d_indices.as<uint32_t>(), d_indptr.as<long long>());
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaDeviceSynchronize());
    return cudaSuccess;
}

// hand the device arrays over: to host buffers, or (out != nullptr) to a resident b2e_graph
static cudaError_t deliver(DeviceBuffer &d_indptr, DeviceBuffer &d_indices, uint64_t n, uint64_t nnz,
                           int64_t *indptr, uint32_t *indices, uint64_t capacity, ResidentCsr *out,
                           std::string &error) {
    if (out) {
        out->indptr = d_indptr.as<int64_t>();
        out->indices = d_indices.as<uint32_t>();
        out->n = n;
        out->nnz = nnz;
        d_indptr.ptr = d_indices.ptr = nullptr;  // ownership moves
        d_indptr.bytes = d_indices.bytes = 0;
        return cudaSuccess;
    }
    if (nnz > capacity) {
        error = "indices buffer too small for the de-duplicated graph";
        return cudaErrorInvalidValue;
    }
    GB_TRY(cudaMemcpy(indptr, d_indptr.ptr, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
    if (nnz) GB_TRY(cudaMemcpy(indices, d_indices.ptr, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return cudaSuccess;
}

cudaError_t csr_from_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint64_t n,
                           int symmetrise, int64_t *indptr, uint32_t *indices, uint64_t capacity,
                           uint64_t *nnz_out, std::string &error, ResidentCsr *resident) {
    const uint64_t count = n_edges * (symmetrise ? 2 : 1);
    DeviceBuffer d_src, d_dst, keys, scratch, temp, flag;
    GB_TRY(d_src.reserve(std::max<uint64_t>(n_edges, 1) * sizeof(uint32_t)));
    GB_TRY(d_dst.reserve(std::max<uint64_t>(n_edges, 1) * sizeof(uint32_t)));
    GB_TRY(keys.reserve(std::max<uint64_t>(count, 1) * sizeof(unsigned long long)));
    GB_TRY(scratch.reserve(std::max<uint64_t>(count, 1) * sizeof(unsigned long long)));
    GB_TRY(flag.reserve(sizeof(int)));
    GB_TRY(cudaMemset(flag.ptr, 0, sizeof(int)));
    GB_TRY(cudaMemcpy(d_src.ptr, src, n_edges * sizeof(uint32_t), cudaMemcpyHostToDevice));
    GB_TRY(cudaMemcpy(d_dst.ptr, dst, n_edges * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (n_edges) {
        edge_keys_kernel<<<blocks_for(n_edges), 256>>>(d_src.as<uint32_t>(), d_dst.as<uint32_t>(), n_edges, n,
                                                       symmetrise, keys.as<unsigned long long>(),
                                                       flag.as<int>());
        GB_TRY(cudaGetLastError());
    }
    int out_of_range = 0;
    GB_TRY(cudaMemcpy(&out_of_range, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (out_of_range) {
        error = "an edge endpoint is not below the number of nodes";
        return cudaErrorInvalidValue;
    }
    d_src.release();
    d_dst.release();
    DeviceBuffer d_indptr, d_indices;
    GB_TRY(keys_to_csr(temp, keys.as<unsigned long long>(), scratch.as<unsigned long long>(), count, n, true,
                       d_indptr, d_indices, nnz_out, error));
    return deliver(d_indptr, d_indices, n, *nnz_out, indptr, indices, capacity, resident, error);
}

// flags[j] = 1 when keys[j] is a real key that the sorted pool does not hold yet
__global__ void __launch_bounds__(256) new_key_flags_kernel(const unsigned long long *keys, uint64_t count,
                                                            const unsigned long long *pool, uint64_t have,
                                                            unsigned char *flags) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const unsigned long long key = keys[k];
    uint64_t lo = 0, hi = have;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (pool[mid] < key) lo = mid + 1; else hi = mid;
    }
    flags[k] = key != INVALID_KEY && !(lo < have && pool[lo] == key);
}

template <typename T>
static cudaError_t select_flagged(DeviceBuffer &temp, const T *in, const unsigned char *flags, T *out,
                                  unsigned long long *d_count, uint64_t count, std::string &error) {
    size_t bytes = 0;
    GB_TRY(cub::DeviceSelect::Flagged(nullptr, bytes, in, flags, out, d_count, count));
    GB_TRY(temp.reserve(bytes));
    GB_TRY(cub::DeviceSelect::Flagged(temp.ptr, bytes, in, flags, out, d_count, count));
    return cudaSuccess;
}

// The first m distinct undirected edges of the Philox stream, in draw order.  Candidates are
// drawn in batches; a batch is sorted, de-duplicated (earliest draw wins), filtered against the
// sorted pool of edges already kept, and merged into it.  Every pool edge was first drawn before
// any edge of a later batch, so only the last batch can overshoot m, and only its new edges
// need their draw index to be cut back to "the m earliest".  Memory: two pool buffers of m keys
// plus a few batch-sized arrays, which is what makes the 2-billion-edge shape fit one GPU.
cudaError_t synthetic_csr(int kind, uint64_t n, uint32_t scale, uint64_t m, uint64_t seed,
                          unsigned long long t_a, unsigned long long t_ab, unsigned long long t_abc,
                          int64_t *indptr, uint32_t *indices, uint64_t capacity, uint64_t *nnz_out,
                          std::string &error, ResidentCsr *resident) {
    typedef unsigned long long u64;
    const uint64_t batch_max = 1ull << 29;
    DeviceBuffer pool[2], keys_in, ids_in, keys_sorted, ids_sorted, keys_unique, ids_unique, keys_new,
        ids_new, flags, temp, d_count;
    GB_TRY(d_count.reserve(sizeof(u64)));
    GB_TRY(pool[0].reserve(std::max<uint64_t>(m, 1) * sizeof(u64)));
    GB_TRY(pool[1].reserve(std::max<uint64_t>(m, 1) * sizeof(u64)));
    int cur = 0;
    uint64_t have = 0, drawn = 0;
    for (int round = 0; have < m; ++round) {
        if (round > 400) {
            error = "could not draw enough distinct edges (graph too dense for its shape?)";
            return cudaErrorInvalidValue;
        }
        const uint64_t want = std::min<uint64_t>(batch_max, (uint64_t)((double)(m - have) * 1.3) + 1024);
        for (DeviceBuffer *buffer : {&keys_in, &ids_in, &keys_sorted, &ids_sorted, &keys_unique, &ids_unique,
                                     &keys_new, &ids_new})
            GB_TRY(buffer->reserve(want * sizeof(u64)));
        GB_TRY(flags.reserve(want));
        draw_edges_kernel<<<blocks_for(want), 256>>>(kind, n, scale, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                     t_a, t_ab, t_abc, drawn, want, keys_in.as<u64>(),
                                                     ids_in.as<u64>());
        GB_TRY(cudaGetLastError());
        drawn += want;
        // stable sort by key: equal keys stay in ascending draw order, so "first" = earliest draw
        GB_TRY(sort_pairs(temp, keys_in.as<u64>(), keys_sorted.as<u64>(), ids_in.as<u64>(),
                          ids_sorted.as<u64>(), want, error));
        size_t bytes = 0;
        GB_TRY(cub::DeviceSelect::UniqueByKey(nullptr, bytes, keys_sorted.as<u64>(), ids_sorted.as<u64>(),
                                              keys_unique.as<u64>(), ids_unique.as<u64>(), d_count.as<u64>(),
                                              want));
        GB_TRY(temp.reserve(bytes));
        GB_TRY(cub::DeviceSelect::UniqueByKey(temp.ptr, bytes, keys_sorted.as<u64>(), ids_sorted.as<u64>(),
                                              keys_unique.as<u64>(), ids_unique.as<u64>(), d_count.as<u64>(),
                                              want));
        u64 unique = 0;
        GB_TRY(cudaMemcpy(&unique, d_count.ptr, sizeof(unique), cudaMemcpyDeviceToHost));
        if (unique == 0) continue;
        new_key_flags_kernel<<<blocks_for(unique), 256>>>(keys_unique.as<u64>(), unique, pool[cur].as<u64>(),
                                                          have, flags.as<unsigned char>());
        GB_TRY(cudaGetLastError());
        GB_TRY(select_flagged(temp, keys_unique.as<u64>(), flags.as<unsigned char>(), keys_new.as<u64>(),
                              d_count.as<u64>(), unique, error));
        GB_TRY(select_flagged(temp, ids_unique.as<u64>(), flags.as<unsigned char>(), ids_new.as<u64>(),
                              d_count.as<u64>(), unique, error));
        u64 fresh = 0;
        GB_TRY(cudaMemcpy(&fresh, d_count.ptr, sizeof(fresh), cudaMemcpyDeviceToHost));
        if (fresh == 0) continue;
        if (have + fresh > m) {  // the last batch overshoots: keep its m - have earliest draws
            const uint64_t keep = m - have;
            GB_TRY(sort_pairs(temp, ids_new.as<u64>(), ids_sorted.as<u64>(), keys_new.as<u64>(),
                              keys_sorted.as<u64>(), fresh, error));
            GB_TRY(sort_keys(temp, keys_sorted.as<u64>(), keys_new.as<u64>(), keep, error));
            fresh = keep;
        }
        GB_TRY(cub::DeviceMerge::MergeKeys(nullptr, bytes, pool[cur].as<u64>(), (int64_t)have,
                                           keys_new.as<u64>(), (int64_t)fresh, pool[cur ^ 1].as<u64>()));
        GB_TRY(temp.reserve(bytes));
        GB_TRY(cub::DeviceMerge::MergeKeys(temp.ptr, bytes, pool[cur].as<u64>(), (int64_t)have,
                                           keys_new.as<u64>(), (int64_t)fresh, pool[cur ^ 1].as<u64>()));
        cur ^= 1;
        have += fresh;
    }
    for (DeviceBuffer *buffer : {&keys_in, &ids_in, &keys_sorted, &ids_sorted, &keys_unique, &ids_unique,
                                 &keys_new, &ids_new, &flags, &pool[cur ^ 1]}) {
        cudaFree(buffer->ptr);
        buffer->ptr = nullptr;
        buffer->bytes = 0;
    }
    // both directions, sorted -> CSR
    DeviceBuffer directed, alternate;
    GB_TRY(directed.reserve(2 * have * sizeof(u64)));
    mirror_keys_kernel<<<blocks_for(have), 256>>>(pool[cur].as<u64>(), have, directed.as<u64>());
    GB_TRY(cudaGetLastError());
    GB_TRY(cudaDeviceSynchronize());
    cudaFree(pool[cur].ptr);
    pool[cur].ptr = nullptr;
    pool[cur].bytes = 0;
    GB_TRY(alternate.reserve(2 * have * sizeof(u64)));
    DeviceBuffer d_indptr, d_indices;
    GB_TRY(keys_to_csr(temp, directed.as<u64>(), alternate.as<u64>(), 2 * have, n, false, d_indptr, d_indices,
                       nnz_out, error));
    directed.release();
    alternate.release();
    temp.release();
    return deliver(d_indptr, d_indices, n, *nnz_out, indptr, indices, capacity, resident, error);
}

}  // namespace b2e
