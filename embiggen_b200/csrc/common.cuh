// Shared device helpers and the handle layout of the B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/b2e.h"

namespace b2e {

// Philox stream tags (top byte of counter word 3); the normative layout is in DESIGN.md.
constexpr uint32_t TAG_WALK1 = 1u;  // first-order steps, 4 per block
constexpr uint32_t TAG_WALK2 = 2u;  // second-order trials, 2 per block
constexpr uint32_t TAG_NEG = 3u;    // negative draws
constexpr uint32_t TAG_INIT0 = 4u;  // table 0 initialisation
constexpr uint32_t TAG_INIT1 = 5u;  // table 1 initialisation
constexpr uint32_t TAG_WALK3 = 7u;  // general walks (normalize_by_degree, typed): 1 trial per block
constexpr uint32_t TAG_SKIP = 6u;   // stochastic_downsample_by_degree, one draw per centre
constexpr uint32_t TAG_FOLD = 12u;  // second-order trials with the return edge folded: 1 trial per block
constexpr uint32_t MAX_TRIALS = 1u << 20;
constexpr uint32_t PAD = B2E_PAD_TOKEN;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                               uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

struct DeviceCounters {
    unsigned long long walk_steps;
    unsigned long long walk_trials;
    unsigned long long walk_searches;
    unsigned long long walk_probes;         // gathers spent on adjacency checks (filter words, row bounds, bisection)
    unsigned long long walk_filter_rejects; // adjacency checks answered "not a neighbour" by the row filter alone
    unsigned long long pairs;
    unsigned long long targets;
    double loss_sum;
    unsigned long long work_counter;  // dynamic walk fetch of the SGD kernels
};

struct WalkParams {
    const int64_t *indptr;
    const uint32_t *indices;
    const uint2 *edge_alias;  // per-row alias tables of a weighted graph ({thr, alias} per edge), or nullptr
    const uint32_t *node_types, *edge_types;  // [n] / [nnz] type ids of typed walks, or nullptr
    unsigned long long q_node[2], q_edge[2];  // accept thresholds [same type, changed type]
    const uint32_t *sources;
    uint64_t n_src;
    uint32_t seed_lo, seed_hi;
    uint64_t first_walk, n_walks, walk_id_stride;
    uint32_t walk_length;
    unsigned long long thr_return, thr_common, thr_explore;
    uint32_t *out;
    DeviceCounters *counters;
    uint32_t undirected;     // every edge has its mirror (verified at load): allows the short-row check
    const unsigned long long *filter;  // per-row blocked Bloom filters over the neighbour lists, or nullptr
    unsigned long long fold_excess;    // E > 0: the return edge is folded out of the envelope (TAG_FOLD)
    uint32_t occupancy;                // resident CTAs per SM the second-order kernel is compiled for (4, 5, 6)
    int sm_count;
};

struct TrainParams {
    const uint32_t *walks;
    uint64_t first_walk, n_walks, walk_id_stride;
    uint32_t seed_lo, seed_hi;
    uint32_t n;
    uint32_t walk_length, window, negatives;
    uint32_t row_stride;  // floats between rows in HBM: rows start on 128 B lines
    uint32_t chunks;      // float4 chunks of a row that hold data: ceil(embedding_size / 4)
    float clip, lr, inv_scale;
    uint32_t use_alias, normalize_lr, scale_dot;
    uint32_t downsample;  // stochastic_downsample_by_degree: max degree + 1, 0 = off
    uint32_t prefetch;  // 1: L2-prefetch the rows of the next draw site
    uint32_t variant;   // tuning variant of the launch (0 = default)
    uint32_t sgd_occupancy;  // CTAs per SM of the SkipGram kernel (0: by table size, see launch_train_pipe)
    uint32_t no_full_rows;  // B2E_NO_FULL_ROWS: keep the generic CBOW kernel for 32-chunk rows (A/B)
    uint32_t shared_negatives;  // SkipGram: one set of negatives per centre (skipgram_shared_kernel)
    uint32_t bulk;      // SkipGram rows by cp.async.bulk + mbarrier instead of per-lane cp.async (experiment)
    const uint2 *alias;  // {threshold, alias} per node
    const int64_t *indptr;
    float *t0, *t1;
    DeviceCounters *counters;
};

cudaError_t launch_walks(const WalkParams &p, bool second_order, cudaStream_t stream);
// GloVe (glove.cu): the co-occurrence triples of an epoch, sorted by (centre << 32 | context)
struct GloveState {
    unsigned long long *d_keys = nullptr;
    uint32_t *d_counts = nullptr;
    uint64_t *d_rowptr = nullptr;
    uint64_t n_triples = 0;
    uint32_t max_count = 1;
    bool finalised = false;
    size_t keys_bytes = 0, counts_bytes = 0, rowptr_bytes = 0;
    unsigned long long *d_scratch_keys = nullptr, *d_merge_keys = nullptr;
    uint32_t *d_merge_counts = nullptr;
    void *d_temp = nullptr, *d_scalar = nullptr;
    size_t scratch_keys_bytes = 0, merge_keys_bytes = 0, merge_counts_bytes = 0, temp_bytes = 0, scalar_bytes = 0;
    // co-occurrence by centre range: the walks of an epoch, their token positions bucketed by range
    uint32_t *d_epoch_walks = nullptr, *d_histogram = nullptr, *d_bounds = nullptr;
    unsigned long long *d_positions = nullptr, *d_cursor = nullptr;
    size_t epoch_walks_bytes = 0, histogram_bytes = 0, bounds_bytes = 0, positions_bytes = 0, cursor_bytes = 0;
    uint32_t last_ranges = 0;  // how many centre ranges the last epoch took (1 = one piece)
};
uint64_t glove_chunk_walks(uint32_t walk_length, uint32_t window);
cudaError_t glove_accumulate(GloveState &g, const uint32_t *d_walks, uint64_t n_walks, uint32_t walk_length,
                             uint32_t window, cudaStream_t stream);
cudaError_t glove_finalise(GloveState &g, uint64_t n, cudaStream_t stream);
cudaError_t glove_train(const GloveState &g, uint64_t n, uint32_t row_stride, uint32_t embedding_size,
                        float alpha, float clip, float lr, float *t0, float *t1, DeviceCounters *counters,
                        bool deterministic, int sm_count, uint64_t max_warps, uint32_t variant,
                        cudaStream_t stream);
// co-occurrence by centre range (glove.cu): histogram of the epoch's tokens, occurrences bucketed
// by range, the triples of one range
cudaError_t glove_token_histogram(const uint32_t *d_walks, uint64_t tokens, uint32_t *d_histogram, uint64_t n,
                                  cudaStream_t stream);
cudaError_t glove_bucket_positions(const uint32_t *d_walks, uint64_t tokens, const uint32_t *d_bounds, uint32_t ranges,
                                   unsigned long long *d_cursor, unsigned long long *d_positions,
                                   cudaStream_t stream);
cudaError_t glove_range_triples(GloveState &g, const uint32_t *d_walks, uint32_t L, uint32_t W,
                                const unsigned long long *d_positions, uint64_t count, cudaStream_t stream);
cudaError_t glove_reserve(void **ptr, size_t *have, size_t want);
cudaError_t glove_max_count(GloveState &g, uint32_t *max_count, cudaStream_t stream);
void glove_free(GloveState &g);

cudaError_t launch_walklet_split(const uint32_t *raw, uint64_t n_walks, uint32_t walk_length, uint32_t scale,
                                 uint32_t *out, cudaStream_t stream);
cudaError_t launch_symmetry_check(const int64_t *indptr, const uint32_t *indices, uint64_t n,
                                  uint64_t nnz, int *d_flag, cudaStream_t stream);
// *d_flags |= 1: a destination id >= n; |= 2: a row that is not strictly ascending
cudaError_t launch_csr_check(const int64_t *indptr, const uint32_t *indices, uint64_t n, int *d_flags,
                             int sm_count, cudaStream_t stream);
// alias_build.cu: offsets check (*d_flags |= 4), start nodes + maximum degree + alias table on the device
cudaError_t check_indptr_device(const int64_t *d_indptr, uint64_t n, uint64_t nnz, int *d_flags, cudaStream_t stream);
cudaError_t build_node_tables(const int64_t *d_indptr, const int64_t *host_indptr, uint64_t n, double alpha,
                              uint32_t *d_sources, uint64_t *n_src_out, uint64_t *max_degree_out, uint2 *d_table,
                              cudaStream_t stream, std::string &error);
uint64_t row_filter_words(uint64_t nnz);
cudaError_t launch_row_filter_build(const int64_t *indptr, const uint32_t *indices, uint64_t n,
                                    unsigned long long *filter, int sm_count, cudaStream_t stream);
cudaError_t launch_exchange_average(float *const t0[], float *const t1[], uint32_t world, uint32_t rank,
                                    uint64_t n, uint32_t row_stride, uint32_t chunks, int sm_count,
                                    cudaStream_t stream, uint32_t rows_per_iteration = 0);
cudaError_t launch_pack_rows(const float *table, uint64_t rows, uint32_t row_stride, uint32_t dim, float *dense,
                             int sm_count, cudaStream_t stream);
cudaError_t launch_tables_digest(const float *t0, const float *t1, uint64_t n, uint32_t row_stride,
                                 uint32_t dim, void *d_out, int sm_count, cudaStream_t stream);
cudaError_t launch_init_tables(float *t0, float *t1, uint64_t n, uint32_t embedding_size,
                               uint32_t row_stride, uint64_t seed, cudaStream_t stream);
// max_warps caps how many walks are trained concurrently (Hogwild staleness on small graphs)
cudaError_t launch_train(const TrainParams &p, uint32_t model, bool deterministic, int sm_count,
                         uint64_t max_warps, cudaStream_t stream);
bool pipe_supported(const TrainParams &p, uint32_t model);
bool shared_negatives_supported(const TrainParams &p);
cudaError_t launch_train_pipe(const TrainParams &p, uint32_t model, bool deterministic, int sm_count,
                              uint64_t max_warps, cudaStream_t stream);

// a CSR that stays in HBM (b2e_graph): built by graph_build.cu, handed to a handle without a copy
struct ResidentCsr {
    int64_t *indptr = nullptr;
    uint32_t *indices = nullptr;
    uint64_t n = 0, nnz = 0;
};
// `resident` != nullptr: the device arrays move into it and the host buffers are not touched
cudaError_t csr_from_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint64_t n,
                           int symmetrise, int64_t *indptr, uint32_t *indices, uint64_t capacity,
                           uint64_t *nnz_out, std::string &error, ResidentCsr *resident = nullptr);
cudaError_t synthetic_csr(int kind, uint64_t n, uint32_t scale, uint64_t m, uint64_t seed,
                          unsigned long long t_a, unsigned long long t_ab, unsigned long long t_abc,
                          int64_t *indptr, uint32_t *indices, uint64_t capacity, uint64_t *nnz_out,
                          std::string &error, ResidentCsr *resident = nullptr);

}  // namespace b2e

struct b2e_graph {
    b2e::ResidentCsr csr;
    int device = 0;
    int references = 1;  // the caller's object + every handle that walks on it
};

struct b2e_handle {
    b2e_config cfg;
    b2e_graph *shared_graph = nullptr;  // d_indptr / d_indices belong to it (b2e_load_graph)
    int sm_count = 0;
    uint64_t n = 0, nnz = 0, n_src = 0;
    uint32_t row_stride = 0;
    int64_t *d_indptr = nullptr;
    uint32_t *d_indices = nullptr;
    uint2 *d_edge_alias = nullptr;
    uint32_t *d_node_types = nullptr, *d_edge_types = nullptr;
    b2e::GloveState glove;
    uint32_t *d_walk_raw = nullptr;  // Walklets: the chunk as walked, before it is split by stride
    uint32_t max_degree = 0;
    uint32_t *d_sources = nullptr;
    unsigned long long *d_filter = nullptr;  // row filters of the adjacency check (second-order walks)
    uint64_t fold_excess = 0;                // see WalkParams
    unsigned long long thr_fold[3] = {0, 0, 0};
    uint2 *d_alias = nullptr;
    float *d_t0 = nullptr, *d_t1 = nullptr;
    uint32_t *d_walks[2] = {nullptr, nullptr};
    uint64_t chunk_cap = 0;
    uint64_t slot_first[2] = {0, 0}, slot_count[2] = {0, 0}, slot_stride[2] = {1, 1};
    cudaEvent_t walk_done[2] = {nullptr, nullptr}, train_done[2] = {nullptr, nullptr};
    cudaStream_t walk_stream = nullptr, train_stream = nullptr;
    bool own_streams = false;
    b2e::DeviceCounters *d_counters = nullptr;
    unsigned long long thr[3] = {0, 0, 0};
    bool second_order = false;
    bool undirected = false;
    uint32_t prefetch = 1;
    uint32_t variant = 0;
    uint32_t sgd_occupancy = 0;  // B2E_SGD_OCC (0: automatic)
    uint32_t bulk = 0;  // B2E_BULK
    uint32_t exchange_rows = 0;  // B2E_EXCHANGE_ROWS (0: the default of launch_exchange_average)
    uint32_t walk_occupancy = 6;  // B2E_WALK_OCC: see walk_kernel
    uint64_t launches = 0;
    std::vector<uint32_t> h_alias_thr, h_alias_idx;
    // the exchange step: replicas of the tables on the other GPUs of the node
    uint32_t world = 1, rank = 0;
    float *peer_t0[B2E_MAX_WORLD] = {nullptr}, *peer_t1[B2E_MAX_WORLD] = {nullptr};
    bool peers_are_ipc = false;
};
