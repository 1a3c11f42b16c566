// K3 on the GPU: start-node list, maximum degree and the alias table over deg^alpha of the
// negative draws (`use_scale_free_distribution`,
// /root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:101-102), built from
// the device-resident offsets without a pass on the host.
//
// The table is the one oracle/alias.c defines (integer masses m_i, bucket capacity c, light and
// heavy nodes swept in node order), computed in closed form instead of by the sequential sweep.
// With D_k the deficits (c - m) of the light nodes before light node k and S_j the surpluses
// (m - c) of the heavy nodes up to and including heavy node j:
//   light node k  keeps m in its bucket and is topped up by heavy node  #{ j : S_j < D_k };
//   heavy node j  is exhausted by the first light node k with D_k + deficit_k > S_j; it then owns
//                 its bucket with  c + S_j - (D_k + deficit_k)  and is topped up by heavy node
//                 j + 1; a heavy node that is never exhausted owns a full bucket.
// Two compactions, two prefix sums (CUB: library plumbing) and two binary-search kernels; every
// quantity is an integer below 2^62, so the result does not depend on the order of the sums and
// is bit-identical to the oracle's (tests/test_gpu_sgns.py::test_alias_table_bit_exact).
#include <cub/cub.cuh>

#include <cmath>
#include <string>

#include "common.cuh"

namespace b2e {

struct DeviceScratch {
    void *ptr = nullptr;
    ~DeviceScratch() { cudaFree(ptr); }
    cudaError_t get(size_t bytes) {
        cudaFree(ptr);
        ptr = nullptr;
        return cudaMalloc(&ptr, bytes ? bytes : 1);
    }
    template <typename T> T *as() { return static_cast<T *>(ptr); }
};

#define AB_TRY(expr)                                                     \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) {                                         \
            error = std::string(#expr) + ": " + cudaGetErrorString(_e);  \
            return _e;                                                   \
        }                                                                \
    } while (0)

__device__ __forceinline__ double degree_weight_device(unsigned long long deg, int form) {
    const double d = (double)deg;
    if (deg == 0) return 0.0;
    if (form == 0) return 1.0;                      // alpha = 0
    if (form == 1) return d;                        // alpha = 1
    if (form == 2) return sqrt(d);                  // alpha = 0.5
    return sqrt(sqrt(__dmul_rn(__dmul_rn(d, d), d)));  // alpha = 0.75
}

// offsets sane?  flags |= 4 otherwise (checked before any kernel trusts them)
__global__ void __launch_bounds__(256) indptr_check_kernel(const int64_t *__restrict__ indptr, uint64_t n, uint64_t nnz,
                                                           int *flags) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int64_t a = indptr[v], b = indptr[v + 1];
    if (a < 0 || b < a || (uint64_t)b > nnz) atomicOr(flags, 4);
}

// fixed-point weight of every node (or the host-provided one), "has edges" flag, degree
__global__ void __launch_bounds__(256) alias_weight_kernel(const int64_t *__restrict__ indptr, uint64_t n, int form,
                                                           int bits, const unsigned long long *__restrict__ given,
                                                           unsigned long long *__restrict__ weight,
                                                           uint32_t *__restrict__ has_edges,
                                                           unsigned long long *__restrict__ degree) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const unsigned long long deg = (unsigned long long)(indptr[v + 1] - indptr[v]);
    has_edges[v] = deg > 0;
    degree[v] = deg;
    if (weight) weight[v] = given ? given[v] : (unsigned long long)floor(ldexp(degree_weight_device(deg, form), bits));
}

__global__ void __launch_bounds__(256) sources_kernel(const uint32_t *__restrict__ has_edges,
                                                      const uint32_t *__restrict__ rank, uint64_t n,
                                                      uint32_t *__restrict__ sources) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n && has_edges[v]) sources[rank[v]] = (uint32_t)v;
}

// final masses and the light flag
__global__ void __launch_bounds__(256) alias_mass_kernel(unsigned long long *__restrict__ mass,
                                                         const uint32_t *__restrict__ has_edges,
                                                         const uint32_t *__restrict__ rank, uint64_t n,
                                                         unsigned long long each, unsigned long long first,
                                                         unsigned long long capacity, uint32_t *__restrict__ is_light) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    unsigned long long m = mass[v];
    if (has_edges[v]) m += each + (rank[v] < first ? 1ull : 0ull);
    mass[v] = m;
    is_light[v] = m < capacity;
}

// node ids and deficits / surpluses of the two classes, each in node order
__global__ void __launch_bounds__(256) alias_split_kernel(const unsigned long long *__restrict__ mass,
                                                          const uint32_t *__restrict__ is_light,
                                                          const uint32_t *__restrict__ light_rank, uint64_t n,
                                                          unsigned long long capacity, uint32_t *__restrict__ light,
                                                          uint32_t *__restrict__ heavy,
                                                          unsigned long long *__restrict__ deficit,
                                                          unsigned long long *__restrict__ surplus) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t k = light_rank[v];
    if (is_light[v]) {
        light[k] = (uint32_t)v;
        deficit[k] = capacity - mass[v];
    } else {
        heavy[v - k] = (uint32_t)v;
        surplus[v - k] = mass[v] - capacity;
    }
}

// floor(mass 2^32 / capacity) for mass < capacity < 2^62 (oracle/alias.c: threshold)
__device__ __forceinline__ uint32_t alias_threshold(unsigned long long mass, unsigned long long capacity) {
    unsigned long long r = mass;
    uint32_t q = 0;
#pragma unroll 4
    for (int bit = 0; bit < 32; ++bit) {
        r <<= 1;
        q <<= 1;
        if (r >= capacity) { r -= capacity; q |= 1u; }
    }
    return q;
}

__global__ void __launch_bounds__(256) alias_light_kernel(const uint32_t *__restrict__ light,
                                                          const uint32_t *__restrict__ heavy,
                                                          const unsigned long long *__restrict__ deficit,
                                                          const unsigned long long *__restrict__ deficit_sum,
                                                          const unsigned long long *__restrict__ surplus_sum,
                                                          uint64_t n_light, uint64_t n_heavy,
                                                          unsigned long long capacity, uint2 *__restrict__ table) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_light) return;
    const unsigned long long before = deficit_sum[k] - deficit[k];
    uint64_t lo = 0, hi = n_heavy;  // heavy nodes exhausted before this one: surplus_sum < before
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(surplus_sum + mid) < before) lo = mid + 1; else hi = mid;
    }
    table[light[k]] = make_uint2(alias_threshold(capacity - deficit[k], capacity), heavy[lo]);
}

__global__ void __launch_bounds__(256) alias_heavy_kernel(const uint32_t *__restrict__ heavy,
                                                          const unsigned long long *__restrict__ deficit_sum,
                                                          const unsigned long long *__restrict__ surplus_sum,
                                                          uint64_t n_light, uint64_t n_heavy,
                                                          unsigned long long capacity, uint2 *__restrict__ table) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_heavy) return;
    const unsigned long long mine = surplus_sum[j];
    uint64_t lo = 0, hi = n_light;  // first light node whose running deficit exceeds my running surplus
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(deficit_sum + mid) > mine) hi = mid; else lo = mid + 1;
    }
    const uint32_t v = heavy[j];
    if (lo >= n_light || j + 1 >= n_heavy) {
        table[v] = make_uint2(0xFFFFFFFFu, v);  // never exhausted: a full bucket
    } else {
        table[v] = make_uint2(alias_threshold(capacity + mine - deficit_sum[lo], capacity), heavy[j + 1]);
    }
}

static inline unsigned grid_for(uint64_t count) { return (unsigned)((count + 255) / 256); }

static int exact_form(double alpha) {
    if (alpha == 0.0) return 0;
    if (alpha == 1.0) return 1;
    if (alpha == 0.5) return 2;
    if (alpha == 0.75) return 3;
    return -1;
}

static double degree_weight_host(uint64_t deg, double alpha) {
    const double d = (double)deg;
    if (deg == 0) return 0.0;
    if (alpha == 0.0) return 1.0;
    if (alpha == 1.0) return d;
    if (alpha == 0.5) return std::sqrt(d);
    if (alpha == 0.75) return std::sqrt(std::sqrt(d * d * d));
    return std::pow(d, alpha);
}

int alias_fraction_bits(uint64_t n, uint64_t max_degree, double alpha) {
    const double w_max = degree_weight_host(max_degree, alpha);
    int bits = 40;
    while (bits > 0 && (double)n * (std::floor(std::ldexp(w_max, bits)) + 1.0) >= 4611686018427387904.0) --bits;
    return bits;
}

cudaError_t check_indptr_device(const int64_t *d_indptr, uint64_t n, uint64_t nnz, int *d_flags, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    indptr_check_kernel<<<grid_for(n), 256, 0, stream>>>(d_indptr, n, nnz, d_flags);
    return cudaGetLastError();
}

// Start nodes (always) and the alias table (when `d_table` is given) from the device offsets.
// `host_indptr` is needed only for an exponent without an exact form (libm's pow is the oracle's).
cudaError_t build_node_tables(const int64_t *d_indptr, const int64_t *host_indptr, uint64_t n, double alpha,
                              uint32_t *d_sources, uint64_t *n_src_out, uint64_t *max_degree_out, uint2 *d_table,
                              cudaStream_t stream, std::string &error) {
    typedef unsigned long long u64;
    DeviceScratch weight, has_edges, rank, degree, temp, scalar, given;
    AB_TRY(has_edges.get(n * sizeof(uint32_t)));
    AB_TRY(rank.get(n * sizeof(uint32_t)));
    AB_TRY(degree.get(n * sizeof(u64)));
    AB_TRY(scalar.get(2 * sizeof(u64)));
    size_t bytes = 0, need = 0;
    // pass 1: degrees, flags; the largest degree fixes the fixed-point format
    alias_weight_kernel<<<grid_for(n), 256, 0, stream>>>(d_indptr, n, 0, 0, nullptr, nullptr, has_edges.as<uint32_t>(),
                                                         degree.as<u64>());
    AB_TRY(cudaGetLastError());
    AB_TRY(cub::DeviceReduce::Max(nullptr, bytes, degree.as<u64>(), scalar.as<u64>(), n, stream));
    need = bytes;
    AB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, has_edges.as<uint32_t>(), rank.as<uint32_t>(), n, stream));
    need = std::max(need, bytes);
    AB_TRY(cub::DeviceScan::InclusiveSum(nullptr, bytes, degree.as<u64>(), degree.as<u64>(), n, stream));
    need = std::max(need, bytes);
    AB_TRY(cub::DeviceReduce::Sum(nullptr, bytes, degree.as<u64>(), scalar.as<u64>(), n, stream));
    need = std::max(need, bytes);
    AB_TRY(temp.get(need));
    bytes = need;
    AB_TRY(cub::DeviceReduce::Max(temp.ptr, bytes, degree.as<u64>(), scalar.as<u64>(), n, stream));
    bytes = need;
    AB_TRY(cub::DeviceScan::ExclusiveSum(temp.ptr, bytes, has_edges.as<uint32_t>(), rank.as<uint32_t>(), n, stream));
    sources_kernel<<<grid_for(n), 256, 0, stream>>>(has_edges.as<uint32_t>(), rank.as<uint32_t>(), n, d_sources);
    AB_TRY(cudaGetLastError());
    u64 max_degree = 0;
    uint32_t last_rank = 0, last_flag = 0;
    AB_TRY(cudaMemcpyAsync(&max_degree, scalar.ptr, sizeof(u64), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaMemcpyAsync(&last_rank, rank.as<uint32_t>() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaMemcpyAsync(&last_flag, has_edges.as<uint32_t>() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaStreamSynchronize(stream));
    const uint64_t n_src = (uint64_t)last_rank + last_flag;
    *n_src_out = n_src;
    *max_degree_out = max_degree;
    if (!d_table) return cudaSuccess;
    if (n_src == 0) {
        error = "alias table: total weight is zero";
        return cudaErrorInvalidValue;
    }

    // pass 2: fixed-point weights and their total
    const int bits = alias_fraction_bits(n, max_degree, alpha);
    const int form = exact_form(alpha);
    AB_TRY(weight.get(n * sizeof(u64)));
    if (form < 0) {  // libm's pow on the host, like the oracle; one upload of 8 bytes per node
        if (!host_indptr) {
            error = "a negative_sampling_exponent other than 0, 0.5, 0.75, 1 needs the offsets on the host";
            return cudaErrorInvalidValue;
        }
        std::vector<u64> host(n);
        for (uint64_t v = 0; v < n; ++v)
            host[v] = (u64)std::floor(std::ldexp(degree_weight_host((uint64_t)(host_indptr[v + 1] - host_indptr[v]), alpha), bits));
        AB_TRY(given.get(n * sizeof(u64)));
        AB_TRY(cudaMemcpyAsync(given.ptr, host.data(), n * sizeof(u64), cudaMemcpyHostToDevice, stream));
        AB_TRY(cudaStreamSynchronize(stream));
    }
    alias_weight_kernel<<<grid_for(n), 256, 0, stream>>>(d_indptr, n, form < 0 ? 0 : form, bits,
                                                         form < 0 ? given.as<u64>() : nullptr, weight.as<u64>(),
                                                         has_edges.as<uint32_t>(), degree.as<u64>());
    AB_TRY(cudaGetLastError());
    bytes = need;
    AB_TRY(cub::DeviceReduce::Sum(temp.ptr, bytes, weight.as<u64>(), scalar.as<u64>(), n, stream));
    u64 total = 0;
    AB_TRY(cudaMemcpyAsync(&total, scalar.ptr, sizeof(u64), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaStreamSynchronize(stream));
    if (total == 0) {
        error = "alias table: total weight is zero";
        return cudaErrorInvalidValue;
    }
    const u64 capacity = (total + n - 1) / n;
    const u64 excess = capacity * n - total, each = excess / n_src, first = excess % n_src;

    // pass 3: masses, classes, the two node lists with their deficits / surpluses
    DeviceScratch is_light, light_rank, light, heavy, deficit, surplus;
    AB_TRY(is_light.get(n * sizeof(uint32_t)));
    AB_TRY(light_rank.get(n * sizeof(uint32_t)));
    alias_mass_kernel<<<grid_for(n), 256, 0, stream>>>(weight.as<u64>(), has_edges.as<uint32_t>(), rank.as<uint32_t>(), n,
                                                       each, first, capacity, is_light.as<uint32_t>());
    AB_TRY(cudaGetLastError());
    bytes = need;
    AB_TRY(cub::DeviceScan::ExclusiveSum(temp.ptr, bytes, is_light.as<uint32_t>(), light_rank.as<uint32_t>(), n, stream));
    uint32_t tail_rank = 0, tail_flag = 0;
    AB_TRY(cudaMemcpyAsync(&tail_rank, light_rank.as<uint32_t>() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaMemcpyAsync(&tail_flag, is_light.as<uint32_t>() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    AB_TRY(cudaStreamSynchronize(stream));
    const uint64_t n_light = (uint64_t)tail_rank + tail_flag, n_heavy = n - n_light;
    if (n_heavy == 0) {  // impossible: the masses add up to n capacities
        error = "alias table: no heavy node";
        return cudaErrorInvalidValue;
    }
    AB_TRY(light.get(n_light * sizeof(uint32_t)));
    AB_TRY(heavy.get(n_heavy * sizeof(uint32_t)));
    AB_TRY(deficit.get(n_light * sizeof(u64)));
    AB_TRY(surplus.get(n_heavy * sizeof(u64)));
    alias_split_kernel<<<grid_for(n), 256, 0, stream>>>(weight.as<u64>(), is_light.as<uint32_t>(),
                                                        light_rank.as<uint32_t>(), n, capacity, light.as<uint32_t>(),
                                                        heavy.as<uint32_t>(), deficit.as<u64>(), surplus.as<u64>());
    AB_TRY(cudaGetLastError());
    // running sums: the deficits keep their own copy (the light kernel needs both), in `degree`
    u64 *deficit_sum = degree.as<u64>();
    if (n_light) {
        bytes = need;
        AB_TRY(cub::DeviceScan::InclusiveSum(temp.ptr, bytes, deficit.as<u64>(), deficit_sum, n_light, stream));
    }
    bytes = need;
    AB_TRY(cub::DeviceScan::InclusiveSum(temp.ptr, bytes, surplus.as<u64>(), surplus.as<u64>(), n_heavy, stream));
    if (n_light) {
        alias_light_kernel<<<grid_for(n_light), 256, 0, stream>>>(light.as<uint32_t>(), heavy.as<uint32_t>(),
                                                                  deficit.as<u64>(), deficit_sum, surplus.as<u64>(),
                                                                  n_light, n_heavy, capacity, d_table);
        AB_TRY(cudaGetLastError());
    }
    alias_heavy_kernel<<<grid_for(n_heavy), 256, 0, stream>>>(heavy.as<uint32_t>(), deficit_sum, surplus.as<u64>(), n_light,
                                                              n_heavy, capacity, d_table);
    AB_TRY(cudaGetLastError());
    AB_TRY(cudaStreamSynchronize(stream));
    return cudaSuccess;
}

}  // namespace b2e
