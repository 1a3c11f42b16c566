// Microbenchmark (measurement tool, not product): how many random 4-byte gathers per second can
// B200's memory system serve from an array much larger than L2 (the access pattern of the walk
// kernel), as a function of cudaLimitMaxL2FetchGranularity and of the loads in flight per thread?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_gather scripts/microbench_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int ILP, bool DEPENDENT>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t *table, uint32_t n, uint32_t iters, uint32_t *sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t state[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) state[k] = hash32(tid * ILP + k + 1);
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = __ldg(table + __umulhi(state[k], n));
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            acc += v[k];
            state[k] = hash32(state[k] + (DEPENDENT ? v[k] : it));  // next address depends on the value (a walk)
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int ILP, bool DEPENDENT>
static void run(const uint32_t *table, uint32_t n, int blocks_per_sm, uint32_t *sink) {
    const uint32_t iters = 2048 / ILP;
    const int grid = 148 * blocks_per_sm;
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    gather_kernel<ILP, DEPENDENT><<<grid, 256>>>(table, n, iters, sink);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(a));
    gather_kernel<ILP, DEPENDENT><<<grid, 256>>>(table, n, iters, sink);
    CHECK(cudaEventRecord(b));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    const double gathers = (double)grid * 256 * iters * ILP;
    printf("  %s ILP %d  blocks/SM %d  %8.3f ms  %7.2f G gathers/s  (%.0f GB/s at 32 B per gather)\n",
           DEPENDENT ? "dependent  " : "independent", ILP, blocks_per_sm, ms, gathers / ms / 1e6, gathers * 32 / ms / 1e6);
}

int main() {
    const uint32_t n = 400000000u;  // 1.6 GB, like the indices of C3
    uint32_t *table, *sink;
    CHECK(cudaMalloc(&table, (size_t)n * 4));
    CHECK(cudaMemset(table, 1, (size_t)n * 4));
    CHECK(cudaMalloc(&sink, 4));
    for (size_t gran : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
        if (gran) CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
        size_t now = 0;
        CHECK(cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity));
        printf("cudaLimitMaxL2FetchGranularity = %zu%s\n", now, gran ? "" : " (default)");
        run<1, true>(table, n, 8, sink);
        run<2, true>(table, n, 8, sink);
        run<4, true>(table, n, 8, sink);
        run<4, true>(table, n, 4, sink);
        run<8, true>(table, n, 4, sink);
        run<8, false>(table, n, 4, sink);
    }
    return 0;
}
