// Microbenchmark (measurement tool, not product): what HBM bandwidth can B200 sustain for the
// access pattern of the SGD kernels -- random row gather + scatter of `row_floats`-wide fp32
// rows, one warp per row, `BATCH` rows in flight per warp?  Gives the practical ceiling the
// roofline fraction of train_kernel should be read against.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_rows scripts/microbench_rows.cu
//   scripts/microbench_rows
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int BATCH, bool WRITE, bool PREFETCH>
__global__ void __launch_bounds__(256) rows_kernel(float *table, uint32_t n_rows, uint32_t stride,
                                                   uint32_t row_floats, uint32_t rows_per_warp,
                                                   float *sink) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t chunks = row_floats >> 2;
    float accum = 0.f;
    uint32_t next_ids[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) next_ids[b] = __umulhi(hash32(warp * 7919u + b), n_rows);
    for (uint32_t it = 0; it < rows_per_warp; it += BATCH) {
        uint32_t ids[BATCH];
        float4 rows[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            ids[b] = next_ids[b];
            next_ids[b] = __umulhi(hash32((warp * 7919u + it + BATCH + b) * 2654435761u), n_rows);
        }
        if (PREFETCH) {
            // one 128 B line per lane: BATCH rows x 5 lines
            for (uint32_t t0 = 0; t0 < BATCH * 5u; t0 += 32u) {
                const uint32_t t = t0 + lane, slot = t / 5u, line = t - slot * 5u;
                uint32_t id = 0;
#pragma unroll
                for (int b = 0; b < BATCH; ++b) if (slot == (uint32_t)b) id = next_ids[b];
                if (slot < BATCH) {
                    const float *base = table + (uint64_t)id * stride;
                    const uint32_t head = (uint32_t)((uintptr_t)base & 127u);
                    if (line * 128u < head + row_floats * 4u)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)base - head + line * 128u));
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            const float *row = table + (uint64_t)ids[b] * stride;
            rows[b] = lane < chunks ? *reinterpret_cast<const float4 *>(row + 4 * lane)
                                    : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            accum += rows[b].x;
            if (WRITE) {
                rows[b].x += 1.0f; rows[b].y += 1.0f;
                float *row = table + (uint64_t)ids[b] * stride;
                if (lane < chunks) *reinterpret_cast<float4 *>(row + 4 * lane) = rows[b];
            }
        }
    }
    if (accum == 123.456f) sink[0] = accum;
}

template <int BATCH, bool WRITE, bool PREFETCH>
static void run(const char *name, float *table, uint32_t n_rows, uint32_t stride, uint32_t row_floats,
                int blocks_per_sm, float *sink) {
    const uint32_t rows_per_warp = 4096 / BATCH * BATCH;
    const int grid = 148 * blocks_per_sm;
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    rows_kernel<BATCH, WRITE, PREFETCH><<<grid, 256>>>(table, n_rows, stride, row_floats, rows_per_warp, sink);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(a));
    rows_kernel<BATCH, WRITE, PREFETCH><<<grid, 256>>>(table, n_rows, stride, row_floats, rows_per_warp, sink);
    CHECK(cudaEventRecord(b));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    const double rows = (double)grid * 8 * rows_per_warp;
    const double bytes = rows * row_floats * 4.0 * (WRITE ? 2.0 : 1.0);
    printf("%-34s stride %3u row %3u  blocks/SM %d  batch %2d  %8.3f ms  %7.1f GB/s (algorithmic)  %6.1f Mrows/s\n",
           name, stride, row_floats, blocks_per_sm, BATCH, ms, bytes / ms / 1e6, rows / ms / 1e3);
}

int main(int argc, char **argv) {
    const uint32_t n_rows = argc > 1 ? (uint32_t)atoi(argv[1]) : 2000000u;  // 2M rows ~ T0+T1 of C2
    float *sink;
    CHECK(cudaMalloc(&sink, 4));
    struct Case { uint32_t stride, row; };
    const Case cases[] = {{100u, 100u}, {104u, 100u}, {112u, 100u}, {128u, 100u}, {128u, 128u}};
    for (const Case &c : cases) {
        const uint32_t stride = c.stride, row = c.row;
        float *table;
        CHECK(cudaMalloc(&table, (size_t)n_rows * stride * 4));
        CHECK(cudaMemset(table, 0, (size_t)n_rows * stride * 4));
        printf("--- table %u rows x %u floats = %.0f MB, %u floats of each row touched\n", n_rows, stride,
               n_rows * (double)stride * 4 / 1e6, row);
        for (int occ : {2, 4}) {
            run<11, false, false>("gather", table, n_rows, stride, row, occ, sink);
            run<4, true, false>("gather+scatter", table, n_rows, stride, row, occ, sink);
            run<11, true, false>("gather+scatter", table, n_rows, stride, row, occ, sink);
        }
        CHECK(cudaFree(table));
    }
    return 0;
}
