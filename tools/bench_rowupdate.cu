// Micro-benchmark: what HBM throughput does the access pattern of the sparse optimizer update allow?
// U sorted-random rows of a [V][3][S] fp32 table are read (3 planes = 3*S*4 bytes contiguous) and written back.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bench_rowupdate.cu -o /tmp/bench_rowupdate
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int ROWS_PER_WARP_IN_FLIGHT>
__global__ void row_update(float *table, const int *ids, int U, int S, int nwarps_per_row_unused) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int F = 3 * S / 4;  // float4 per row record
    for (int u0 = warp * ROWS_PER_WARP_IN_FLIGHT; u0 < U; u0 += nwarps * ROWS_PER_WARP_IN_FLIGHT) {
        float4 v[ROWS_PER_WARP_IN_FLIGHT][8];
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP_IN_FLIGHT; ++j) {
            const int u = u0 + j;
            if (u < U) {
                const float4 *row = reinterpret_cast<const float4 *>(table + (size_t)ids[u] * 3 * S);
#pragma unroll
                for (int r = 0; r < 8; ++r) { const int f = lane + 32 * r; if (f < F) v[j][r] = row[f]; }
            }
        }
#pragma unroll
        for (int j = 0; j < ROWS_PER_WARP_IN_FLIGHT; ++j) {
            const int u = u0 + j;
            if (u < U) {
                float4 *row = reinterpret_cast<float4 *>(table + (size_t)ids[u] * 3 * S);
#pragma unroll
                for (int r = 0; r < 8; ++r) { const int f = lane + 32 * r; if (f < F) { float4 t = v[j][r]; t.x += 1.f; row[f] = t; } }
            }
        }
    }
}

int main() {
    const int V = 400000, S = 304, U = 47000, reps = 20;
    float *table; cudaMalloc(&table, (size_t)V * 3 * S * 4); cudaMemset(table, 0, (size_t)V * 3 * S * 4);
    int *ids; cudaMalloc(&ids, U * 4);
    float *flush; cudaMalloc(&flush, 512u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm : {2, 4, 8}) {
        for (int variant = 0; variant < 2; ++variant) {
            float best = 1e9, tot = 0;
            for (int rep = 0; rep < reps; ++rep) {
                std::vector<int> h(U);
                for (auto &x : h) x = (int)((double)rand() / RAND_MAX * (V - 1));
                std::sort(h.begin(), h.end());
                cudaMemcpy(ids, h.data(), U * 4, cudaMemcpyHostToDevice);
                cudaMemset(flush, rep, 512u << 20);  // evict L2
                cudaEventRecord(e0);
                if (variant == 0) row_update<1><<<148 * blocks_per_sm, 256>>>(table, ids, U, S, 0);
                else row_update<2><<<148 * blocks_per_sm, 256>>>(table, ids, U, S, 0);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep >= 2) { best = std::min(best, ms); tot += ms; }
            }
            const double bytes = 2.0 * U * 3 * S * 4;
            printf("warps/SM %2d rows_in_flight/warp %d: avg %.1f us  best %.1f us  -> %.0f GB/s (avg)\n", blocks_per_sm * 8,
                   variant + 1, 1e3 * tot / (reps - 2), 1e3 * best, bytes / (tot / (reps - 2) * 1e-3) / 1e9);
        }
    }
    return 0;
}
