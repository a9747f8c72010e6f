// How fast can a B200 gather random 1 KB rows (and random 128 B rows) from HBM, with nothing else to do?
// The "gather roofline" for the Phase III / Phase II rescoring kernels (BASELINE config 5: 4096 x 1000 random candidates):
// a warp reads ROW_BYTES contiguous bytes per random row, U rows in flight per warp, XORs them into a sink.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int ROW_BYTES, int U>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ rows, const int64_t* __restrict__ idx, int64_t n_idx,
                                                     uint32_t* __restrict__ sink) {
    constexpr int V = ROW_BYTES / 512;  // uint4 per lane per row (1 KB: 2; 128 B rows are read as one uint32 per lane)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t acc = 0;
    for (int64_t i0 = warp * U; i0 < n_idx; i0 += nwarps * U) {
        if (ROW_BYTES >= 512) {
            uint4 v[U][V > 0 ? V : 1];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = i0 + u < n_idx ? idx[i0 + u] : idx[i0];
                const uint4* src = reinterpret_cast<const uint4*>(rows + (size_t)r * ROW_BYTES);
#pragma unroll
                for (int k = 0; k < V; k++) v[u][k] = __ldg(src + 32 * k + lane);
            }
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int k = 0; k < V; k++) acc ^= v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
        } else {
            uint32_t v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = i0 + u < n_idx ? idx[i0 + u] : idx[i0];
                v[u] = __ldg(reinterpret_cast<const uint32_t*>(rows + (size_t)r * ROW_BYTES) + lane);
            }
#pragma unroll
            for (int u = 0; u < U; u++) acc ^= v[u];
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int ROW_BYTES, int U>
void run(const uint8_t* rows, const int64_t* idx, int64_t n_idx, uint32_t* sink, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    gather_kernel<ROW_BYTES, U><<<grid, 256>>>(rows, idx, n_idx, sink);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; it++) gather_kernel<ROW_BYTES, U><<<grid, 256>>>(rows, idx, n_idx, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("row %4d B  U=%2d rows/warp in flight  %d CTA/SM  %7.3f ms  %7.1f GB/s gathered  (%.0f KB in flight per SM)\n", ROW_BYTES, U,
           ctas_per_sm, ms, (double)n_idx * ROW_BYTES / ms / 1e6, (double)U * 8 * ctas_per_sm * ROW_BYTES / 1024.0);
}

int main() {
    const int64_t n_rows = 32000000, n_idx = 4096 * 1000;
    uint8_t* rows;
    int64_t* idx;
    uint32_t* sink;
    if (cudaMalloc(&rows, (size_t)n_rows * 1024) != cudaSuccess) return 1;
    cudaMalloc(&idx, sizeof(int64_t) * n_idx);
    cudaMalloc(&sink, 4);
    cudaMemset(rows, 1, (size_t)n_rows * 1024);
    int64_t* h = (int64_t*)malloc(sizeof(int64_t) * n_idx);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n_idx; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        h[i] = (int64_t)(s % (uint64_t)n_rows);
    }
    cudaMemcpy(idx, h, sizeof(int64_t) * n_idx, cudaMemcpyHostToDevice);
    run<1024, 1>(rows, idx, n_idx, sink, 8);
    run<1024, 2>(rows, idx, n_idx, sink, 8);
    run<1024, 4>(rows, idx, n_idx, sink, 2);
    run<1024, 4>(rows, idx, n_idx, sink, 4);
    run<1024, 4>(rows, idx, n_idx, sink, 8);
    run<1024, 8>(rows, idx, n_idx, sink, 4);
    run<1024, 8>(rows, idx, n_idx, sink, 8);
    run<1024, 16>(rows, idx, n_idx, sink, 4);
    run<128, 4>(rows, idx, n_idx, sink, 8);
    run<128, 16>(rows, idx, n_idx, sink, 8);
    run<128, 32>(rows, idx, n_idx, sink, 8);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
