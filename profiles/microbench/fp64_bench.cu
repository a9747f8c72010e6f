// Micro-benchmark: DFMA / DADD issue rate on sm_100a (B200), to size the float64 rescoring kernels.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, UNROLL = 8;
template <int MODE>
__global__ void k(double* out, double seed, long long* cycles) {
    double a[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++) a[i] = seed * (threadIdx.x + 1 + i);
    double b = seed * 1.0000001, c = seed * 0.5;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
            else asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(c));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * threads); cudaMalloc(&cyc, sizeof(long long) * sms);
    k<MODE><<<sms, threads>>>(out, 1.5, cyc); cudaDeviceSynchronize();
    k<MODE><<<sms, threads>>>(out, 1.5, cyc); cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    printf("%-6s threads/SM=%4d cycles=%9.0f lane-ops/clk/SM=%6.2f\n", name, threads, avg, (double)ITERS * UNROLL * threads / avg);
    cudaFree(out); cudaFree(cyc);
}
int main() { for (int t : {128, 256, 512, 1024}) { run<0>("DFMA", t); run<1>("DADD", t); } return 0; }
