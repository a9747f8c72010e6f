// Micro-benchmark: issue rate of POPC / LOP3 / IADD3 on sm_100a, alone and mixed the way the Hamming scan mixes
// them.  SURVEY.md App. D assumed POPC = 16 lanes/clk/SM; "measure, don't guess".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_bench popc_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int UNROLL = 16;

template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t a[UNROLL], acc[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; acc[i] = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (MODE == 0) {            // POPC only (dependent chain per slot, UNROLL independent slots)
                asm volatile("popc.b32 %0, %1;" : "=r"(a[i]) : "r"(a[i]));
            } else if (MODE == 1) {     // LOP3 only
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(seed), "r"(it));
            } else if (MODE == 2) {     // XOR + POPC + ADD: the naive scan inner op
                uint32_t x;
                asm volatile("xor.b32 %0, %1, %2;" : "=r"(x) : "r"(a[i]), "r"(seed + it));
                uint32_t p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x));
                acc[i] += p;
            } else if (MODE == 3) {     // IADD only
                asm volatile("add.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(seed));
            } else if (MODE == 4) {     // 1 POPC : 4 LOP3 (CSA-heavy mix)
                uint32_t x = a[i];
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x) : "r"(x), "r"(seed), "r"(it));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(x) : "r"(x), "r"(seed), "r"(it));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x) : "r"(x), "r"(acc[i]), "r"(it));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(a[i]) : "r"(x), "r"(seed), "r"(it));
                uint32_t p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x));
                acc[i] += p;
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++) s += a[i] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_slot, int threads) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    k<MODE><<<sms, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double slots = (double)ITERS * UNROLL * threads;  // per SM
    printf("%-28s threads/SM=%4d  cycles=%9.0f  slot-ops/clk/SM=%7.2f  (x%d instr => %7.2f lane-instr/clk/SM)  %.3f ms  eff.clock=%.0f MHz\n",
           name, threads, avg, slots / avg, ops_per_slot, slots * ops_per_slot / avg, ms, avg / ms / 1e3);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("POPC only", 1, threads);
        run<1>("LOP3 only", 1, threads);
        run<3>("IADD only", 1, threads);
        run<2>("XOR+POPC+ADD", 3, threads);
        run<4>("4xLOP3+POPC+ADD", 6, threads);
    }
    return 0;
}
