// mxf4_peak.cu - measured tensor-pipe ceiling for the instruction the dense Hamming scan is built on:
//     tcgen05.mma.cta_group::{1,2}.kind::mxf4.block_scale.scale_vec::2X   (packed e2m1 operands, f32 accumulate)
// A bare issue loop: operands resident (A in tensor memory exactly as scan_mma.cu keeps its query tile, B in shared
// memory behind the same SWIZZLE_128B descriptor), no TMA, no expanders, no epilogue - one elected thread per CTA
// (pair) issues back-to-back MMAs into one or two TMEM accumulators and commits to an mbarrier once per "tile" of 16
// MMAs (K = 1024), keeping two tiles in flight like the real kernel.  The number this prints is the denominator of
// bench.py's `roofline.peak` (MEASURED_PEAKS.json only carries a cuBLAS bf16 figure).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mxf4_peak mxf4_peak.cu && ./mxf4_peak
//
// Output: one line per (cta_group, N, data pattern): TFLOP/s (2 ops per multiply-add), MMA cycles per 128 x N x 64 tile
// slice from clock64, and the SM clock the run averaged.  Patterns: "zeros" (no datapath toggling: the clock the power
// cap allows is highest), "scan" (the +-{0.5,1,2} x {0,0.5,1,2} one-hot nibbles the Hamming scan feeds), "random"
// (every nibble random: the worst case for power).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

template <int CG>
__device__ __forceinline__ void umma_f4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                        uint32_t accumulate) {
    if constexpr (CG == 2)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %6, 0;\n"
            "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %6, 0;\n"
            "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
            : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)1)
                     : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr uint32_t TMEM_COLS = 512, A_COL = 0, SFA_COL = 128, SFB_COL = 160, D_COL = 256;

// N = MMA N (database rows per tile: 128 with two accumulators as in scan_mma.cu, or 256 with one)
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) mxf4_issue_kernel(int tiles, int pattern, unsigned long long* cycles_out) {
    constexpr int MY_ROWS = N / CG;             // rows of B this CTA holds
    constexpr int KB_BYTES = MY_ROWS * 128;     // one K-block (256 e2m1 elements per row = 128 bytes)
    constexpr int NACC = N <= 128 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* b_mem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 4 K-blocks
    __shared__ unsigned long long done_bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

    // B operand bytes (the swizzle only permutes 16-byte chunks inside a row: irrelevant for throughput)
    for (int i = tid; i < 4 * KB_BYTES / 4; i += blockDim.x) {
        const uint32_t h = hash32((uint32_t)i * 2654435761u + blockIdx.x * 977u + 17u);
        uint32_t w = 0u;
        if (pattern == 1) w = (i & 3) == 3 ? ((h >> 1) & 0x44444444u) : (h & (0x11111111u << (i & 3)));
        if (pattern == 2) w = h;
        reinterpret_cast<uint32_t*>(b_mem)[i] = w;
    }
    if (tid == 0) {
        mbar_init(smem_u32(&done_bar[0]), 1);
        mbar_init(smem_u32(&done_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    {
        // A operand: 128 lanes x 128 columns (K = 1024 e2m1 elements per lane), scale factors all 1.0 (UE8M0 0x7F)
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 128; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t h = hash32((uint32_t)(tid * 131 + c + t) * 2246822519u + blockIdx.x);
                uint32_t w = 0u;
                if (pattern == 1) {
                    const uint32_t mag = (t & 3) == 0 ? 0x44444444u : ((t & 3) == 1 ? 0x22222222u : 0x11111111u);
                    w = (mag | 0x88888888u) ^ ((h & 0x11111111u) << 3);
                }
                if (pattern == 2) w = h;
                v[t] = w;
            }
            tmem_st8(lane_base + A_COL + c, v);
        }
        uint32_t one[8];
#pragma unroll
        for (int t = 0; t < 8; t++) one[t] = 0x7F7F7F7Fu;
        for (int c = 0; c < 64; c += 8) tmem_st8(lane_base + SFA_COL + c, one);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    tc_fence_after();

    if (warp == 0 && rank == 0) {
        // idesc: A = B = e2m1 (format 1), K-major, UE8M0 scales (bit 23), N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t M = CG == 2 ? 256 : 128;
        const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
        const uint64_t desc0 = umma_desc_sw128(smem_u32(b_mem));
        const uint32_t bar0 = smem_u32(&done_bar[0]);
        unsigned long long t0 = 0, t1 = 0;
        uint32_t elected;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
        if (elected) {
            t0 = clock64();
            for (int t = 0; t < tiles; t++) {
                const int as = t % NACC;
                if (t >= 2) mbar_wait(bar0 + 8 * (t & 1), ((uint32_t)(t - 2) >> 1) & 1u);  // tile t - 2 is complete
                const uint32_t d_tmem = tmem + D_COL + (uint32_t)as * 128;
#pragma unroll
                for (int kb = 0; kb < 4; kb++)
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++)
                        umma_f4<CG>(d_tmem, tmem + A_COL + (uint32_t)(kb * 4 + k4) * 8, desc0 + (uint64_t)(kb * (KB_BYTES >> 4) + k4 * 2), idesc,
                                    tmem + SFA_COL, tmem + SFB_COL, (kb | k4) != 0);
                tc_commit<CG>(bar0 + 8 * (t & 1));
            }
            for (int t = tiles > 2 ? tiles - 2 : 0; t < tiles; t++) mbar_wait(bar0 + 8 * (t & 1), ((uint32_t)t >> 1) & 1u);
            t1 = clock64();
            if (cycles_out) cycles_out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

template <int CG, int N>
void run(int sms, int tiles, int pattern, unsigned long long* cyc_dev) {
    constexpr int MY_ROWS = N / CG;
    const size_t smem = 1024 + (size_t)4 * MY_ROWS * 128 + 64;
    auto kern = mxf4_issue_kernel<CG, N>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas = CG == 2 ? (sms / 2) * 2 : sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, kern, tiles / 8, pattern, (unsigned long long*)nullptr));  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    unsigned long long cyc = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(cyc_dev, 0, sizeof(unsigned long long) * 256));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, kern, tiles, pattern, cyc_dev));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
            best = ms;
            unsigned long long h[256];
            CK(cudaMemcpy(h, cyc_dev, sizeof(h), cudaMemcpyDeviceToHost));
            cyc = 0;
            for (int i = 0; i < ctas; i++) cyc = h[i] > cyc ? h[i] : cyc;
        }
    }
    const double M = CG == 2 ? 256.0 : 128.0;
    const double groups = (double)ctas / CG;
    const double flop = 2.0 * M * N * 1024.0 * (double)tiles * groups;  // 16 MMAs of K = 64 per tile
    const char* pn = pattern == 0 ? "zeros" : (pattern == 1 ? "scan" : "random");
    printf("cta_group::%d M=%3d N=%3d K=64 e2m1 %-6s : %8.1f TFLOP/s  %.3f ms  %7.1f cycles/tile (ideal %d)  clock %.0f MHz  [%d CTAs]\n", CG, (int)M, N, pn,
           flop / (best * 1e-3) / 1e12, best, (double)cyc / tiles, N * 8, (double)cyc / (best * 1e3), ctas);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int tiles = argc > 1 ? atoi(argv[1]) : 40000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs; %d tiles of 16 MMAs per CTA (pair)\n", prop.name, prop.multiProcessorCount, tiles);
    unsigned long long* cyc_dev;
    CK(cudaMalloc(&cyc_dev, sizeof(unsigned long long) * 256));
    if (argc > 2 && !strcmp(argv[2], "smalln")) {
        // the swapped-operand kernels (database rows = M = 128, queries = N): does the MMA rate hold at small N?
        run<1, 16>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        run<1, 32>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        run<1, 48>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        run<1, 64>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        run<1, 96>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        run<1, 128>(prop.multiProcessorCount, tiles, 1, cyc_dev);
        return 0;
    }
    for (int pattern = 0; pattern < 3; pattern++) {
        run<2, 128>(prop.multiProcessorCount, tiles, pattern, cyc_dev);
        run<2, 256>(prop.multiProcessorCount, tiles, pattern, cyc_dev);
        run<1, 128>(prop.multiProcessorCount, tiles, pattern, cyc_dev);
        run<1, 256>(prop.multiProcessorCount, tiles, pattern, cyc_dev);
    }
    return 0;
}
