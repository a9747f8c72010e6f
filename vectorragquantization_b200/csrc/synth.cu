// Counter-based synthetic embeddings (there is no network, so no Cohere / Ollama vectors).  Integer-exact and
// identical to the test oracle's generator (oracle/): every element is a pure function of
// (seed, row, column), so any shard on any GPU - and the CPU oracle - regenerates the same database.
#include "vrq_internal.cuh"

namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float synth_elem(uint64_t base, uint64_t r, uint32_t c, int d) {
    const uint64_t h = splitmix64(base + r * (uint64_t)d + c);
    const int s = (int)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48));
    const int mc = (int)(splitmix64((uint64_t)c ^ 0xC01DBEEFCAFEF00Dull) & 0x7FFF) - 16384;
    return __fmul_rn((float)(s - 131070 + mc), 0x1p-20f);
}

__global__ void __launch_bounds__(256) synth_f32_kernel(uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale,
                                                        float* __restrict__ out) {
    const uint64_t base = seed * 0xD1342543DE82EF95ull;
    const int64_t total = nrows * (int64_t)d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const uint32_t c = (uint32_t)(e - r * d);
        const uint64_t rr = (uint64_t)(row0 + r);
        float v = synth_elem(base, rr, c, d);
        if (row_scale) {
            const uint64_t hr = splitmix64(base ^ (rr + 0x5851F42D4C957F2Dull));
            v = __fmul_rn(v, __int_as_float((127 + (int)(hr & 3) - 1) << 23));  // * 2^((hr&3)-1), exact
        }
        out[e] = v;
    }
}

// One thread per output byte of the code: 8 elements -> 1 code byte and 8 int8 values (two 32-bit stores).
__global__ void __launch_bounds__(256) synth_codes_int8_kernel(uint64_t seed, int64_t row0, int64_t nrows, int d,
                                                               uint8_t* __restrict__ codes, int8_t* __restrict__ i8) {
    const uint64_t base = seed * 0xD1342543DE82EF95ull;
    const int nb = d >> 3;
    const int64_t total = nrows * (int64_t)nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / nb;
        const int b = (int)(t - r * nb);
        const uint64_t rr = (uint64_t)(row0 + r);
        uint32_t byte = 0, lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float x = synth_elem(base, rr, (uint32_t)(8 * b + j), d);
            byte = (byte << 1) | (uint32_t)(x > 0.0f);
            float v = __fsub_rn(__fmul_rn(x, 1259.0f), 0.69f);
            v = fminf(fmaxf(rintf(v), -128.f), 127.f);
            const uint32_t q = (uint32_t)(__float2int_rz(v) & 0xFF);
            if (j < 4)
                lo |= q << (8 * j);
            else
                hi |= q << (8 * (j - 4));
        }
        if (codes) codes[t] = (uint8_t)byte;
        if (i8) reinterpret_cast<uint2*>(i8)[t] = make_uint2(lo, hi);
    }
}

__global__ void iota_i64_kernel(int64_t* out, int64_t n, int64_t start) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = start + i;
}

int grid_for(vrq_ctx* ctx, int64_t work) {
    int64_t blocks = (work + 255) / 256, cap = (int64_t)ctx->sm_count * 32;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

int vrq_launch_synth_f32(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale, float* out,
                         cudaStream_t st) {
    if (nrows == 0) return 0;
    synth_f32_kernel<<<grid_for(ctx, nrows * (int64_t)d), 256, 0, st>>>(seed, row0, nrows, d, row_scale, out);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_synth_codes_int8(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, uint8_t* codes,
                                int8_t* i8, cudaStream_t st) {
    if (nrows == 0) return 0;
    synth_codes_int8_kernel<<<grid_for(ctx, nrows * (int64_t)(d / 8)), 256, 0, st>>>(seed, row0, nrows, d, codes, i8);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_iota_i64(vrq_ctx* ctx, int64_t* out, int64_t n, int64_t start, cudaStream_t st) {
    if (n == 0) return 0;
    iota_i64_kernel<<<grid_for(ctx, n), 256, 0, st>>>(out, n, start);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}
