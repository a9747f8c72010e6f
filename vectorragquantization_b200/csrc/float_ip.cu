// Brute-force float32 inner-product top-k: the faiss.IndexIDMap(faiss.IndexFlatIP(d)) the reference's float baseline class
// uses (CohereVectorDBFloat.py:62 ctor, :156 search, :133 add_with_ids) - the recall yardstick the quantised classes are
// compared with (main.py).  Rows are the VRQ_PAYLOAD_F32 payload of a vrq_index.
//
//   scores[q][r] = sum_j q[j] * x[r][j]            float32 accumulation (faiss: BLAS sgemm / SIMD loops, order unspecified)
//   result       = the k largest scores per query, descending; ties by ascending position
//
// Two kernels per batch of <= 8 queries: (1) one warp per database row streams the row once (coalesced float4 loads) and
// produces its dot product with each query of the batch held in shared memory - HBM-bound, 4 d bytes per row per batch;
// (2) one CTA per query radix-selects the k largest of the n scores (as 64-bit keys ~ordered(score) << 32 | position, the
// selection machinery of the Hamming scan) and bitonic-sorts them.
#include <math.h>

#include <algorithm>

#include "topk_utils.cuh"
#include "vrq_internal.cuh"

namespace {

using namespace vrq;
constexpr int IP_QB = 8;         // queries per pass over the rows
constexpr int IP_THREADS = 256;  // 8 warps = 8 rows in flight per block
constexpr int IPK_THREADS = 512;
constexpr int IP_MAX_K = 4096;

__global__ void __launch_bounds__(IP_THREADS) ip_scores_kernel(const float* __restrict__ rows, int64_t n, int d, const float* __restrict__ q,
                                                              int nqb, float* __restrict__ scores /* [nqb][n] */) {
    extern __shared__ float qs[];  // [IP_QB][d]
    for (int i = threadIdx.x; i < IP_QB * d; i += IP_THREADS) qs[i] = (i / d) < nqb ? q[i] : 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d4 = d >> 2;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < n; r += (int64_t)gridDim.x * 8) {
        const float4* x = reinterpret_cast<const float4*>(rows + (size_t)r * d);
        float acc[IP_QB];
#pragma unroll
        for (int b = 0; b < IP_QB; b++) acc[b] = 0.f;
        for (int c = lane; c < d4; c += 32) {
            const float4 v = __ldg(x + c);
#pragma unroll
            for (int b = 0; b < IP_QB; b++) {
                const float4 w = *reinterpret_cast<const float4*>(qs + b * d + 4 * c);
                acc[b] = fmaf(v.x, w.x, acc[b]);
                acc[b] = fmaf(v.y, w.y, acc[b]);
                acc[b] = fmaf(v.z, w.z, acc[b]);
                acc[b] = fmaf(v.w, w.w, acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < IP_QB; b++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
        }
        if (lane < nqb) {
            float v = acc[0];
#pragma unroll
            for (int b = 1; b < IP_QB; b++)
                if (lane == b) v = acc[b];
            scores[(size_t)lane * n + r] = v;
        }
    }
}

// key: larger score -> smaller key; equal scores -> lower position first
__device__ __forceinline__ unsigned long long ip_key(float s, int64_t r) {
    return ((0xFFFFFFFFull - ordered_from_float(s)) << 32) | (unsigned long long)(uint32_t)r;
}

__global__ void __launch_bounds__(IPK_THREADS) ip_topk_kernel(const float* __restrict__ scores, int64_t n, int k, const int64_t* __restrict__ id_map,
                                                              int64_t id0, float* __restrict__ out_scores, int64_t* __restrict__ out_labels) {
    extern __shared__ unsigned long long sel[];  // next_pow2(k)
    __shared__ SelectScratch sc;
    const int q = blockIdx.x, tid = threadIdx.x;
    const float* s = scores + (size_t)q * n;
    const int n2 = next_pow2(k);
    for (int i = tid; i < n2; i += IPK_THREADS) sel[i] = ~0ull;
    unsigned long long kth = ~0ull;
    if (n > k) kth = radix_select_kth<IPK_THREADS>([&](int i) { return ip_key(s[i], i); }, (int)n, k, tid, &sc, 0);
    if (tid == 0) sc.counter = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += IPK_THREADS) {
        const unsigned long long key = ip_key(s[i], i);
        if (key <= kth) sel[atomicAdd(&sc.counter, 1)] = key;
    }
    bitonic_sort<IPK_THREADS, false>(sel, nullptr, n2, tid, 0);
    for (int i = tid; i < k; i += IPK_THREADS) {
        const unsigned long long key = sel[i];
        const size_t o = (size_t)q * k + i;
        if (key == ~0ull) {  // fewer than k rows: faiss pads with label -1 and the lowest score
            out_labels[o] = -1;
            out_scores[o] = -INFINITY;
        } else {
            const int64_t r = (int64_t)(key & 0xFFFFFFFFull);
            out_labels[o] = id_map ? id_map[r] : id0 + r;
            out_scores[o] = s[r];
        }
    }
}

}  // namespace

int vrq_launch_ip_topk(vrq_ctx* ctx, const float* rows, int64_t n, int d, const float* q, int64_t nq, int k, const int64_t* id_map,
                       int64_t id0, float* out_scores, int64_t* out_labels, cudaStream_t st) {
    if (nq == 0) return 0;
    if (k <= 0 || k > IP_MAX_K) {
        vrq_set_error("float inner-product top-k supports 1 <= k <= %d (got %d)", IP_MAX_K, k);
        return k <= 0 ? VRQ_ERR_ARG : VRQ_ERR_UNSUPPORTED;
    }
    if (d % 4 != 0 || d > 4096 || n >= (1ll << 31)) {
        vrq_set_error("float inner-product search needs d %% 4 == 0, d <= 4096 and fewer than 2^31 rows");
        return VRQ_ERR_UNSUPPORTED;
    }
    if (n == 0) {
        VRQ_CUDA(cudaMemsetAsync(out_labels, 0xFF, sizeof(int64_t) * (size_t)nq * k, st));
        return 0;
    }
    void* sc_v;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_IP_SCORES, sizeof(float) * (size_t)IP_QB * (size_t)n, &sc_v));
    const size_t smem_a = sizeof(float) * (size_t)IP_QB * d;
    VRQ_CUDA(cudaFuncSetAttribute(ip_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_a, 48 * 1024)));
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 8);
    for (int64_t q0 = 0; q0 < nq; q0 += IP_QB) {
        const int nqb = (int)std::min<int64_t>(IP_QB, nq - q0);
        ip_scores_kernel<<<(unsigned)blocks, IP_THREADS, smem_a, st>>>(rows, n, d, q + (size_t)q0 * d, nqb, (float*)sc_v);
        ip_topk_kernel<<<nqb, IPK_THREADS, sizeof(unsigned long long) * (size_t)n2, st>>>((const float*)sc_v, n, k, id_map, id0,
                                                                                           out_scores + (size_t)q0 * k, out_labels + (size_t)q0 * k);
        vrq_count_launch(ctx, 2);
        VRQ_CUDA(cudaGetLastError());
    }
    return 0;
}
