// Phase II / III rescoring, the 2-phase payload rescoring of the VectorDB* classes, and the per-query
// select/sort kernels that replace the reference's Python list.sort calls.
//
// Reference lines: Phase II CohereEnhancedVectorDB.py:283-297, Phase III :302-322, 2-phase VectorDBInt8.py:226-242.
// Every sort here reproduces a STABLE descending list.sort: the sort key is (score descending, previous rank
// ascending), which is a total order, so the result is unique and equals the stable sort.
#include <math.h>

#include <algorithm>
#include <stdlib.h>

#include "topk_utils.cuh"
#include "vrq_internal.cuh"

#ifndef VRQ_RESCORE_IMMA_DEFAULT
#define VRQ_RESCORE_IMMA_DEFAULT 1  // cfg5 measurement (profiles/r02): the tensor-core path beats the CUDA-core path
#endif
#ifndef VRQ_RESCORE_BIN_DEFAULT
#define VRQ_RESCORE_BIN_DEFAULT 2
#endif

namespace {

using namespace vrq;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ int64_t cand_row(const uint64_t* keys, const int64_t* pos, size_t idx, int64_t pos_base) {
    if (keys) {
        uint64_t k = keys[idx];
        return k == VRQ_KEY_NONE ? -1 : (int64_t)(k & VRQ_KEY_POS_MASK) - pos_base;
    }
    int64_t p = pos[idx];
    return p < 0 ? -1 : p - pos_base;
}

// ---- Phase II: sum_i q[i] * (2*bit_i - 1), float64 accumulation ------------------------------------------------
// One warp per (query, candidate).  Lane l owns dimensions {l, l+32, l+64, ...}: the query values it needs sit in
// consecutive shared-memory words (conflict-free, already widened to float64 once per CTA) and its bit of code word t
// is always bit 8*(l/8) + 7 - (l%8) (np.packbits is MSB-first inside each byte).  The code is never unpacked to
// memory: the bit only flips the sign bit of the float64 addend.
__global__ void __launch_bounds__(256) rescore_binary_kernel(const uint8_t* __restrict__ codes, int d,
                                                             const uint64_t* __restrict__ keys,
                                                             const int64_t* __restrict__ pos, int64_t pos_base, int m,
                                                             const float* __restrict__ qf, double* __restrict__ score) {
    extern __shared__ double qd[];
    const int q = blockIdx.y;
    for (int i = threadIdx.x; i < d; i += blockDim.x) qd[i] = (double)qf[(size_t)q * d + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words = d >> 5;
    const int bitpos = 8 * (lane >> 3) + 7 - (lane & 7);
    for (int i = blockIdx.x * 8 + warp; i < m; i += gridDim.x * 8) {
        const size_t idx = (size_t)q * m + i;
        const int64_t row = cand_row(keys, pos, idx, pos_base);
        if (row < 0) {
            if (lane == 0) score[idx] = -INFINITY;
            continue;
        }
        const uint32_t* code = reinterpret_cast<const uint32_t*>(codes + (size_t)row * (d >> 3));
        // lane l fetches word l (+32, ...) once, coalesced; every lane then needs every word -> shuffle broadcast
        double acc = 0.0;
        for (int t0 = 0; t0 < words; t0 += 32) {
            const uint32_t mine = (t0 + lane < words) ? __ldg(code + t0 + lane) : 0u;
            const int nt = min(32, words - t0);
            for (int t = 0; t < nt; t++) {
                const uint32_t w = __shfl_sync(FULL, mine, t);
                const double v = qd[32 * (t0 + t) + lane];
                const int hi = __double2hiint(v) ^ (int)((~(w >> bitpos) & 1u) << 31);  // bit 0 -> negate
                acc += __hiloint2double(hi, __double2loint(v));
            }
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) score[idx] = acc;
    }
}

// d % 32 != 0 (faiss only asks for d % 8 == 0): code rows are not word-aligned; lane l walks dimensions l, l + 32, ... and
// picks its bit out of the code byte by byte.  Correct, not tuned - such widths are not on any benchmark path.
__global__ void __launch_bounds__(256) rescore_binary_bytes_kernel(const uint8_t* __restrict__ codes, int d,
                                                                   const uint64_t* __restrict__ keys, const int64_t* __restrict__ pos,
                                                                   int64_t pos_base, int m, const float* __restrict__ qf,
                                                                   double* __restrict__ score) {
    extern __shared__ double qd[];
    const int q = blockIdx.y;
    for (int i = threadIdx.x; i < d; i += blockDim.x) qd[i] = (double)qf[(size_t)q * d + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * 8 + warp; i < m; i += gridDim.x * 8) {
        const size_t idx = (size_t)q * m + i;
        const int64_t row = cand_row(keys, pos, idx, pos_base);
        if (row < 0) {
            if (lane == 0) score[idx] = -INFINITY;
            continue;
        }
        const uint8_t* code = codes + (size_t)row * (d >> 3);
        double acc = 0.0;
        for (int e = lane; e < d; e += 32) acc += ((__ldg(code + (e >> 3)) >> (7 - (e & 7))) & 1) ? qd[e] : -qd[e];
        acc = warp_sum_f64(acc);
        if (lane == 0) score[idx] = acc;
    }
}

// d == 1024: lane l owns code word l, i.e. the 32 dimensions 32 l .. 32 l + 31, with their query values resident in
// registers as float64 (bit 8 c + 7 - i of the little-endian word is dimension 32 l + 8 c + i: np.packbits is MSB-first
// inside each byte).  No shared memory and no shuffles in the inner loop: per dimension one funnel shift, one LOP3 that
// folds the (inverted) bit into the sign of the addend, one DADD; four candidates are in flight per warp.
__global__ void __launch_bounds__(256, 2) rescore_binary1024_kernel(const uint8_t* __restrict__ codes,
                                                                    const uint64_t* __restrict__ keys,
                                                                    const int64_t* __restrict__ pos, int64_t pos_base, int m,
                                                                    const float* __restrict__ qf, double* __restrict__ score) {
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int qhi[32], qlo[32];  // the 32 query values of this lane as float64 halves
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
        const float4 f = *reinterpret_cast<const float4*>(qf + (size_t)q * 1024 + 32 * lane + e);
        const double d0 = (double)f.x, d1 = (double)f.y, d2 = (double)f.z, d3 = (double)f.w;
        qhi[e] = __double2hiint(d0), qlo[e] = __double2loint(d0);
        qhi[e + 1] = __double2hiint(d1), qlo[e + 1] = __double2loint(d1);
        qhi[e + 2] = __double2hiint(d2), qlo[e + 2] = __double2loint(d2);
        qhi[e + 3] = __double2hiint(d3), qlo[e + 3] = __double2loint(d3);
    }
    // Candidates are processed in chunks of 32 per warp: lane j fetches the position of candidate j (one load per
    // chunk), the code words of candidates j .. j + 7 are in flight in a register ring (one register per slot), and lane j
    // keeps the score of candidate j so that the chunk's 32 scores leave in one store.
    constexpr int R = 8;
    const int step = gridDim.x * 8, first = blockIdx.x * 8 + warp;
    const int ncand = first < m ? (m - first + step - 1) / step : 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int nchunk = min(32, ncand - c0);
        const size_t myidx = (size_t)q * m + first + (size_t)(c0 + lane) * step;
        const int64_t myrow = (lane < nchunk) ? cand_row(keys, pos, myidx, pos_base) : -1;
        double my_acc = 0.0;
        uint32_t ring[R];
#pragma unroll
        for (int p = 0; p < R; p++) {
            const int64_t r = __shfl_sync(FULL, myrow, p);
            ring[p] = (p < nchunk && r >= 0) ? ~__ldg(reinterpret_cast<const uint32_t*>(codes + (size_t)r * 128) + lane) : 0u;
        }
#pragma unroll 1
        for (int j = 0; j < nchunk; j++) {
            const uint32_t nw = ring[0];
#pragma unroll
            for (int p = 0; p < R - 1; p++) ring[p] = ring[p + 1];
            const int64_t rn = __shfl_sync(FULL, myrow, (j + R) & 31);
            ring[R - 1] = (j + R < nchunk && rn >= 0) ? ~__ldg(reinterpret_cast<const uint32_t*>(codes + (size_t)rn * 128) + lane) : 0u;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int b = 0; b < 32; b += 4) {
                // bit b of the word <-> dimension offset 8 (b / 8) + 7 - (b % 8); a CLEAR bit negates the addend
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int bit = b + t, e = 8 * (bit >> 3) + 7 - (bit & 7);
                    const int hi = qhi[e] ^ (int)((nw << (31 - bit)) & 0x80000000u);
                    const double v = __hiloint2double(hi, qlo[e]);
                    if (t == 0) a0 += v;
                    if (t == 1) a1 += v;
                    if (t == 2) a2 += v;
                    if (t == 3) a3 += v;
                }
            }
            const double acc = warp_sum_f64((a0 + a1) + (a2 + a3));
            if (lane == j) my_acc = acc;
        }
        if (lane < nchunk) score[myidx] = myrow < 0 ? -INFINITY : my_acc;
    }
}

// ---- Phase III: dot(q, int8 row) / ||row||, -inf when the norm is 0 ----------------------------------------------
// The reference does the dot in float32 (BLAS sdot, order unspecified); here it is accumulated in float64 (every
// product is exact in float64), which is within the 1e-5 parity tolerance and closer to the true value.
// HBM-gather-bound: 1 KB per (query, candidate), no operand reuse.  To stay under the gather time the inner loop has
// no int->float conversion: byte b (offset by 128) is dropped into the mantissa of the double 4096 + (b + 128) with
// one PRMT, so each element costs one PRMT + one DFMA; the offset is removed once per candidate with sum(q).
template <bool D1024>
__global__ void __launch_bounds__(256, 2) rescore_int8cos_kernel(const int8_t* __restrict__ rows, int d,
                                                              const uint64_t* __restrict__ keys,
                                                              const int64_t* __restrict__ pos, int64_t pos_base, int m,
                                                              const float* __restrict__ qf, double* __restrict__ score) {
    extern __shared__ double qd[];  // generic path only
    __shared__ double qsum_s[8];
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // lane-private query slice, float64: for d == 1024 the 32 elements of the two 16-byte chunks this lane loads
    double qr[D1024 ? 32 : 1];
    double qsum = 0.0;
    if (D1024) {
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                // opaque conversion: keeps the 32 float64 values resident in registers (otherwise the compiler
                // re-converts the float32 inside the loop and the kernel becomes bound by the conversion unit)
                const float f = qf[(size_t)q * 1024 + 512 * h + 16 * lane + e];
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(qr[16 * h + e]) : "f"(f));
                qsum += qr[16 * h + e];
            }
    } else {
        for (int i = threadIdx.x; i < d; i += blockDim.x) qd[i] = (double)qf[(size_t)q * d + i];
        __syncthreads();
        for (int i = lane; i < d; i += 32) qsum += qd[i];
    }
    qsum = warp_sum_f64(qsum);
    (void)qsum_s;
    const double offset = 4224.0 * qsum;  // sum_i q_i * (4096 + 128)
    if (D1024) {
        // Per candidate the dependent chain is position -> row address -> 1 KB row -> 8-deep DFMA chain -> shuffle tree
        // -> sqrt -> divide.  Candidates are processed in chunks of 32: lane j fetches the position of candidate j
        // (one coalesced load per chunk), rows are double-buffered in registers, and lane j keeps the reduced
        // (dot, sum of squares) of candidate j so that the 32 sqrt + divide + store run once per chunk, in parallel.
        const int step = gridDim.x * 8, first = blockIdx.x * 8 + warp;
        const int ncand = first < m ? (m - first + step - 1) / step : 0;
        for (int c0 = 0; c0 < ncand; c0 += 32) {
            const int nchunk = min(32, ncand - c0);
            const size_t myidx = (size_t)q * m + first + (size_t)(c0 + lane) * step;
            const int64_t myrow = (lane < nchunk) ? cand_row(keys, pos, myidx, pos_base) : -1;
            double my_acc = 0.0;
            int my_n2 = 0;
            // register ring: the rows of candidates j .. j+3 are in flight (4 KB per warp, ~64 KB per SM)
            uint4 ra[4], rb[4];
            int64_t rr[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                ra[p] = make_uint4(0, 0, 0, 0);
                rb[p] = ra[p];
                rr[p] = __shfl_sync(FULL, myrow, p);
                if (p < nchunk && rr[p] >= 0) {
                    const uint4* src = reinterpret_cast<const uint4*>(rows + (size_t)rr[p] * 1024);
                    ra[p] = __ldg(src + lane);
                    rb[p] = __ldg(src + 32 + lane);
                }
            }
#pragma unroll 1
            for (int j = 0; j < nchunk; j++) {
                const uint4 v0 = ra[0], v1 = rb[0];
                const int64_t row = rr[0];
#pragma unroll
                for (int p = 0; p < 3; p++) {
                    ra[p] = ra[p + 1];
                    rb[p] = rb[p + 1];
                    rr[p] = rr[p + 1];
                }
                rr[3] = __shfl_sync(FULL, myrow, (j + 4) & 31);
                if (j + 4 < nchunk && rr[3] >= 0) {
                    const uint4* src = reinterpret_cast<const uint4*>(rows + (size_t)rr[3] * 1024);
                    ra[3] = __ldg(src + lane);
                    rb[3] = __ldg(src + 32 + lane);
                }
                if (row >= 0) {
                    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                    int n2 = 0;
                    const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        n2 = __dp4a((int)w0[c], (int)w0[c], n2);
                        n2 = __dp4a((int)w1[c], (int)w1[c], n2);
                        const uint32_t u0 = w0[c] ^ 0x80808080u, u1 = w1[c] ^ 0x80808080u;  // bytes + 128, unsigned
#pragma unroll
                        for (int b = 0; b < 4; b += 2) {
                            // hi word 0x40B0_uu00: the double 4096 + u
                            acc0 = fma(qr[4 * c + b], __hiloint2double((int)__byte_perm(0x40B00000u, u0, 0x3200 | ((4 + b) << 4)), 0), acc0);
                            acc1 = fma(qr[16 + 4 * c + b], __hiloint2double((int)__byte_perm(0x40B00000u, u1, 0x3200 | ((4 + b) << 4)), 0), acc1);
                            acc2 = fma(qr[4 * c + b + 1], __hiloint2double((int)__byte_perm(0x40B00000u, u0, 0x3200 | ((5 + b) << 4)), 0), acc2);
                            acc3 = fma(qr[16 + 4 * c + b + 1], __hiloint2double((int)__byte_perm(0x40B00000u, u1, 0x3200 | ((5 + b) << 4)), 0), acc3);
                        }
                    }
                    const double acc = warp_sum_f64((acc0 + acc1) + (acc2 + acc3));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(FULL, n2, o);
                    if (lane == j) {
                        my_acc = acc;
                        my_n2 = n2;
                    }
                }
            }
            if (lane < nchunk)
                score[myidx] = (myrow < 0 || my_n2 == 0) ? -INFINITY : (my_acc - offset) / sqrt((double)my_n2);
        }
    } else {
        for (int i = blockIdx.x * 8 + warp; i < m; i += gridDim.x * 8) {
            const size_t idx = (size_t)q * m + i;
            const int64_t row = cand_row(keys, pos, idx, pos_base);
            if (row < 0) {
                if (lane == 0) score[idx] = -INFINITY;
                continue;
            }
            double acc = 0.0;
            int n2 = 0;
            const int* src = reinterpret_cast<const int*>(rows + (size_t)row * d);
            for (int w = lane; w < (d >> 2); w += 32) {
                const uint32_t x = (uint32_t)__ldg(src + w);
                n2 = __dp4a((int)x, (int)x, n2);
                const uint32_t u = x ^ 0x80808080u;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int hi = (int)(0x40B00000u | (((u >> (8 * b)) & 0xFFu) << 8));
                    acc = fma(qd[4 * w + b], __hiloint2double(hi, 0), acc);
                }
            }
            acc = warp_sum_f64(acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(FULL, n2, o);
            if (lane == 0) score[idx] = (n2 == 0) ? -INFINITY : (acc - offset) / sqrt((double)n2);
        }
    }
}

// ---- Phase III, d == 1024, rows in flight parked in shared memory ---------------------------------------------------
// Random 1 KB rows stream at the full HBM rate only with >= 128 KB in flight per SM (profiles/microbench/
// gather_bench_r01.txt: 6.8 TB/s; 4.4 TB/s at 64 KB).  The register ring of the kernel above holds 64 KB per SM and cannot
// grow (the float64 query slice takes 64 registers).  Here every lane parks its own 2 x 16 bytes of each row in shared
// memory with cp.async (LDGSTS) - a ring of P3_RING rows per warp, 8 KB - and reads back exactly the bytes it copied, so
// the ring needs no barrier at all: cp.async.wait_group orders a lane's own copies.  Same arithmetic as above.
constexpr int P3_RING = 8;

// BATCH (opt-in, VRQ_RESCORE_BATCHRED=1; held to 1e-13 of the float64 evaluation with the other variants by
// tests/test_gpu_kernels.py::test_rescore_int8cos_kernel_variants; superseded as the default by the tensor-core kernel of
// rescore_mma.cu): the per-row butterfly
// reductions (15 shuffles + 10 adds per row, the top stall of the ncu capture) are replaced by a transpose through shared
// memory - every lane parks its partial dot / partial norm of row j in red[j & 7][lane]; after 8 rows lane (qd, r) adds
// the 8 partials of quarter qd of row r (columns visited in the skewed order 8 qd + ((i + r) & 7): conflict-free), two
// xor-shuffle steps close the sum over the quarters, and lane j picks up the total of row j.
template <bool BATCH>
__global__ void __launch_bounds__(256, 2) rescore_int8cos_async_kernel(const int8_t* __restrict__ rows, const uint64_t* __restrict__ keys,
                                                                       const int64_t* __restrict__ pos, int64_t pos_base, int m,
                                                                       const float* __restrict__ qf, double* __restrict__ score) {
    extern __shared__ __align__(16) uint8_t p3_ring[];  // [8 warps][P3_RING][2][32 lanes][16 bytes]
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t my_slot0 = (uint32_t)__cvta_generic_to_shared(p3_ring) + (uint32_t)(warp * P3_RING * 1024 + lane * 16);
    // BATCH: [8 warps][8 rows][32 lanes] doubles, then the same shape of ints, behind the ring
    double* red_acc = reinterpret_cast<double*>(p3_ring + 8 * P3_RING * 1024) + warp * 256;
    int* red_n2 = reinterpret_cast<int*>(p3_ring + 8 * P3_RING * 1024 + 8 * 256 * sizeof(double)) + warp * 256;
    double qr[32];
    double qsum = 0.0;
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const float f = qf[(size_t)q * 1024 + 512 * h + 16 * lane + e];
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(qr[16 * h + e]) : "f"(f));
            qsum += qr[16 * h + e];
        }
    qsum = warp_sum_f64(qsum);
    const double offset = 4224.0 * qsum;  // sum_i q_i * (4096 + 128)
    const int step = gridDim.x * 8, first = blockIdx.x * 8 + warp;
    const int ncand = first < m ? (m - first + step - 1) / step : 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int nchunk = min(32, ncand - c0);
        const size_t myidx = (size_t)q * m + first + (size_t)(c0 + lane) * step;
        const int64_t myrow = (lane < nchunk) ? cand_row(keys, pos, myidx, pos_base) : -1;
        double my_acc = 0.0;
        int my_n2 = 0;
        auto issue = [&](int j) {  // park row j of this chunk in ring slot j % P3_RING (one commit group per row, even if empty)
            const int64_t r = __shfl_sync(FULL, myrow, j & 31);
            if (j < nchunk && r >= 0) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(rows) + (size_t)r * 1024 + lane * 16;
                const uint32_t dst = my_slot0 + (uint32_t)((j % P3_RING) * 1024);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 512) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int j = 0; j < P3_RING; j++) issue(j);
#pragma unroll 1
        for (int j = 0; j < nchunk; j++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(P3_RING - 1) : "memory");
            const int64_t row = __shfl_sync(FULL, myrow, j);
            const uint32_t src = my_slot0 + (uint32_t)((j % P3_RING) * 1024);
            uint4 v0, v1;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w) : "r"(src) : "memory");
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "r"(src + 512) : "memory");
            if (row >= 0) {
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                int n2 = 0;
                const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    n2 = __dp4a((int)w0[c], (int)w0[c], n2);
                    n2 = __dp4a((int)w1[c], (int)w1[c], n2);
                    const uint32_t u0 = w0[c] ^ 0x80808080u, u1 = w1[c] ^ 0x80808080u;  // bytes + 128, unsigned
#pragma unroll
                    for (int b = 0; b < 4; b += 2) {
                        acc0 = fma(qr[4 * c + b], __hiloint2double((int)__byte_perm(0x40B00000u, u0, 0x3200 | ((4 + b) << 4)), 0), acc0);
                        acc1 = fma(qr[16 + 4 * c + b], __hiloint2double((int)__byte_perm(0x40B00000u, u1, 0x3200 | ((4 + b) << 4)), 0), acc1);
                        acc2 = fma(qr[4 * c + b + 1], __hiloint2double((int)__byte_perm(0x40B00000u, u0, 0x3200 | ((5 + b) << 4)), 0), acc2);
                        acc3 = fma(qr[16 + 4 * c + b + 1], __hiloint2double((int)__byte_perm(0x40B00000u, u1, 0x3200 | ((5 + b) << 4)), 0), acc3);
                    }
                }
                if (BATCH) {
                    red_acc[(j & 7) * 32 + lane] = (acc0 + acc1) + (acc2 + acc3);
                    red_n2[(j & 7) * 32 + lane] = n2;
                } else {
                    const double acc = warp_sum_f64((acc0 + acc1) + (acc2 + acc3));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(FULL, n2, o);
                    if (lane == j) {
                        my_acc = acc;
                        my_n2 = n2;
                    }
                }
            } else if (BATCH) {
                red_acc[(j & 7) * 32 + lane] = 0.0;
                red_n2[(j & 7) * 32 + lane] = 0;
            }
            if (BATCH && ((j & 7) == 7 || j == nchunk - 1)) {
                __syncwarp();
                const int r = lane & 7, qd = lane >> 3;
                double a = 0.0;
                int nn = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int col = 8 * qd + ((i + r) & 7);
                    a += red_acc[r * 32 + col];
                    nn += red_n2[r * 32 + col];
                }
                a += __shfl_xor_sync(FULL, a, 8);
                a += __shfl_xor_sync(FULL, a, 16);
                nn += __shfl_xor_sync(FULL, nn, 8);
                nn += __shfl_xor_sync(FULL, nn, 16);
                if ((lane >> 3) == (j >> 3)) {  // lanes 8 (j/8) .. +7 take the totals of rows 8 (j/8) + (lane & 7)
                    my_acc = a;
                    my_n2 = nn;
                }
                __syncwarp();
            }
            issue(j + P3_RING);  // the slot just read is free again (this lane's own reads are complete: v0 / v1 were consumed)
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (lane < nchunk)
            score[myidx] = (myrow < 0 || my_n2 == 0) ? -INFINITY : (my_acc - offset) / sqrt((double)my_n2);
    }
}

// ---- Phase III, d == 1024, integer dot products ----------------------------------------------------------------------
// Same ring as above, different arithmetic.  The float64 kernel spends ~93 of its ~180 instructions per row on 32 DFMA +
// 32 PRMT (byte -> high word of a double) + 29 moves re-zeroing operand pairs.  Here the query is turned ONCE per block
// into 64-bit fixed point against its own largest exponent E, Q_i = rint(q_i * 2^(61-E)) (|Q_i| < 2^62; exact for every
// element within 2^-38 of max|q|, i.e. far below float64 accumulation error), and split into four 16-bit limbs
// (three unsigned, the top one signed).  dp2a multiplies two limbs by two int8 bytes of the packed row word and
// accumulates in int32 - no byte extraction, no conversion: 64 IDP.2A per row and lane, sums bounded by
// 32 * 65535 * 128 < 2^31.  sum_i Q_i x_i = A3 2^48 + A2 2^32 + A1 2^16 + A0 is then an exact integer, rounded once into
// a double per lane, reduced across the warp in float64 and scaled back by 2^(E-61) (a power of two: exact).
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(256, 2) rescore_int8cos_dp2a_kernel(const int8_t* __restrict__ rows, const uint64_t* __restrict__ keys,
                                                                      const int64_t* __restrict__ pos, int64_t pos_base, int m,
                                                                      const float* __restrict__ qf, double* __restrict__ score) {
    extern __shared__ __align__(16) uint8_t p3_ring[];  // [8 warps][P3_RING][2][32 lanes][16 bytes]
    const int q = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t my_slot0 = (uint32_t)__cvta_generic_to_shared(p3_ring) + (uint32_t)(warp * P3_RING * 1024 + lane * 16);

    // ---- the query slice of this lane (elements 16 lane .. +15 and 512 + 16 lane .. +15) as packed 16-bit limbs ------
    // qa[p][l] = limb l of element 2p (low half) and of element 2p+1 (high half)
    uint32_t qa[16][4];
    double unscale;
    {
        float qv[32];
        float amax = 0.f;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                qv[16 * h + e] = qf[(size_t)q * 1024 + 512 * h + 16 * lane + e];
                amax = fmaxf(amax, fabsf(qv[16 * h + e]));
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL, amax, o));
        const int bexp = (__float_as_int(amax) >> 23) & 0xFF;  // biased exponent of max|q|; 0: the query is all zero / denormal
        const int E = bexp - 127;
        const double sc = bexp == 0 ? 0.0 : __hiloint2double((1023 + 61 - E) << 20, 0);       // 2^(61-E)
        unscale = bexp == 0 ? 0.0 : __hiloint2double((1023 - 61 + E) << 20, 0);               // 2^(E-61)
#pragma unroll
        for (int pr = 0; pr < 16; pr++) {
            const long long Q0 = __double2ll_rn((double)qv[2 * pr] * sc), Q1 = __double2ll_rn((double)qv[2 * pr + 1] * sc);
            const uint32_t lo0 = (uint32_t)Q0, hi0 = (uint32_t)((unsigned long long)Q0 >> 32);
            const uint32_t lo1 = (uint32_t)Q1, hi1 = (uint32_t)((unsigned long long)Q1 >> 32);
            qa[pr][0] = __byte_perm(lo0, lo1, 0x5410);
            qa[pr][1] = __byte_perm(lo0, lo1, 0x7632);
            qa[pr][2] = __byte_perm(hi0, hi1, 0x5410);
            qa[pr][3] = __byte_perm(hi0, hi1, 0x7632);
        }
    }

    const int step = gridDim.x * 8, first = blockIdx.x * 8 + warp;
    const int ncand = first < m ? (m - first + step - 1) / step : 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int nchunk = min(32, ncand - c0);
        const size_t myidx = (size_t)q * m + first + (size_t)(c0 + lane) * step;
        const int64_t myrow = (lane < nchunk) ? cand_row(keys, pos, myidx, pos_base) : -1;
        double my_acc = 0.0;
        int my_n2 = 0;
        auto issue = [&](int j) {  // park row j of this chunk in ring slot j % P3_RING (one commit group per row, even if empty)
            const int64_t r = __shfl_sync(FULL, myrow, j & 31);
            if (j < nchunk && r >= 0) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(rows) + (size_t)r * 1024 + lane * 16;
                const uint32_t dst = my_slot0 + (uint32_t)((j % P3_RING) * 1024);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 512) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int j = 0; j < P3_RING; j++) issue(j);
#pragma unroll 1
        for (int j = 0; j < nchunk; j++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(P3_RING - 1) : "memory");
            const int64_t row = __shfl_sync(FULL, myrow, j);
            const uint32_t src = my_slot0 + (uint32_t)((j % P3_RING) * 1024);
            uint4 v0, v1;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w) : "r"(src) : "memory");
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "r"(src + 512) : "memory");
            if (row >= 0) {
                int a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
                int n2 = 0;
                const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    n2 = __dp4a((int)w0[c], (int)w0[c], n2);
                    n2 = __dp4a((int)w1[c], (int)w1[c], n2);
#pragma unroll
                    for (int l = 0; l < 3; l++) {
                        a0[l] = dp2a_lo_us(qa[2 * c][l], w0[c], a0[l]);
                        a0[l] = dp2a_hi_us(qa[2 * c + 1][l], w0[c], a0[l]);
                        a1[l] = dp2a_lo_us(qa[8 + 2 * c][l], w1[c], a1[l]);
                        a1[l] = dp2a_hi_us(qa[8 + 2 * c + 1][l], w1[c], a1[l]);
                    }
                    a0[3] = dp2a_lo_ss(qa[2 * c][3], w0[c], a0[3]);
                    a0[3] = dp2a_hi_ss(qa[2 * c + 1][3], w0[c], a0[3]);
                    a1[3] = dp2a_lo_ss(qa[8 + 2 * c][3], w1[c], a1[3]);
                    a1[3] = dp2a_hi_ss(qa[8 + 2 * c + 1][3], w1[c], a1[3]);
                }
                double s = (double)(a0[3] + a1[3]);
                s = fma(s, 65536.0, (double)(a0[2] + a1[2]));
                s = fma(s, 65536.0, (double)(a0[1] + a1[1]));
                s = fma(s, 65536.0, (double)(a0[0] + a1[0]));
                const double acc = warp_sum_f64(s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(FULL, n2, o);
                if (lane == j) {
                    my_acc = acc;
                    my_n2 = n2;
                }
            }
            issue(j + P3_RING);  // the slot just read is free again (this lane's own reads are complete: v0 / v1 were consumed)
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (lane < nchunk) score[myidx] = (myrow < 0 || my_n2 == 0) ? -INFINITY : (my_acc * unscale) / sqrt((double)my_n2);
    }
}

// ---- 2-phase rescoring of the VectorDB* classes: float32 dot(q, dequantised payload row) -------------------------
struct PayloadParams {
    int kind;
    const void* payload;
    const void* aux;
    float scale_f32;
    double scale_f64;
    int d;
    const uint64_t* keys;
    int64_t pos_base;
    int m;
    const float* qf;
    float* score;
};

__global__ void __launch_bounds__(256) rescore_payload_kernel(PayloadParams p) {
    extern __shared__ float qs[];
    const int q = blockIdx.y, d = p.d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = p.qf[(size_t)q * d + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = blockIdx.x * 8 + warp; i < p.m; i += gridDim.x * 8) {
        const size_t idx = (size_t)q * p.m + i;
        const int64_t row = cand_row(p.keys, nullptr, idx, p.pos_base);
        if (row < 0) {
            if (lane == 0) p.score[idx] = -INFINITY;
            continue;
        }
        float sc32 = p.scale_f32;
        double sc64 = p.scale_f64;
        bool zero = false;
        if (p.kind == VRQ_PAYLOAD_INT8_PERDOC) {
            const float lo = static_cast<const float*>(p.aux)[2 * row], hi = static_cast<const float*>(p.aux)[2 * row + 1];
            sc32 = __fdiv_rn(fmaxf(fabsf(lo), fabsf(hi)), 127.f);
            zero = (lo == hi);
        } else if (p.kind == VRQ_PAYLOAD_INT4_PERDOC) {
            const double lo = static_cast<const double*>(p.aux)[2 * row], hi = static_cast<const double*>(p.aux)[2 * row + 1];
            sc64 = fmax(fabs(lo), fabs(hi)) / 7.0;
            zero = (lo == hi);
        }
        double acc = 0.0;
        for (int e = lane; e < d; e += 32) {
            float v;
            switch (p.kind) {
                case VRQ_PAYLOAD_INT8_PERDOC:
                case VRQ_PAYLOAD_INT8_GLOBAL:
                    v = __fmul_rn((float)static_cast<const int8_t*>(p.payload)[(size_t)row * d + e], sc32);
                    break;
                case VRQ_PAYLOAD_INT16_GLOBAL:
                    v = __fmul_rn((float)static_cast<const int16_t*>(p.payload)[(size_t)row * d + e], sc32);
                    break;
                case VRQ_PAYLOAD_INT4_PERDOC:
                case VRQ_PAYLOAD_INT4_GLOBAL: {
                    const uint8_t b = static_cast<const uint8_t*>(p.payload)[(size_t)row * (d >> 1) + (e >> 1)];
                    const int nib = (e & 1) ? (b & 0xF) : (b >> 4);
                    v = (float)__dmul_rn((double)(nib - 8), sc64);
                    break;
                }
                case VRQ_PAYLOAD_F32:
                    v = static_cast<const float*>(p.payload)[(size_t)row * d + e];
                    break;
                case VRQ_PAYLOAD_CODES_PM1: {  // np.unpackbits -> where(bits == 0, -1, 1).astype(float32)
                    const uint8_t b = static_cast<const uint8_t*>(p.payload)[(size_t)row * (d >> 3) + (e >> 3)];
                    v = ((b >> (7 - (e & 7))) & 1) ? 1.f : -1.f;
                    break;
                }
                default:
                    v = 0.f;
            }
            if (zero) v = 0.f;
            acc = fma((double)qs[e], (double)v, acc);
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) p.score[idx] = (float)acc;
    }
}

// ---- keys -> (distance, label) -------------------------------------------------------------------------------------
__global__ void keys_to_dist_labels_kernel(const uint64_t* __restrict__ keys, int64_t count, int64_t pos_base,
                                           const int64_t* __restrict__ id_map, int64_t id0, int32_t* __restrict__ dist,
                                           int64_t* __restrict__ labels) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t k = keys[i];
    if (k == VRQ_KEY_NONE) {
        if (dist) dist[i] = 0x7fffffff;
        if (labels) labels[i] = -1;
        return;
    }
    if (dist) dist[i] = (int32_t)(k >> VRQ_KEY_POS_BITS);
    if (labels) {
        const int64_t lp = (int64_t)(k & VRQ_KEY_POS_MASK) - pos_base;
        labels[i] = id_map ? id_map[lp] : id0 + lp;
    }
}

// ---- the three list.sort calls of CohereEnhancedVectorDB.search, per query, + the multi-GPU merge ----------------
constexpr int M3_THREADS = 256;

__global__ void __launch_bounds__(M3_THREADS) merge3_kernel(int world, int nq, int bk, int64_t rank_stride,
                                                            const uint64_t* __restrict__ keys,
                                                            const int64_t* __restrict__ labels,
                                                            const double* __restrict__ sbin,
                                                            const double* __restrict__ scos, int k, int k2,
                                                            int64_t* __restrict__ out_labels, int32_t* __restrict__ out_ham,
                                                            double* __restrict__ out_sbin, double* __restrict__ out_scos,
                                                            int32_t* __restrict__ out_count, uint32_t* __restrict__ g_src1,
                                                            uint32_t* __restrict__ g_r2) {
    extern __shared__ unsigned long long sm3[];
    const int P = next_pow2(bk);
    unsigned long long* skey = sm3;                   // P
    uint32_t* sval = (uint32_t*)(skey + P);           // P
    // P: phase-I rank -> flat source index (w * bk + j), and P: phase-II rank -> phase-I rank.  Shared memory up to P = 8192;
    // for larger candidate lists the two index arrays live in global scratch (written and read by this block only)
    uint32_t* src1 = g_src1 ? g_src1 + (size_t)blockIdx.x * P : sval + P;
    uint32_t* r2 = g_r2 ? g_r2 + (size_t)blockIdx.x * P : sval + 2 * P;
    __shared__ SelectScratch sc;
    __shared__ int total_s;
    const int q = blockIdx.x, tid = threadIdx.x;
    auto gidx = [&](uint32_t flat) -> size_t { return (size_t)(flat / bk) * (size_t)rank_stride + (size_t)q * bk + (flat % bk); };

    // (A) global phase-I cut: the bk smallest (hamming, position) keys over all ranks
    if (tid == 0) total_s = 0;
    __syncthreads();
    int mine = 0;
    for (int f = tid; f < world * bk; f += M3_THREADS) mine += (keys[gidx(f)] != VRQ_KEY_NONE);
    if (mine) atomicAdd(&total_s, mine);
    for (int i = tid; i < P; i += M3_THREADS) {
        skey[i] = VRQ_KEY_NONE;
        sval[i] = 0xffffffffu;
    }
    __syncthreads();
    const int total = total_s;
    const int c1 = total < bk ? total : bk;
    unsigned long long kth = VRQ_KEY_NONE - 1;
    if (total > bk) {
        kth = radix_select_kth<M3_THREADS>([&](int x) { return (unsigned long long)keys[gidx((uint32_t)x)]; }, world * bk, bk, tid, &sc, 0);
    }
    if (tid == 0) sc.counter = 0;
    __syncthreads();
    for (int f = tid; f < world * bk; f += M3_THREADS) {
        const unsigned long long key = keys[gidx(f)];
        if (key <= kth) {
            const int s = atomicAdd(&sc.counter, 1);
            skey[s] = key;
            sval[s] = (uint32_t)f;
        }
    }
    bitonic_sort<M3_THREADS, true>(skey, sval, P, tid, 0);
    // keep phase-I order; (B) sort by score_binary descending, stable w.r.t. phase-I rank
    for (int i = tid; i < P; i += M3_THREADS) src1[i] = sval[i];
    __syncthreads();
    for (int i = tid; i < P; i += M3_THREADS) {
        if (i < c1) {
            skey[i] = ~ordered_from_double(sbin[gidx(src1[i])]);
            sval[i] = (uint32_t)i;
        } else {
            skey[i] = VRQ_KEY_NONE;
            sval[i] = 0xffffffffu;
        }
    }
    bitonic_sort<M3_THREADS, true>(skey, sval, P, tid, 0);
    const int c2 = c1 < k2 ? c1 : k2;
    for (int i = tid; i < P; i += M3_THREADS) r2[i] = sval[i];
    __syncthreads();
    // (C) sort the first c2 by score_cosine descending, stable w.r.t. phase-II rank
    for (int i = tid; i < P; i += M3_THREADS) {
        if (i < c2) {
            skey[i] = ~ordered_from_double(scos[gidx(src1[r2[i]])]);
            sval[i] = (uint32_t)i;
        } else {
            skey[i] = VRQ_KEY_NONE;
            sval[i] = 0xffffffffu;
        }
    }
    bitonic_sort<M3_THREADS, true>(skey, sval, P, tid, 0);
    const int c3 = c2 < k ? c2 : k;
    for (int i = tid; i < k; i += M3_THREADS) {
        const size_t o = (size_t)q * k + i;
        if (i < c3) {
            const size_t g = gidx(src1[r2[sval[i]]]);
            out_labels[o] = labels[g];
            out_ham[o] = (int32_t)(keys[g] >> VRQ_KEY_POS_BITS);
            out_sbin[o] = sbin[g];
            out_scos[o] = scos[g];
        } else {
            out_labels[o] = -1;
            out_ham[o] = 0x7fffffff;
            out_sbin[o] = -INFINITY;
            out_scos[o] = -INFINITY;
        }
    }
    if (tid == 0) out_count[q] = c3;
}

// ---- 2-phase classes: stable sort of all phase-I hits by float32 score descending, first k --------------------------
__global__ void __launch_bounds__(M3_THREADS) select2_kernel(int m, const uint64_t* __restrict__ keys,
                                                             const int64_t* __restrict__ labels,
                                                             const float* __restrict__ score, int k,
                                                             int64_t* __restrict__ out_labels, float* __restrict__ out_score,
                                                             int32_t* __restrict__ out_count) {
    extern __shared__ unsigned long long sm2[];
    const int P = next_pow2(m);
    unsigned long long* skey = sm2;
    uint32_t* sval = (uint32_t*)(skey + P);
    __shared__ int valid_s;
    const int q = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) valid_s = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < P; i += M3_THREADS) {
        const bool ok = i < m && keys[(size_t)q * m + i] != VRQ_KEY_NONE;
        mine += ok;
        skey[i] = ok ? (0xFFFFFFFFull - ordered_from_float(score[(size_t)q * m + i])) : VRQ_KEY_NONE;
        sval[i] = ok ? (uint32_t)i : 0xffffffffu;
    }
    if (mine) atomicAdd(&valid_s, mine);
    bitonic_sort<M3_THREADS, true>(skey, sval, P, tid, 0);
    const int c = valid_s < k ? valid_s : k;
    for (int i = tid; i < k; i += M3_THREADS) {
        const size_t o = (size_t)q * k + i;
        if (i < c) {
            const size_t g = (size_t)q * m + sval[i];
            out_labels[o] = labels[g];
            out_score[o] = score[g];
        } else {
            out_labels[o] = -1;
            out_score[o] = -INFINITY;
        }
    }
    if (tid == 0) out_count[q] = c;
}

int grid_x_for(vrq_ctx* ctx, int64_t nq, int m) {
    int64_t want = (m + 7) / 8;
    int64_t cap = ((int64_t)ctx->sm_count * 8 + nq - 1) / nq;
    if (cap < 1) cap = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace

int vrq_launch_rescore_binary(vrq_ctx* ctx, const uint8_t* codes, int d, const uint64_t* keys, const int64_t* pos,
                              int64_t pos_base, int64_t nq, int m, const float* qf, double* score, cudaStream_t st) {
    if (nq == 0 || m == 0) return 0;
    if (d % 8 != 0 || d > 12288) {
        vrq_set_error("rescore_binary needs d %% 8 == 0 and d <= 12288 (got %d)", d);
        return VRQ_ERR_UNSUPPORTED;
    }
    vrq_timer_scope ts(ctx, VRQ_CAT_RESCORE, st);
    if (d % 32 != 0) {
        rescore_binary_bytes_kernel<<<dim3(grid_x_for(ctx, nq, m), (unsigned)nq), 256, sizeof(double) * d, st>>>(codes, d, keys, pos, pos_base, m,
                                                                                                                qf, score);
        vrq_count_launch(ctx);
        VRQ_CUDA(cudaGetLastError());
        return 0;
    }
    // d == 1024: the kernels of rescore_mma.cu.  VRQ_RESCORE_BIN = 2 (default): tensor cores (mma.sync s8 over the code bits),
    // 1: nibble table in shared memory, 0: the register kernel below
    const int bin_mode = getenv("VRQ_RESCORE_BIN") ? atoi(getenv("VRQ_RESCORE_BIN")) : VRQ_RESCORE_BIN_DEFAULT;
    if (d == 1024 && bin_mode == 2 && ((uintptr_t)codes % 16) == 0)
        return vrq_launch_rescore_binary_imma(ctx, codes, keys, pos, pos_base, nq, m, qf, score, st);
    if (d == 1024 && bin_mode == 1 && ((uintptr_t)codes % 16) == 0)
        return vrq_launch_rescore_binary_lut(ctx, codes, keys, pos, pos_base, nq, m, qf, score, st);
    dim3 grid(grid_x_for(ctx, nq, m), (unsigned)nq);
    if (d == 1024)
        rescore_binary1024_kernel<<<grid, 256, 0, st>>>(codes, keys, pos, pos_base, m, qf, score);
    else
        rescore_binary_kernel<<<grid, 256, sizeof(double) * d, st>>>(codes, d, keys, pos, pos_base, m, qf, score);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_rescore_int8cos(vrq_ctx* ctx, const int8_t* rows, int d, const uint64_t* keys, const int64_t* pos,
                               int64_t pos_base, int64_t nq, int m, const float* qf, double* score, cudaStream_t st) {
    if (nq == 0 || m == 0) return 0;
    if (d % 8 != 0 || d > 12288) {
        vrq_set_error("rescore_int8cos needs d %% 8 == 0 and d <= 12288 (got %d)", d);
        return VRQ_ERR_UNSUPPORTED;
    }
    vrq_timer_scope ts(ctx, VRQ_CAT_RESCORE, st);
    dim3 grid(grid_x_for(ctx, nq, m), (unsigned)nq);
    const bool use_async = !(getenv("VRQ_RESCORE_ASYNC") && atoi(getenv("VRQ_RESCORE_ASYNC")) == 0);  // read per call (tests switch it)
    const bool use_dp2a = getenv("VRQ_RESCORE_DP2A") && atoi(getenv("VRQ_RESCORE_DP2A")) != 0;  // opt-in: integer dot products
    // tensor-core path (rescore_mma.cu): VRQ_RESCORE_IMMA=1/0; the default is whichever the cfg5 measurement favours
    const bool use_imma = getenv("VRQ_RESCORE_IMMA") ? atoi(getenv("VRQ_RESCORE_IMMA")) != 0 : VRQ_RESCORE_IMMA_DEFAULT;
    if (d == 1024 && use_imma && ((uintptr_t)rows % 16) == 0)
        return vrq_launch_rescore_int8cos_imma(ctx, rows, keys, pos, pos_base, nq, m, qf, score, st);
    if (d == 1024 && use_async && use_dp2a) {
        const size_t smem = (size_t)8 * P3_RING * 1024;
        VRQ_CUDA(cudaFuncSetAttribute(rescore_int8cos_dp2a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rescore_int8cos_dp2a_kernel<<<grid, 256, smem, st>>>(rows, keys, pos, pos_base, m, qf, score);
    } else if (d == 1024 && use_async) {
        const size_t smem = (size_t)8 * P3_RING * 1024;
        if (getenv("VRQ_RESCORE_BATCHRED") && atoi(getenv("VRQ_RESCORE_BATCHRED")) != 0) {  // opt-in, see the kernel's comment
            const size_t smem_b = smem + 8 * 256 * (sizeof(double) + sizeof(int));
            VRQ_CUDA(cudaFuncSetAttribute(rescore_int8cos_async_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
            rescore_int8cos_async_kernel<true><<<grid, 256, smem_b, st>>>(rows, keys, pos, pos_base, m, qf, score);
        } else {
            VRQ_CUDA(cudaFuncSetAttribute(rescore_int8cos_async_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rescore_int8cos_async_kernel<false><<<grid, 256, smem, st>>>(rows, keys, pos, pos_base, m, qf, score);
        }
    } else if (d == 1024)
        rescore_int8cos_kernel<true><<<grid, 256, 0, st>>>(rows, d, keys, pos, pos_base, m, qf, score);
    else
        rescore_int8cos_kernel<false><<<grid, 256, sizeof(double) * d, st>>>(rows, d, keys, pos, pos_base, m, qf, score);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_rescore_payload_dot(vrq_ctx* ctx, const vrq_rescore2_args& a, cudaStream_t st) {
    if (a.nq == 0 || a.m == 0) return 0;
    if (a.d % 8 != 0 || a.d > 12288) {
        vrq_set_error("payload rescoring needs d %% 8 == 0 and d <= 12288 (got %d)", a.d);
        return VRQ_ERR_UNSUPPORTED;
    }
    PayloadParams p{};
    p.kind = a.kind;
    p.payload = a.payload;
    p.aux = a.aux;
    p.d = a.d;
    p.keys = a.keys;
    p.pos_base = a.pos_base;
    p.m = a.m;
    p.qf = a.qf;
    p.score = a.score;
    if (a.kind == VRQ_PAYLOAD_INT8_GLOBAL) p.scale_f32 = (float)(a.limit / 127.0);
    if (a.kind == VRQ_PAYLOAD_INT16_GLOBAL) p.scale_f32 = (float)(a.limit / 32767.0);
    if (a.kind == VRQ_PAYLOAD_INT4_GLOBAL) p.scale_f64 = a.limit / 7.0;
    vrq_timer_scope ts(ctx, VRQ_CAT_RESCORE, st);
    dim3 grid(grid_x_for(ctx, a.nq, a.m), (unsigned)a.nq);
    rescore_payload_kernel<<<grid, 256, sizeof(float) * a.d, st>>>(p);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_keys_to_dist_labels(vrq_ctx* ctx, const uint64_t* keys, int64_t count, int64_t pos_base,
                                   const int64_t* id_map, int64_t id0, int32_t* dist, int64_t* labels, cudaStream_t st) {
    if (count == 0) return 0;
    keys_to_dist_labels_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(keys, count, pos_base, id_map, id0, dist, labels);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_merge3(vrq_ctx* ctx, int world, int64_t nq, int bk, int64_t rank_stride, const uint64_t* keys, const int64_t* labels,
                      const double* sbin, const double* scos, int k, int k2, int64_t* out_labels, int32_t* out_ham,
                      double* out_sbin, double* out_scos, int32_t* out_count, cudaStream_t st) {
    if (nq == 0) return 0;
    if (bk <= 0 || bk > VRQ_MAX_K || k <= 0 || k2 <= 0 || world <= 0) {
        vrq_set_error("merge3: need 1 <= binary_k <= %d, k > 0, k2 > 0, world > 0", VRQ_MAX_K);
        return VRQ_ERR_ARG;
    }
    int P = 1;
    while (P < bk) P <<= 1;
    const bool big = P > 8192;
    size_t smem = (size_t)P * (big ? (8 + 4) : (8 + 4 + 4 + 4));
    void *g1 = nullptr, *g2 = nullptr;
    if (big) {
        VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_M3_SRC, sizeof(uint32_t) * (size_t)nq * P, &g1));
        VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_M3_R2, sizeof(uint32_t) * (size_t)nq * P, &g2));
    }
    if (smem > 40 * 1024)
        VRQ_CUDA(cudaFuncSetAttribute(merge3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 160 * 1024)));
    vrq_timer_scope ts(ctx, VRQ_CAT_MERGE, st);
    if (rank_stride <= 0) rank_stride = nq * (int64_t)bk;
    merge3_kernel<<<(unsigned)nq, M3_THREADS, smem, st>>>(world, (int)nq, bk, rank_stride, keys, labels, sbin, scos, k, k2, out_labels,
                                                          out_ham, out_sbin, out_scos, out_count, (uint32_t*)g1, (uint32_t*)g2);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_select2(vrq_ctx* ctx, int64_t nq, int m, const uint64_t* keys, const int64_t* labels, const float* score,
                       int k, int64_t* out_labels, float* out_score, int32_t* out_count, cudaStream_t st) {
    if (nq == 0) return 0;
    if (m <= 0 || m > VRQ_MAX_K || k <= 0) {
        vrq_set_error("select2: need 1 <= m <= %d and k > 0", VRQ_MAX_K);
        return VRQ_ERR_ARG;
    }
    int P = 1;
    while (P < m) P <<= 1;
    size_t smem = (size_t)P * 12;
    if (smem > 40 * 1024)
        VRQ_CUDA(cudaFuncSetAttribute(select2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 96 * 1024)));
    vrq_timer_scope ts(ctx, VRQ_CAT_MERGE, st);
    select2_kernel<<<(unsigned)nq, M3_THREADS, smem, st>>>(m, keys, labels, score, k, out_labels, out_score, out_count);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}
