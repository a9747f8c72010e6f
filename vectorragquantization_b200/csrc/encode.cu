// Quantise / dequantise / 1-bit-code kernels (sm_100a).
//
// Every encoder is ONE pass over x with a warp per row: min/max/sum by warp shuffles, the quantised row (+ the packbits
// code) written with full-line stores.  d == 1024 runs encode1024_ring_kernel (rows prefetched into a per-warp
// shared-memory ring with cp.async; encode1024_kernel, the register variant, stays as the literal-formula fallback),
// any other d % 8 == 0 the generic kernel.  HBM-bound: 4096 B read + {1024, 2048, 512} (+128) B written per 1024-d row.
//
// Bit-exactness contract: all float ops are single IEEE roundings (__fmul_rn / __fadd_rn /
// __fdiv_rn, no FMA contraction, file compiled with --fmad=false as a second fence); np.mean's float32 pairwise
// summation tree is reproduced add for add (A.2).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "vrq_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

inline int vrq_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

struct EncParams {
    const float* x;
    int64_t n;
    int d;
    float lim;    // np.float32(limit)
    float scale;  // np.float32(qmax/limit)
    void* q;
    void* mn;
    void* mx;
    uint8_t* ubin;
    int ge;
};

// ---- per-element quantisers --------------------------------------------------------------------------
template <int CODEC>
struct RowScale {
    float scale;
    bool constant;
};

// float -> int conversion with x86 cvttss2si semantics (what NumPy's astype compiles to on the reference's hosts):
// NaN and out-of-range values give the "integer indefinite" 0x80000000, whose low byte is 0.  Only reachable when a
// row's max|x| is so small (< ~4e-37) that the per-document scale overflows to inf.
__device__ __forceinline__ int cvt_x86(float t) { return (fabsf(t) < 2147483648.f) ? __float2int_rz(t) : (int)0x80000000; }

__device__ __forceinline__ int q_perdoc8(float v, float scale) {
    return cvt_x86(__fmul_rn(v, scale));  // astype(int8) truncates (VectorDBInt8.py:126)
}
__device__ __forceinline__ int q_global(float v, float lim, float scale, float qmax) {
    float c = fminf(fmaxf(v, -lim), lim);       // np.clip(x, -limit, limit)
    float s = rintf(__fmul_rn(c, scale));       // np.round: half to even
    s = fminf(fmaxf(s, -qmax), qmax);           // np.clip(., -qmax, qmax)
    return __float2int_rz(s);
}
__device__ __forceinline__ int q_int4(float v, float scale) {
    float s = rintf(__fmul_rn(v, scale));
    s = (s != s) ? s : fminf(fmaxf(s, -8.f), 7.f);  // np.clip propagates NaN (0 * inf when the scale overflowed)
    return (cvt_x86(s) + 8) & 0xF;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// =====================================================================================================
// Fast path, d == 1024.  Lane l holds elements 128*j + 4*l .. +3 of the row for j = 0..7 (8 x LDG.128).
// =====================================================================================================
constexpr int ROW_PAD = 136;  // floats per 128-element block in shared memory (8 floats of padding)

template <int CODEC, bool UBIN>
__global__ void __launch_bounds__(256) encode1024_kernel(EncParams p) {
    __shared__ __align__(16) float sbuf[UBIN ? 8 : 1][UBIN ? 8 * ROW_PAD : 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < p.n; row += stride) {
        const float4* src = reinterpret_cast<const float4*>(p.x + row * 1024);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = ldg_stream_f4(src + j * 32 + lane);

        // ---- np.mean(x) with NumPy's pairwise tree ----------------------------------
        // 8 blocks of 128; inside a block 8 strided accumulators r[jj] += x[128b + 8s + jj], s = 0..15 in order;
        // block = ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); total = ((B0+B1)+(B2+B3))+((B4+B5)+(B6+B7)).
        // The row is transposed through padded shared memory so that lane (b = l/4, jp = l%4) owns the two
        // sequential chains jj = 2jp, 2jp+1 of block b (conflict-free LDS.64), then combined with xor-shuffles.
        float mean = 0.f;
        if (UBIN) {
            float* sb = sbuf[warp];
#pragma unroll
            for (int j = 0; j < 8; j++) *reinterpret_cast<float4*>(sb + ROW_PAD * j + 4 * lane) = v[j];
            __syncwarp();
            const int b = lane >> 2, jp = lane & 3;
            const float2* cp = reinterpret_cast<const float2*>(sb + ROW_PAD * b + 2 * jp);
            float2 e = cp[0];
            float r0 = e.x, r1 = e.y;
#pragma unroll
            for (int s = 1; s < 16; s++) {
                e = cp[4 * s];
                r0 = __fadd_rn(r0, e.x);
                r1 = __fadd_rn(r1, e.y);
            }
            float t = __fadd_rn(r0, r1);
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 1));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 2));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 4));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 8));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 16));
            mean = __fdiv_rn(t, 1024.f);
            __syncwarp();
        }

        // ---- per-row statistics ---------------------------------------------------------------------
        float scale = p.scale;
        bool constant = false;
        if (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT4) {
            float lo = v[0].x, hi = v[0].x;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                lo = fminf(fminf(fminf(lo, v[j].x), fminf(v[j].y, v[j].z)), v[j].w);
                hi = fmaxf(fmaxf(fmaxf(hi, v[j].x), fmaxf(v[j].y, v[j].z)), v[j].w);
            }
            lo = warp_min(lo);
            hi = warp_max(hi);
            constant = (lo == hi);
            const float m = fmaxf(fabsf(lo), fabsf(hi));
            if (CODEC == VRQ_CODEC_INT8_PERDOC) {
                scale = __fdiv_rn(127.f, m);  // np.float32 division (VectorDBInt8.py:125)
                if (lane == 0) {
                    if (p.mn) static_cast<float*>(p.mn)[row] = lo;
                    if (p.mx) static_cast<float*>(p.mx)[row] = hi;
                }
            } else {
                scale = (float)(7.0 / (double)m);  // Python float division, then cast (VectorDBInt4.py:136)
                if (lane == 0) {
                    if (p.mn) static_cast<double*>(p.mn)[row] = (double)lo;
                    if (p.mx) static_cast<double*>(p.mx)[row] = (double)hi;
                }
            }
        }

        // ---- quantise + store -----------------------------------------------------------------------
        if (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT8_GLOBAL) {
            uint32_t* dst = static_cast<uint32_t*>(p.q) + row * 256;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int a, b, c, d;
                if (CODEC == VRQ_CODEC_INT8_PERDOC) {
                    a = q_perdoc8(v[j].x, scale), b = q_perdoc8(v[j].y, scale);
                    c = q_perdoc8(v[j].z, scale), d = q_perdoc8(v[j].w, scale);
                } else {
                    a = q_global(v[j].x, p.lim, scale, 127.f), b = q_global(v[j].y, p.lim, scale, 127.f);
                    c = q_global(v[j].z, p.lim, scale, 127.f), d = q_global(v[j].w, p.lim, scale, 127.f);
                }
                uint32_t w = (a & 0xFF) | ((b & 0xFF) << 8) | ((c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24);
                dst[j * 32 + lane] = constant ? 0u : w;
            }
        } else if (CODEC == VRQ_CODEC_INT16_GLOBAL) {
            uint2* dst = static_cast<uint2*>(p.q) + row * 256;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int a = q_global(v[j].x, p.lim, scale, 32767.f), b = q_global(v[j].y, p.lim, scale, 32767.f);
                int c = q_global(v[j].z, p.lim, scale, 32767.f), d = q_global(v[j].w, p.lim, scale, 32767.f);
                uint2 w;
                w.x = (a & 0xFFFF) | ((uint32_t)(b & 0xFFFF) << 16);
                w.y = (c & 0xFFFF) | ((uint32_t)(d & 0xFFFF) << 16);
                dst[j * 32 + lane] = w;
            }
        } else if (CODEC == VRQ_CODEC_INT4) {
            uint16_t* dst = static_cast<uint16_t*>(p.q) + row * 256;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t b0 = (q_int4(v[j].x, scale) << 4) | q_int4(v[j].y, scale);
                uint32_t b1 = (q_int4(v[j].z, scale) << 4) | q_int4(v[j].w, scale);
                dst[j * 32 + lane] = constant ? (uint16_t)0 : (uint16_t)(b0 | (b1 << 8));
            }
        }

        // ---- np.packbits(x > mean): MSB-first; lane l owns nibble (l odd: low, l even: high) of byte 16j + l/2 ----
        if (UBIN) {
            uint32_t nibs = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t nb;
                if (p.ge)
                    nb = ((v[j].x >= mean) << 3) | ((v[j].y >= mean) << 2) | ((v[j].z >= mean) << 1) | (v[j].w >= mean);
                else
                    nb = ((v[j].x > mean) << 3) | ((v[j].y > mean) << 2) | ((v[j].z > mean) << 1) | (v[j].w > mean);
                nibs |= nb << (4 * j);
            }
            // lane (8m + j') assembles 32-bit word 4j' + m of the 128-byte code from the 8 lanes of its group
            const int jsel = lane & 7, grp = lane & 24;
            uint32_t word = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t ns = __shfl_sync(FULL, nibs, grp + s);
                uint32_t nb = (ns >> (4 * jsel)) & 0xF;
                word |= nb << (8 * (s >> 1) + ((s & 1) ? 0 : 4));
            }
            reinterpret_cast<uint32_t*>(p.ubin + row * 128)[4 * jsel + (lane >> 3)] = word;
        }
    }
}

// =====================================================================================================
// Fast path with a prefetch ring, d == 1024 (the default).  The register kernel above keeps at most one row per warp
// in flight and stalls on it (ncu r01: 40 % warps active, long-scoreboard bound, 5.1-5.8 TB/s).  Here every warp owns a
// ring of RING_STAGES rows in shared memory filled by cp.async (LDGSTS), so RING_STAGES-1 .. RING_STAGES rows per warp
// are always on their way from HBM while one is encoded.  A row sits in the ring as 256 16-byte chunks, chunk c stored
// at c ^ ring_swz(c >> 3): with that XOR all three access patterns are bank-conflict free -
//   fill   : lane l copies chunks 32j + l (coalesced 512 B global reads),
//   mean   : lane (b, jp) walks its two strided accumulator chains of 128-block b with LDS.64 (np.mean's tree, A.2),
//   encode : lane l reads its 32 CONSECUTIVE elements 32l .. 32l+31 with 8 LDS.128.
// Owning consecutive elements is what makes the rest cheap: the packbits word of the lane is one 32-bit store and
// the quantised elements go out as 16-byte stores, with no shuffles.  Rounding uses the 1.5 * 2^23 magic-number add
// (round-half-even of the FADD itself, the integer lands in the low mantissa bits) instead of FRND + F2I, which run at
// quarter rate; clipping to the integer bounds commutes with rint, so it is done first and the result is the same for
// every finite input.  Rows whose scale is not an ordinary float (max|x| outside [1e-36, 1e36], NaN) take the
// literal formulas of the register kernel.
// =====================================================================================================
constexpr int RING_MAX_WARPS = 8;  // warps per block (run-time: blockDim.x / 32)
constexpr int RING_MAX_STAGES = 6;

__device__ __forceinline__ int ring_swz(int line) { return (line & 7) ^ (((line >> 3) & 1) << 1); }

template <int STAGES>
__device__ __forceinline__ void ring_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
}

__device__ __forceinline__ uint32_t pack4_low_bytes(int a, int b, int c, int d) {
    const uint32_t lo = __byte_perm((uint32_t)a, (uint32_t)b, 0x0040);
    const uint32_t hi = __byte_perm((uint32_t)c, (uint32_t)d, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

// w |= BIT if a > b (a >= b): FSETP + predicated LOP3, exact IEEE comparison (false for NaN, -0 == +0)
template <uint32_t BIT>
__device__ __forceinline__ void or_if_gt(uint32_t& w, float a, float b) {
    asm("{\n\t.reg .pred q;\n\tsetp.gt.f32 q, %1, %2;\n\t@q or.b32 %0, %0, %3;\n\t}" : "+r"(w) : "f"(a), "f"(b), "n"(BIT));
}
template <uint32_t BIT>
__device__ __forceinline__ void or_if_ge(uint32_t& w, float a, float b) {
    asm("{\n\t.reg .pred q;\n\tsetp.ge.f32 q, %1, %2;\n\t@q or.b32 %0, %0, %3;\n\t}" : "+r"(w) : "f"(a), "f"(b), "n"(BIT));
}
// the four bits of chunk U of a run (elements 4U .. 4U+3): bit 7 - (4(U&1) + c) of byte U/2
template <int U, bool GE>
__device__ __forceinline__ uint32_t chunk_bits(const float4& f, float mean) {
    constexpr int BASE = 8 * (U >> 1) + 7 - 4 * (U & 1);
    uint32_t w = 0;
    if (GE) {
        or_if_ge<1u << BASE>(w, f.x, mean);
        or_if_ge<1u << (BASE - 1)>(w, f.y, mean);
        or_if_ge<1u << (BASE - 2)>(w, f.z, mean);
        or_if_ge<1u << (BASE - 3)>(w, f.w, mean);
    } else {
        or_if_gt<1u << BASE>(w, f.x, mean);
        or_if_gt<1u << (BASE - 1)>(w, f.y, mean);
        or_if_gt<1u << (BASE - 2)>(w, f.z, mean);
        or_if_gt<1u << (BASE - 3)>(w, f.w, mean);
    }
    return w;
}
template <int CPR, bool GE>
__device__ __forceinline__ uint32_t run_bits(const float4* v, float mean) {
    uint32_t w = chunk_bits<0, GE>(v[0], mean);
    if (CPR > 1) w |= chunk_bits<1, GE>(v[1 % CPR], mean);
    if (CPR > 2) w |= chunk_bits<2, GE>(v[2 % CPR], mean) | chunk_bits<3, GE>(v[3 % CPR], mean);
    if (CPR > 4)
        w |= (chunk_bits<4, GE>(v[4 % CPR], mean) | chunk_bits<5, GE>(v[5 % CPR], mean)) |
             (chunk_bits<6, GE>(v[6 % CPR], mean) | chunk_bits<7, GE>(v[7 % CPR], mean));
    return w;
}

// np.clip(x, -lim, lim) for lim > 0 in one instruction: min(|x|, lim) carrying the sign of x
__device__ __forceinline__ float clip_sym(float x, float lim) {
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lim));
    return r;
}

constexpr float MAGIC_RINT = 12582912.f;  // 1.5 * 2^23: (t + MAGIC) has ulp 1, so the add rounds t half-to-even

template <int CODEC, bool UBIN, int STAGES>
__global__ void __launch_bounds__(RING_MAX_WARPS * 32, 3) encode1024_ring_kernel(EncParams p) {
    extern __shared__ __align__(1024) unsigned char ring_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 32-bit shared address of this warp's ring; 1024-byte aligned, so the XOR parts below (< 512) never carry into it
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring_raw) + (uint32_t)warp * STAGES * 4096;
    const int nwarps = blockDim.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    const int64_t row0 = (int64_t)blockIdx.x * nwarps + warp;

    // Every swizzled byte offset splits into (lane part) ^ (compile-time part) + (compile-time add); only the lane
    // parts live in registers (checked against ring_swz for all lanes in profiles/debug/check_ring_swizzle.py):
    //   fill, chunk 32j + lane          : 512j + (fill_l ^ 16 * (4(j&1) ^ 2((j>>1)&1)))
    //   mean, lane (b, jp), step s      : (mean_l ^ 16 * (2(s&3) ^ (s>>2))) + 128 (s>>2)
    //   encode, run k, chunk u          : (4096/K) k + (enc_l ^ 16u)          [K = 4: 1024k + (enc_l ^ 16 (u ^ 2(k&1)))]
    constexpr int K = (CODEC == VRQ_CODEC_INT16_GLOBAL) ? 4 : (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT8_GLOBAL) ? 2 : 1;
    constexpr int CPR = 8 / K;  // chunks per run
    const uint32_t fill_l = (uint32_t)(lane ^ (lane >> 3)) * 16;
    const int mb = lane >> 2, mjp = lane & 3;
    const uint32_t mean_l = 512u * mb + (uint32_t)((mjp >> 1) ^ (4 * (mb & 1)) ^ (2 * ((mb >> 1) & 1))) * 16 + (mjp & 1) * 8;
    const uint32_t enc_l = K == 1   ? (uint32_t)((8 * lane) ^ ((lane & 7) ^ (((lane >> 3) & 1) << 1))) * 16
                           : K == 2 ? (uint32_t)((4 * lane) ^ (((lane >> 1) & 7) ^ (((lane >> 4) & 1) << 1))) * 16
                                    : (uint32_t)((2 * lane) ^ ((lane >> 2) & 7)) * 16;

    auto fill = [&](int64_t row, int stage) {
        if (row < p.n) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.x + row * 1024) + lane * 16;
            const uint32_t dst = ring_s + stage * 4096;
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512 * j + (fill_l ^ (16u * ((4 * (j & 1)) ^ (2 * ((j >> 1) & 1)))))),
                             "l"(src + 512 * j)
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

#pragma unroll
    for (int s = 0; s < STAGES; s++) fill(row0 + s * stride, s);

    int stage = 0;
    for (int64_t row = row0; row < p.n; row += stride) {
        ring_wait<STAGES>();
        __syncwarp();
        const uint32_t sb = ring_s + stage * 4096;

        // ---- np.mean(x): same tree as encode1024_kernel, read through the swizzle ------------------------------
        float mean = 0.f;
        if (UBIN) {
            const uint32_t a1 = sb + mean_l;
            float r0 = 0.f, r1 = 0.f;
#pragma unroll
            for (int s = 0; s < 16; s++) {
                float2 e;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                             : "=f"(e.x), "=f"(e.y)
                             : "r"((a1 ^ (16u * ((2 * (s & 3)) ^ (s >> 2)))) + 128 * (s >> 2)));
                r0 = s ? __fadd_rn(r0, e.x) : e.x;
                r1 = s ? __fadd_rn(r1, e.y) : e.y;
            }
            float t = __fadd_rn(r0, r1);
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 1));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 2));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 4));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 8));
            t = __fadd_rn(t, __shfl_xor_sync(FULL, t, 16));
            mean = __fmul_rn(t, 0.0009765625f);  // == t / 1024 for every t (a power-of-two scaling rounds the same way)
        }

        // ---- the lane's elements: K runs of 32/K CONSECUTIVE elements, run k = elements 1024k/K + (32/K) * lane .. -------
        // K = 16-byte stores the quantised row takes per lane (int4: 1, int8: 2, int16: 4), so every store instruction
        // of the warp writes 512 contiguous bytes.  v[k * 8/K + u] = chunk k * 256/K + lane * 8/K + u (conflict-free too).
        float4 v[8];
        {
            const uint32_t a2 = sb + enc_l;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = i / CPR, u = i % CPR;
                const uint32_t addr = (K == 4) ? (a2 ^ (16u * (u ^ (2 * (k & 1))))) + 1024 * k : (a2 ^ (16u * u)) + (4096 / K) * k;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(addr));
            }
        }

        // np.packbits(x > mean), MSB first: FSETP + predicated OR per element (run_bits)
        if (UBIN) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t word = p.ge ? run_bits<CPR, true>(v + k * CPR, mean) : run_bits<CPR, false>(v + k * CPR, mean);
                uint8_t* code = p.ubin + row * 128;
                if (K == 1)
                    reinterpret_cast<uint32_t*>(code)[lane] = word;
                else if (K == 2)
                    reinterpret_cast<uint16_t*>(code)[32 * k + lane] = (uint16_t)word;
                else
                    code[32 * k + lane] = (uint8_t)word;
            }
        }

        // ---- per-row statistics (they consume every loaded register: the stage can be refilled after them) -----
        float scale = p.scale;
        bool constant = false, ordinary = true;
        if (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT4) {
            float lo = v[0].x, hi = v[0].x;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                lo = fminf(fminf(fminf(lo, v[u].x), fminf(v[u].y, v[u].z)), v[u].w);
                hi = fmaxf(fmaxf(fmaxf(hi, v[u].x), fmaxf(v[u].y, v[u].z)), v[u].w);
            }
            lo = warp_min(lo);
            hi = warp_max(hi);
            constant = (lo == hi);
            const float m = fmaxf(fabsf(lo), fabsf(hi));
            ordinary = (m >= 1e-36f) && (m <= 1e36f);
            if (CODEC == VRQ_CODEC_INT8_PERDOC) {
                scale = __fdiv_rn(127.f, m);
                if (lane == 0) {
                    if (p.mn) static_cast<float*>(p.mn)[row] = lo;
                    if (p.mx) static_cast<float*>(p.mx)[row] = hi;
                }
            } else {
                scale = (float)(7.0 / (double)m);
                if (lane == 0) {
                    if (p.mn) static_cast<double*>(p.mn)[row] = (double)lo;
                    if (p.mx) static_cast<double*>(p.mx)[row] = (double)hi;
                }
            }
        }

        // every lane's loads from this stage have been consumed by the arithmetic above (packbits compares / min-max):
        // hand the stage back now; without either the refill waits until the quantiser below has used the registers
        constexpr bool EARLY = UBIN || CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT4;
        if (EARLY) {
            __syncwarp();
            fill(row + (int64_t)STAGES * stride, stage);
        }

        // ---- quantise + store ------------------------------------------------------------------------------------
        if (CODEC == VRQ_CODEC_INT8_PERDOC) {
            uint32_t w[8];
            if (ordinary) {
#pragma unroll
                for (int u = 0; u < 8; u++)
                    w[u] = pack4_low_bytes(__float2int_rz(__fmul_rn(v[u].x, scale)), __float2int_rz(__fmul_rn(v[u].y, scale)),
                                           __float2int_rz(__fmul_rn(v[u].z, scale)), __float2int_rz(__fmul_rn(v[u].w, scale)));
            } else {
#pragma unroll
                for (int u = 0; u < 8; u++)
                    w[u] = pack4_low_bytes(q_perdoc8(v[u].x, scale), q_perdoc8(v[u].y, scale), q_perdoc8(v[u].z, scale),
                                           q_perdoc8(v[u].w, scale));
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<int8_t*>(p.q) + row * 1024) + lane;
            dst[0] = constant ? make_uint4(0, 0, 0, 0) : make_uint4(w[0], w[1], w[2], w[3]);
            dst[32] = constant ? make_uint4(0, 0, 0, 0) : make_uint4(w[4], w[5], w[6], w[7]);
        } else if (CODEC == VRQ_CODEC_INT8_GLOBAL) {
            uint32_t w[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                int b[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float cl = clip_sym(e[c], p.lim);
                    b[c] = __float_as_int(__fadd_rn(__fmul_rn(cl, scale), MAGIC_RINT));
                }
                w[u] = pack4_low_bytes(b[0], b[1], b[2], b[3]);
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<int8_t*>(p.q) + row * 1024) + lane;
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[32] = make_uint4(w[4], w[5], w[6], w[7]);
        } else if (CODEC == VRQ_CODEC_INT16_GLOBAL) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<int16_t*>(p.q) + row * 1024) + lane;
#pragma unroll
            for (int h = 0; h < 4; h++) {
                uint32_t w[4];
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const float4 f = v[2 * h + t];
                    const float e[4] = {f.x, f.y, f.z, f.w};
                    int b[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float cl = clip_sym(e[c], p.lim);
                        b[c] = __float_as_int(__fadd_rn(__fmul_rn(cl, scale), MAGIC_RINT));
                    }
                    w[2 * t] = __byte_perm((uint32_t)b[0], (uint32_t)b[1], 0x5410);
                    w[2 * t + 1] = __byte_perm((uint32_t)b[2], (uint32_t)b[3], 0x5410);
                }
                dst[32 * h] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else if (CODEC == VRQ_CODEC_INT4) {
            uint32_t w[4];
            if (ordinary) {
                // (c + MAGIC + 8) carries rint(c) + 8 = the nibble in its low 4 bits; two of them make a byte with one IMAD
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    int by[4];
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const float4 f = v[2 * h + t];
                        const float e[4] = {f.x, f.y, f.z, f.w};
                        int b[4];
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            // |x * scale| <= 7 (1 + 2^-23) on an ordinary row, so np.clip(., -8, 7) never acts
                            b[c] = __float_as_int(__fadd_rn(__fmul_rn(e[c], scale), MAGIC_RINT + 8.f));
                        }
                        by[2 * t] = b[0] * 16 + b[1];
                        by[2 * t + 1] = b[2] * 16 + b[3];
                    }
                    w[h] = pack4_low_bytes(by[0], by[1], by[2], by[3]);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    int by[4];
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const float4 f = v[2 * h + t];
                        by[2 * t] = (q_int4(f.x, scale) << 4) | q_int4(f.y, scale);
                        by[2 * t + 1] = (q_int4(f.z, scale) << 4) | q_int4(f.w, scale);
                    }
                    w[h] = pack4_low_bytes(by[0], by[1], by[2], by[3]);
                }
            }
            uint4* dst = reinterpret_cast<uint4*>(static_cast<int8_t*>(p.q) + row * 512) + lane;
            dst[0] = constant ? make_uint4(0, 0, 0, 0) : make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (!EARLY) {
            __syncwarp();
            fill(row + (int64_t)STAGES * stride, stage);
        }
        stage = (stage + 1 == STAGES) ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// =====================================================================================================
// Generic path: any d with d % 8 == 0, 8 <= d <= 8192.  One warp per row, the row staged in shared memory,
// np.mean's recursion (split at n/2 rounded down to a multiple of 8 until n <= 128) evaluated leaf by leaf.
// =====================================================================================================
constexpr int GEN_WARPS = 4;
constexpr int GEN_MAX_D = 8192;
constexpr int GEN_MAX_LEAVES = GEN_MAX_D / 64;

__device__ void enum_leaves(int off, int n, int* loff, int* llen, int& cnt) {
    if (n <= 128) {
        loff[cnt] = off;
        llen[cnt] = n;
        cnt++;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    enum_leaves(off, n2, loff, llen, cnt);
    enum_leaves(off + n2, n - n2, loff, llen, cnt);
}
__device__ float combine_leaves(const float* leaf, int& idx, int n) {
    if (n <= 128) return leaf[idx++];
    int n2 = n / 2;
    n2 -= n2 % 8;
    float a = combine_leaves(leaf, idx, n2);
    float b = combine_leaves(leaf, idx, n - n2);
    return __fadd_rn(a, b);
}

template <int CODEC, bool UBIN>
__global__ void __launch_bounds__(GEN_WARPS * 32) encode_generic_kernel(EncParams p) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ int loff[GEN_MAX_LEAVES], llen[GEN_MAX_LEAVES];
    __shared__ int nleaves;
    __shared__ float leafsum[GEN_WARPS][GEN_MAX_LEAVES];
    const int d = p.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* row_s = dyn + (size_t)warp * d;
    if (threadIdx.x == 0) {
        int c = 0;
        enum_leaves(0, d, loff, llen, c);
        nleaves = c;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * GEN_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * GEN_WARPS + warp; row < p.n; row += stride) {
        const float4* src = reinterpret_cast<const float4*>(p.x + row * d);
        for (int c = lane; c < d / 4; c += 32) reinterpret_cast<float4*>(row_s)[c] = ldg_stream_f4(src + c);
        __syncwarp();
        float mean = 0.f;
        if (UBIN) {
            const int g = lane >> 3, j8 = lane & 7;
            for (int l0 = 0; l0 < nleaves; l0 += 4) {
                const int l = l0 + g;
                const bool act = l < nleaves;
                const int off = act ? loff[l] : 0, len = act ? llen[l] : 8;
                float r = row_s[off + j8];
                for (int i = 8; i < len; i += 8) r = __fadd_rn(r, row_s[off + i + j8]);
                r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 1));
                r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 2));
                r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 4));
                if (act && j8 == 0) leafsum[warp][l] = r;
            }
            __syncwarp();
            float s = 0.f;
            if (lane == 0) {
                int idx = 0;
                s = combine_leaves(leafsum[warp], idx, d);
            }
            s = __shfl_sync(FULL, s, 0);
            mean = __fdiv_rn(s, (float)d);
        }
        float scale = p.scale;
        bool constant = false;
        if (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT4) {
            float lo = row_s[0], hi = row_s[0];
            for (int c = lane; c < d; c += 32) {
                lo = fminf(lo, row_s[c]);
                hi = fmaxf(hi, row_s[c]);
            }
            lo = warp_min(lo);
            hi = warp_max(hi);
            constant = (lo == hi);
            const float m = fmaxf(fabsf(lo), fabsf(hi));
            if (CODEC == VRQ_CODEC_INT8_PERDOC) {
                scale = __fdiv_rn(127.f, m);
                if (lane == 0) {
                    if (p.mn) static_cast<float*>(p.mn)[row] = lo;
                    if (p.mx) static_cast<float*>(p.mx)[row] = hi;
                }
            } else {
                scale = (float)(7.0 / (double)m);
                if (lane == 0) {
                    if (p.mn) static_cast<double*>(p.mn)[row] = (double)lo;
                    if (p.mx) static_cast<double*>(p.mx)[row] = (double)hi;
                }
            }
        }
        if (CODEC == VRQ_CODEC_INT8_PERDOC || CODEC == VRQ_CODEC_INT8_GLOBAL) {
            int8_t* dst = static_cast<int8_t*>(p.q) + row * d;
            for (int c = lane; c < d; c += 32) {
                int a = (CODEC == VRQ_CODEC_INT8_PERDOC) ? q_perdoc8(row_s[c], scale)
                                                         : q_global(row_s[c], p.lim, scale, 127.f);
                dst[c] = constant ? (int8_t)0 : (int8_t)a;
            }
        } else if (CODEC == VRQ_CODEC_INT16_GLOBAL) {
            int16_t* dst = static_cast<int16_t*>(p.q) + row * d;
            for (int c = lane; c < d; c += 32) dst[c] = (int16_t)q_global(row_s[c], p.lim, scale, 32767.f);
        } else if (CODEC == VRQ_CODEC_INT4) {
            uint8_t* dst = static_cast<uint8_t*>(p.q) + row * (d / 2);
            for (int c = lane; c < d / 2; c += 32) {
                uint32_t b = (q_int4(row_s[2 * c], scale) << 4) | q_int4(row_s[2 * c + 1], scale);
                dst[c] = constant ? (uint8_t)0 : (uint8_t)b;
            }
        }
        if (UBIN) {
            uint8_t* dst = p.ubin + row * (d / 8);
            for (int b = lane; b < d / 8; b += 32) {
                uint32_t v = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float e = row_s[8 * b + j];
                    v = (v << 1) | (uint32_t)(p.ge ? (e >= mean) : (e > mean));
                }
                dst[b] = (uint8_t)v;
            }
        }
        __syncwarp();
    }
}

// ---- _to_binary on int8 / int16 rows: exact integer form d*x > sum(x) -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256) to_binary_int_kernel(const T* __restrict__ x, int64_t n, int d, int ge,
                                                            uint8_t* __restrict__ ubin) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += stride) {
        const T* src = x + row * d;
        int s = 0;
        for (int c = lane; c < d; c += 32) s += (int)src[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        uint8_t* dst = ubin + row * (d / 8);
        for (int b = lane; b < d / 8; b += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int e = (int)src[8 * b + j] * d;
                v = (v << 1) | (uint32_t)(ge ? (e >= s) : (e > s));
            }
            dst[b] = (uint8_t)v;
        }
    }
}

// ---- dequantisers (elementwise) ------------------------------------------------------------------------
struct DeqParams {
    int kind;
    const void* q;
    int64_t n;
    int d;
    const void* mn;
    const void* mx;
    float scale_f32;   // global kinds: np.float32(limit / qmax)
    double scale_f64;  // INT4_GLOBAL: limit / 7.0
    float* out;
};

__global__ void __launch_bounds__(256) dequant_kernel(DeqParams p) {
    const int64_t total = p.n * (int64_t)p.d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / p.d;
        const int col = (int)(e - row * p.d);
        float r;
        switch (p.kind) {
            case VRQ_PAYLOAD_INT8_PERDOC: {
                float lo = static_cast<const float*>(p.mn)[row], hi = static_cast<const float*>(p.mx)[row];
                float sc = __fdiv_rn(fmaxf(fabsf(lo), fabsf(hi)), 127.f);
                r = (lo == hi) ? 0.f : __fmul_rn((float)static_cast<const int8_t*>(p.q)[e], sc);
                break;
            }
            case VRQ_PAYLOAD_INT8_GLOBAL:
                r = __fmul_rn((float)static_cast<const int8_t*>(p.q)[e], p.scale_f32);
                break;
            case VRQ_PAYLOAD_INT16_GLOBAL:
                r = __fmul_rn((float)static_cast<const int16_t*>(p.q)[e], p.scale_f32);
                break;
            case VRQ_PAYLOAD_INT4_PERDOC:
            case VRQ_PAYLOAD_INT4_GLOBAL: {
                uint8_t byte = static_cast<const uint8_t*>(p.q)[row * (p.d / 2) + (col >> 1)];
                int nib = (col & 1) ? (byte & 0xF) : (byte >> 4);
                double sc = p.scale_f64;
                bool zero = false;
                if (p.kind == VRQ_PAYLOAD_INT4_PERDOC) {
                    double lo = static_cast<const double*>(p.mn)[row], hi = static_cast<const double*>(p.mx)[row];
                    sc = fmax(fabs(lo), fabs(hi)) / 7.0;
                    zero = (lo == hi);
                }
                r = zero ? 0.f : (float)__dmul_rn((double)(nib - 8), sc);
                break;
            }
            default:
                r = 0.f;
        }
        p.out[e] = r;
    }
}

// The magic-number rounding of the ring kernel needs |clip(x) * scale| to stay below qmax + 0.5 (it always does for a
// positive, ordinary limit: the product is qmax * (1 +- 2^-23)); anything else goes to the register kernel.
template <int CODEC>
bool ring_params_ok(const EncParams& p) {
    if (CODEC == VRQ_CODEC_INT8_GLOBAL || CODEC == VRQ_CODEC_INT16_GLOBAL) {
        const float qmax = CODEC == VRQ_CODEC_INT8_GLOBAL ? 127.f : 32767.f;
        const float top = p.lim * p.scale;
        return p.lim > 0.f && isfinite(p.scale) && p.scale > 0.f && top < qmax + 0.49f;
    }
    return true;
}

template <int CODEC, bool UBIN, int STAGES>
int launch_ring(const EncParams& p, int grid, int warps, size_t smem, cudaStream_t st) {
    auto kern = encode1024_ring_kernel<CODEC, UBIN, STAGES>;
    VRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, warps * 32, smem, st>>>(p);
    return 0;
}

template <int CODEC>
int launch_codec(vrq_ctx* ctx, const EncParams& p, cudaStream_t st) {
    const bool ub = p.ubin != nullptr;
    if (p.d == 1024 && ring_params_ok<CODEC>(p) && vrq_env_int("VRQ_ENCODE_RING", 1) != 0) {
        // launch shape (B200 sweeps, profiles/r01/encode_ring_sweep*.txt): one block of 8 warps with a 4-stage ring per SM
        // (3-4 rows per warp in flight) is best for every codec except per-document int8, whose min/max shuffle chain +
        // float divide + F2I want 16 warps per SM (2 blocks of 8 warps x 3 stages)
        int stages = vrq_env_int("VRQ_ENCODE_STAGES", CODEC == VRQ_CODEC_INT8_PERDOC ? 3 : 4);
        stages = stages < 2 ? 2 : (stages > RING_MAX_STAGES ? RING_MAX_STAGES : stages);
        if (stages == 5) stages = 4;
        int warps = vrq_env_int("VRQ_ENCODE_WARPS", 8);
        warps = warps < 1 ? 1 : (warps > RING_MAX_WARPS ? RING_MAX_WARPS : warps);
        const size_t smem = (size_t)warps * stages * 4096;
        int per_sm = (int)((size_t)(ctx->smem_optin ? ctx->smem_optin : 227 * 1024) / (smem + 1024));
        const int by_threads = 2048 / (warps * 32), by_regs = 65536 / (80 * warps * 32);
        per_sm = per_sm < by_threads ? per_sm : by_threads;
        per_sm = per_sm < by_regs ? per_sm : by_regs;
        int64_t blocks = (p.n + warps - 1) / warps;
        int64_t cap = (int64_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
        int grid = (int)(blocks < cap ? blocks : cap);
        int rc;
        if (stages == 2)
            rc = ub ? launch_ring<CODEC, true, 2>(p, grid, warps, smem, st) : launch_ring<CODEC, false, 2>(p, grid, warps, smem, st);
        else if (stages == 3)
            rc = ub ? launch_ring<CODEC, true, 3>(p, grid, warps, smem, st) : launch_ring<CODEC, false, 3>(p, grid, warps, smem, st);
        else if (stages == 4)
            rc = ub ? launch_ring<CODEC, true, 4>(p, grid, warps, smem, st) : launch_ring<CODEC, false, 4>(p, grid, warps, smem, st);
        else
            rc = ub ? launch_ring<CODEC, true, 6>(p, grid, warps, smem, st) : launch_ring<CODEC, false, 6>(p, grid, warps, smem, st);
        if (rc) return rc;
    } else if (p.d == 1024) {
        int64_t blocks = (p.n + 7) / 8;
        int64_t cap = (int64_t)ctx->sm_count * 6;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (ub)
            encode1024_kernel<CODEC, true><<<grid, 256, 0, st>>>(p);
        else
            encode1024_kernel<CODEC, false><<<grid, 256, 0, st>>>(p);
    } else {
        size_t smem = (size_t)GEN_WARPS * p.d * sizeof(float);
        auto kt = encode_generic_kernel<CODEC, true>;
        auto kf = encode_generic_kernel<CODEC, false>;
        if (smem > 40 * 1024) {
            VRQ_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            VRQ_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        int64_t blocks = (p.n + GEN_WARPS - 1) / GEN_WARPS;
        int64_t cap = (int64_t)ctx->sm_count * 8;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (ub)
            kt<<<grid, GEN_WARPS * 32, smem, st>>>(p);
        else
            kf<<<grid, GEN_WARPS * 32, smem, st>>>(p);
    }
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int vrq_launch_encode(vrq_ctx* ctx, const vrq_encode_args& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.d % 8 != 0 || a.d < 8) {
        vrq_set_error("embedding_dim must be a positive multiple of 8 (got %d)", a.d);
        return VRQ_ERR_ARG;
    }
    if (a.d != 1024 && a.d > GEN_MAX_D) {
        vrq_set_error("embedding_dim %d > %d is not supported by the encode kernels", a.d, GEN_MAX_D);
        return VRQ_ERR_UNSUPPORTED;
    }
    // every kernel moves rows with 16-byte vector accesses (torch allocations and the library's staging buffers are
    // 256-byte aligned; a caller-made sub-view might not be): refuse instead of faulting
    if (((uintptr_t)a.x | (uintptr_t)a.q) & 15u || ((uintptr_t)a.ubin & 3u)) {
        vrq_set_error("encode: x and q must be 16-byte aligned and ubinary 4-byte aligned device pointers");
        return VRQ_ERR_ARG;
    }
    EncParams p{a.x, a.n, a.d, a.limit_f32, a.scale_f32, a.q, a.mn, a.mx, a.ubin, a.ge};
    vrq_timer_scope ts(ctx, VRQ_CAT_ENCODE, st);
    switch (a.codec) {
        case VRQ_CODEC_NONE:
            return launch_codec<VRQ_CODEC_NONE>(ctx, p, st);
        case VRQ_CODEC_INT8_PERDOC:
            return launch_codec<VRQ_CODEC_INT8_PERDOC>(ctx, p, st);
        case VRQ_CODEC_INT8_GLOBAL:
            return launch_codec<VRQ_CODEC_INT8_GLOBAL>(ctx, p, st);
        case VRQ_CODEC_INT16_GLOBAL:
            return launch_codec<VRQ_CODEC_INT16_GLOBAL>(ctx, p, st);
        case VRQ_CODEC_INT4:
            return launch_codec<VRQ_CODEC_INT4>(ctx, p, st);
    }
    vrq_set_error("unknown codec %d", a.codec);
    return VRQ_ERR_ARG;
}

int vrq_launch_to_binary_int(vrq_ctx* ctx, const void* x, int elem_bytes, int64_t n, int d, int ge, uint8_t* ubin,
                             cudaStream_t st) {
    if (n == 0) return 0;
    if (d % 8 != 0 || d < 8) {
        vrq_set_error("embedding_dim must be a positive multiple of 8 (got %d)", d);
        return VRQ_ERR_ARG;
    }
    if ((int64_t)d * 32768 > 2147483647LL) {
        vrq_set_error("embedding_dim %d too large for the integer threshold kernel", d);
        return VRQ_ERR_UNSUPPORTED;
    }
    int64_t blocks = (n + 7) / 8, cap = (int64_t)ctx->sm_count * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    vrq_timer_scope ts(ctx, VRQ_CAT_ENCODE, st);
    if (elem_bytes == 1)
        to_binary_int_kernel<int8_t><<<grid, 256, 0, st>>>(static_cast<const int8_t*>(x), n, d, ge, ubin);
    else
        to_binary_int_kernel<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t*>(x), n, d, ge, ubin);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

int vrq_launch_dequant(vrq_ctx* ctx, const vrq_dequant_args& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    DeqParams p{};
    p.kind = a.kind;
    p.q = a.q;
    p.n = a.n;
    p.d = a.d;
    p.mn = a.mn;
    p.mx = a.mx;
    p.out = a.out;
    if (a.kind == VRQ_PAYLOAD_INT8_GLOBAL) p.scale_f32 = (float)(a.limit / 127.0);
    if (a.kind == VRQ_PAYLOAD_INT16_GLOBAL) p.scale_f32 = (float)(a.limit / 32767.0);
    if (a.kind == VRQ_PAYLOAD_INT4_GLOBAL) p.scale_f64 = a.limit / 7.0;
    int64_t total = a.n * (int64_t)a.d;
    int64_t blocks = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    dequant_kernel<<<grid, 256, 0, st>>>(p);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}
