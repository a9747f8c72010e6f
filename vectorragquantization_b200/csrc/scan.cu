// Phase I: exact Hamming top-k over a flat array of binary codes (replacement for faiss
// IndexBinaryFlat::search as called at CohereEnhancedVectorDB.py:268 / VectorDBInt8.py:218).
//
// Result contract: for every query the k codes with the smallest key (hamming << 40 | position), ascending -
// i.e. ties broken by ascending position, exactly what faiss's (distance, id)-ordered heap + reorder returns.
//
// Kernel design (B200):
//   * grid = (query tiles) x (row strips); one persistent CTA walks its strip in ascending row order.
//   * A producer warp streams 256-row x 128-byte tiles of codes with TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//     into a ring of shared-memory stages guarded by full/empty mbarriers.
//   * 8 consumer warps: thread = one code.  The 128-byte code is pulled from the swizzled stage with 8
//     conflict-free LDS.128 into registers, the stage is released immediately, and the thread then walks the
//     query tile held in shared memory (broadcast LDS.128) doing XOR + POPC + ADD.
//   * Top-k: a per-(strip, query) threshold tau (the k-th best distance seen so far in rows with LOWER positions)
//     filters candidates with a strict '<' - exact under the (distance, position) key because the strip is walked
//     in ascending position.  Survivors are appended to a per-(strip, query) list in global memory; when a list
//     could overflow, the consumer warps radix-select the k smallest keys in shared memory and tighten tau.
//   * A prefix pass over the first rows of the database seeds tau for the main pass, so lists rarely compact.
//   * A per-query merge kernel radix-selects the k smallest keys over all strips' lists and bitonic-sorts them.
// Bounds: <= ~3 queries per pass the kernel is HBM-bound (128 B per code per pass); for query batches it is bound by
// the integer pipes (32 x (LOP3 + POPC + IADD) per (query, code) pair).  See DESIGN.md section 3.
#include "scan_common.cuh"

#ifndef VRQ_MMA_LOCKSTEP_DEFAULT
#define VRQ_MMA_LOCKSTEP_DEFAULT 0
#endif

namespace {

using namespace vrq;

// ---- Hamming distance of one 1024-bit code held in registers against one query in shared memory ---------------
// Plain form: 32 x (XOR, POPC, ADD) - bound by the POPC pipe (16 lanes/clk/SM measured, profiles/microbench).
__device__ __forceinline__ int hamming128_popc(const uint4 (&c)[8], uint32_t qaddr) {
    int d = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 qq = lds128(qaddr + j * 16);
        d += __popc(c[j].x ^ qq.x) + __popc(c[j].y ^ qq.y) + __popc(c[j].z ^ qq.z) + __popc(c[j].w ^ qq.w);
    }
    return d;
}

__device__ __forceinline__ void csa(uint32_t& s, uint32_t& cy, uint32_t a, uint32_t b, uint32_t d) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(d));   // a ^ b ^ d
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(cy) : "r"(a), "r"(b), "r"(d));  // majority(a, b, d)
}

// Carry-save form: 16 carry-save adders (2 LOP3 each, ALU pipe at 64 lanes/clk/SM) fold the 32 XOR words into
// 4 words of weight 1, 10 of weight 2 and 2 of weight 4, so only 16 POPC remain: 64 LOP3 + 16 POPC per pair keeps
// the ALU pipe and the POPC pipe equally busy (~1 clk per pair per SM instead of 2).  Exact: a CSA preserves
// popc(a) + popc(b) + popc(d) = popc(sum) + 2 popc(carry).
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Same carry-save tree, but the 16 population counts are accumulated with integer multiply-adds whose multiplier
// (1, 2 or 4, derived from a kernel argument so ptxas cannot turn them back into IADD3/LEA) sits in a register: the
// adds leave the saturated ALU pipe for the idle FMA pipe.
__device__ __forceinline__ int hamming128_csa_imad(const uint4 (&c)[8], uint32_t qaddr, uint32_t one, uint32_t two, uint32_t four) {
    uint32_t x[32];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 qq = lds128(qaddr + j * 16);
        x[4 * j + 0] = c[j].x ^ qq.x;
        x[4 * j + 1] = c[j].y ^ qq.y;
        x[4 * j + 2] = c[j].z ^ qq.z;
        x[4 * j + 3] = c[j].w ^ qq.w;
    }
    uint32_t s1[10], c1[10];
#pragma unroll
    for (int t = 0; t < 10; t++) csa(s1[t], c1[t], x[3 * t], x[3 * t + 1], x[3 * t + 2]);
    uint32_t s2[4], c2[4];
    csa(s2[0], c2[0], s1[0], s1[1], s1[2]);
    csa(s2[1], c2[1], s1[3], s1[4], s1[5]);
    csa(s2[2], c2[2], s1[6], s1[7], s1[8]);
    csa(s2[3], c2[3], s1[9], x[30], x[31]);
    uint32_t s3[2], c3[2];
    csa(s3[0], c3[0], c1[0], c1[1], c1[2]);
    csa(s3[1], c3[1], c1[3], c1[4], c1[5]);
    uint32_t a0 = imad(__popc(s2[0]), one, __popc(s2[1]));
    uint32_t a1 = imad(__popc(s2[2]), one, __popc(s2[3]));
    uint32_t a2 = imad(__popc(c3[0]), four, 0u);
    uint32_t a3 = imad(__popc(c3[1]), four, 0u);
    a0 = imad(__popc(c1[6]), two, a0);
    a1 = imad(__popc(c1[7]), two, a1);
    a2 = imad(__popc(c1[8]), two, a2);
    a3 = imad(__popc(c1[9]), two, a3);
    a0 = imad(__popc(c2[0]), two, a0);
    a1 = imad(__popc(c2[1]), two, a1);
    a2 = imad(__popc(c2[2]), two, a2);
    a3 = imad(__popc(c2[3]), two, a3);
    a0 = imad(__popc(s3[0]), two, a0);
    a1 = imad(__popc(s3[1]), two, a1);
    a0 = imad(a1, one, a0);
    a2 = imad(a3, one, a2);
    return (int)imad(a2, one, a0);
}

__device__ __forceinline__ int hamming128_csa(const uint4 (&c)[8], uint32_t qaddr) {
    uint32_t x[32];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 qq = lds128(qaddr + j * 16);
        x[4 * j + 0] = c[j].x ^ qq.x;
        x[4 * j + 1] = c[j].y ^ qq.y;
        x[4 * j + 2] = c[j].z ^ qq.z;
        x[4 * j + 3] = c[j].w ^ qq.w;
    }
    uint32_t s1[10], c1[10];
#pragma unroll
    for (int t = 0; t < 10; t++) csa(s1[t], c1[t], x[3 * t], x[3 * t + 1], x[3 * t + 2]);
    uint32_t s2[4], c2[4];
    csa(s2[0], c2[0], s1[0], s1[1], s1[2]);
    csa(s2[1], c2[1], s1[3], s1[4], s1[5]);
    csa(s2[2], c2[2], s1[6], s1[7], s1[8]);
    csa(s2[3], c2[3], s1[9], x[30], x[31]);
    uint32_t s3[2], c3[2];
    csa(s3[0], c3[0], c1[0], c1[1], c1[2]);
    csa(s3[1], c3[1], c1[3], c1[4], c1[5]);
    const int w1 = __popc(s2[0]) + __popc(s2[1]) + __popc(s2[2]) + __popc(s2[3]);
    const int w2 = __popc(c1[6]) + __popc(c1[7]) + __popc(c1[8]) + __popc(c1[9]) + __popc(c2[0]) + __popc(c2[1]) +
                   __popc(c2[2]) + __popc(c2[3]) + __popc(s3[0]) + __popc(s3[1]);
    const int w4 = __popc(c3[0]) + __popc(c3[1]);
    return w1 + 2 * w2 + 4 * w4;
}

// 14-CSA variant: 4 words of weight 1 + 14 of weight 2 -> 18 POPC, 60 LOP3: balances the two pipes when the final
// adds also sit on the ALU pipe.
__device__ __forceinline__ int hamming128_csa14(const uint4 (&c)[8], uint32_t qaddr) {
    uint32_t x[32];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 qq = lds128(qaddr + j * 16);
        x[4 * j + 0] = c[j].x ^ qq.x;
        x[4 * j + 1] = c[j].y ^ qq.y;
        x[4 * j + 2] = c[j].z ^ qq.z;
        x[4 * j + 3] = c[j].w ^ qq.w;
    }
    uint32_t s1[10], c1[10];
#pragma unroll
    for (int t = 0; t < 10; t++) csa(s1[t], c1[t], x[3 * t], x[3 * t + 1], x[3 * t + 2]);
    uint32_t s2[4], c2[4];
    csa(s2[0], c2[0], s1[0], s1[1], s1[2]);
    csa(s2[1], c2[1], s1[3], s1[4], s1[5]);
    csa(s2[2], c2[2], s1[6], s1[7], s1[8]);
    csa(s2[3], c2[3], s1[9], x[30], x[31]);
    const int w1 = __popc(s2[0]) + __popc(s2[1]) + __popc(s2[2]) + __popc(s2[3]);
    int w2 = __popc(c2[0]) + __popc(c2[1]) + __popc(c2[2]) + __popc(c2[3]);
#pragma unroll
    for (int t = 0; t < 10; t++) w2 += __popc(c1[t]);
    return w1 + 2 * w2;
}

// The append is the cold path (a few candidates per million pairs once tau has converged): keep it out of line so
// the hot loop carries no address arithmetic for it.
__device__ __noinline__ void append_candidate(int* cnt_s, uint64_t* lists, int q, int cap, int d, unsigned long long pos,
                                              const unsigned long long* key_lo) {
    if (key_lo && (((unsigned long long)d << VRQ_KEY_POS_BITS) | pos) <= key_lo[q]) return;  // below the chunk's lower bound
    const int slot = atomicAdd(&cnt_s[q], 1);
    if (slot >= cap) __trap();  // cannot happen (overflow check every group_tiles tiles); never write past a list
    lists[(size_t)q * cap + slot] = ((unsigned long long)d << VRQ_KEY_POS_BITS) | pos;
}

// ---- the scan kernel ---------------------------------------------------------------------------------------
// TMA128 = true : code_bytes == 128, TMA + swizzled shared-memory pipeline (the fast path)
// TMA128 = false: any code_bytes % 4 == 0, codes read straight from global memory (correct, not tuned)
template <bool TMA128, int CW, int CSA>
__global__ void __launch_bounds__(ScanCfg<CW>::THREADS, 1)
hamming_scan_kernel(const __grid_constant__ CUtensorMap tmap, ScanParams p) {
    constexpr int CONSUMER_WARPS = CW;
    constexpr int CONSUMER_THREADS = ScanCfg<CW>::CONSUMER_THREADS;
    constexpr int SCAN_THREADS = ScanCfg<CW>::THREADS;
    constexpr int TILE_ROWS = ScanCfg<CW>::TILE_ROWS;
    constexpr int STAGE_BYTES = ScanCfg<CW>::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    // carve shared memory: [stages | query codes | tau | cnt | scratch | select scratch | barriers]
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_mem = base;
    uint8_t* ptr = base + (TMA128 ? (size_t)p.stages * STAGE_BYTES : 0);
    uint8_t* qsm = ptr;
    ptr += (size_t)p.qtile * p.code_bytes;
    ptr = (uint8_t*)(((uintptr_t)ptr + 15) & ~(uintptr_t)15);
    int* tau_s = (int*)ptr;
    ptr += sizeof(int) * p.qtile;
    int* cnt_s = (int*)ptr;
    ptr += sizeof(int) * p.qtile;
    ptr = (uint8_t*)(((uintptr_t)ptr + 15) & ~(uintptr_t)15);
    unsigned long long* scratch = (unsigned long long*)ptr;
    ptr += sizeof(unsigned long long) * p.cap;
    SelectScratch* sc = (SelectScratch*)ptr;
    ptr += sizeof(SelectScratch);
    ptr = (uint8_t*)(((uintptr_t)ptr + 7) & ~(uintptr_t)7);
    unsigned long long* bars = (unsigned long long*)ptr;  // full[stages], empty[stages]
    int* flag = (int*)(bars + 2 * 8);

    if (p.guard && *p.guard == 0) return;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * p.qtile;
    const int qt = min(p.qtile, p.nq - q0);
    const int strip = blockIdx.y;
    const int64_t s_begin = p.row_begin + (int64_t)strip * p.rows_per_strip;
    const int64_t s_end = min(p.row_end, s_begin + p.rows_per_strip);
    const int64_t nrows = s_end > s_begin ? s_end - s_begin : 0;
    const int ntiles = (int)((nrows + TILE_ROWS - 1) / TILE_ROWS);

    if (TMA128 && tid == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[8 + s]), CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // query tile -> shared memory; thresholds and counters
    if ((p.code_bytes & 3) == 0) {
        for (int i = tid; i < p.qtile * p.code_bytes / 4; i += SCAN_THREADS) {
            const int q = i / (p.code_bytes / 4);
            uint32_t v = 0;
            if (q < qt) v = reinterpret_cast<const uint32_t*>(p.queries + (size_t)q0 * p.code_bytes)[i];
            reinterpret_cast<uint32_t*>(qsm)[i] = v;
        }
    } else {  // d % 32 != 0 (faiss only asks for d % 8 == 0): rows are not word-aligned, everything goes byte by byte
        for (int i = tid; i < p.qtile * p.code_bytes; i += SCAN_THREADS)
            qsm[i] = (i / p.code_bytes < qt) ? p.queries[(size_t)q0 * p.code_bytes + i] : (uint8_t)0;
    }
    for (int q = tid; q < p.qtile; q += SCAN_THREADS) {
        tau_s[q] = (q < qt) ? (p.tau0 ? min(p.tau0[q0 + q], TAU_INF) : TAU_INF) : 0;
        cnt_s[q] = 0;
    }
    if (tid == 0) *flag = 0;
    __syncthreads();

    uint64_t* my_lists = p.lists + ((size_t)strip * p.nq + q0) * p.cap;
    const unsigned long long* key_lo_q0 = p.key_lo ? p.key_lo + q0 : nullptr;

    if (warp == CONSUMER_WARPS) {
        // ===================== TMA producer warp =====================
        if (TMA128 && lane == 0) {
            for (int t = 0; t < ntiles; t++) {
                const int s = t % p.stages;
                const uint32_t ph = (uint32_t)(t / p.stages) & 1u;
                mbar_wait(smem_u32(&bars[8 + s]), ph ^ 1u);
                mbar_expect_tx(smem_u32(&bars[s]), STAGE_BYTES);
#pragma unroll
                for (int b = 0; b < TILE_ROWS / TMA_BOX_ROWS; b++)
                    tma_load_2d(smem_u32(stage_mem + (size_t)s * STAGE_BYTES + (size_t)b * TMA_BOX_ROWS * CODE_BYTES), &tmap, 0,
                                (int)(s_begin + (int64_t)t * TILE_ROWS + b * TMA_BOX_ROWS), smem_u32(&bars[s]));
            }
        }
    } else {
        // ===================== consumer warps: thread = one code =====================
        const int r = warp * 32 + lane;
        const int w_words = p.code_bytes / 4;
        for (int t = 0; t < ntiles; t++) {
            const int64_t lrow = s_begin + (int64_t)t * TILE_ROWS + r;
            const bool valid = lrow < s_end;
            const unsigned long long pos = (unsigned long long)(p.pos_base + lrow);
            if (TMA128) {
                const int s = t % p.stages;
                mbar_wait(smem_u32(&bars[s]), (uint32_t)(t / p.stages) & 1u);
                const uint32_t rowaddr = smem_u32(stage_mem + (size_t)s * STAGE_BYTES) + r * CODE_BYTES;
                uint4 c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) c[j] = lds128(rowaddr + ((j ^ (r & 7)) << 4));
                const uint32_t qbase = smem_u32(qsm), taubase = smem_u32(tau_s);
                uint32_t qaddr = qbase, taddr = taubase;
                const uint32_t one = (uint32_t)p.one, two = one + one, four = two + two;
                // rows past the end of the strip (last tile only) sit the loop out: no per-query validity arithmetic
                if (valid) {
#pragma unroll 4
                for (int q = 0; q < qt; q++, qaddr += CODE_BYTES, taddr += 4) {
                    const int d = (CSA == 17   ? hamming128_csa_imad(c, qaddr, one, two, four)
                                   : CSA == 16 ? hamming128_csa(c, qaddr)
                                   : CSA == 14 ? hamming128_csa14(c, qaddr)
                                               : hamming128_popc(c, qaddr));
                    if (d < lds32(taddr)) append_candidate(cnt_s, my_lists, q, p.cap, d, pos, key_lo_q0);
                }
                }
                // The stage goes back to the producer only here: ld.shared merely ISSUES the read, and an mbarrier
                // arrive placed right behind it can overtake the outstanding loads (observed on B200 in scan_mma.cu:
                // TMA refilled the stage before the reads returned).  After the query loop every c[j] has been an
                // operand of real instructions, so the reads have completed.  (Rows past the end of the strip never
                // use their registers.)
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[8 + s]));
            } else {
                const uint8_t* crow8 = p.codes + (size_t)(valid ? lrow : 0) * p.code_bytes;
                const uint32_t* crow = reinterpret_cast<const uint32_t*>(crow8);
                const bool words_ok = (p.code_bytes & 3) == 0;
                for (int q = 0; q < qt; q++) {
                    const uint8_t* qrow8 = qsm + (size_t)q * p.code_bytes;
                    const uint32_t* qrow = reinterpret_cast<const uint32_t*>(qrow8);
                    int d = 0;
                    if (words_ok) {
                        for (int w = 0; w < w_words; w++) d += __popc(__ldg(crow + w) ^ qrow[w]);
                    } else {
                        for (int b = 0; b < p.code_bytes; b++) d += __popc((uint32_t)(__ldg(crow8 + b) ^ qrow8[b]));
                    }
                    if (d < tau_s[q] && valid) append_candidate(cnt_s, my_lists, q, p.cap, d, pos, key_lo_q0);
                }
            }
            // ---- overflow check every group_tiles tiles: no list may exceed cap during the next group ----
            if ((t + 1) % p.group_tiles == 0 && t + 1 < ntiles) {
                group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
                const int limit = p.cap - p.group_tiles * TILE_ROWS;
                for (int q = tid; q < qt; q += CONSUMER_THREADS)
                    if (cnt_s[q] > limit) atomicOr(flag, 1);
                group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
                if (*flag) {
                    for (int q = 0; q < qt; q++) {
                        const int n = cnt_s[q];
                        if (n > limit) compact_list<CONSUMER_THREADS>(my_lists + (size_t)q * p.cap, n, p.k, scratch, sc, tid, &cnt_s[q], &tau_s[q]);
                    }
                    group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
                    if (tid == 0) *flag = 0;
                    group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
                }
            }
        }
        group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
        // final compaction: every list leaves the kernel with at most k keys (bounds the merge's working set)
        for (int q = 0; q < qt; q++) {
            const int n = cnt_s[q];
            if (n > p.k) compact_list<CONSUMER_THREADS>(my_lists + (size_t)q * p.cap, n, p.k, scratch, sc, tid, &cnt_s[q], &tau_s[q]);
        }
        group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
        for (int q = tid; q < qt; q += CONSUMER_THREADS) p.counts[(size_t)strip * p.nq + q0 + q] = cnt_s[q];
    }
}

// ---- per-query merge of the strips' lists: k smallest keys, sorted ------------------------------------------
// Tree of shared-memory merges: CTA (g, q) copies the lists of strips [g*gs, (g+1)*gs) of query q into shared memory
// (every list holds <= k keys when the scan kernel exits), radix-selects the k smallest there, and either writes them
// as one list of the next level or - at the last level - bitonic-sorts them into the final result.
constexpr int MERGE_THREADS = 512;

__global__ void __launch_bounds__(MERGE_THREADS) merge_group_kernel(const uint64_t* __restrict__ lists,
                                                                    const int* __restrict__ counts, int nq, int cap_in,
                                                                    int s_total, int gs, int k, int buf_cap,
                                                                    uint64_t* __restrict__ out_lists, int* __restrict__ out_counts,
                                                                    uint64_t* __restrict__ final_out, int* __restrict__ tau_out,
                                                                    const int* __restrict__ guard) {
    if (guard && *guard == 0) return;
    extern __shared__ unsigned long long msm[];  // buf[buf_cap] | sel[next_pow2(k)] (final level only)
    __shared__ SelectScratch sc;
    __shared__ int offs[65];
    const int g = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
    const int s0 = g * gs, s1 = min(s_total, s0 + gs), ns = s1 - s0;
    unsigned long long* buf = msm;
    unsigned long long* sel = msm + buf_cap;
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < ns; i++) {
            offs[i] = acc;
            acc += counts[(size_t)(s0 + i) * nq + q];
        }
        offs[ns] = acc;
    }
    __syncthreads();
    const int total = offs[ns];
    if (total > buf_cap) __trap();  // cannot happen: lists are compacted to <= k keys by the scan kernel
    {
        // warp w copies lists w, w + 16, ...; 4 independent 8-byte loads in flight per lane (the copy is latency-bound)
        const int warp = tid >> 5, lane = tid & 31;
        for (int i = warp; i < ns; i += MERGE_THREADS / 32) {
            const uint64_t* l = lists + ((size_t)(s0 + i) * nq + q) * cap_in;
            unsigned long long* dst = buf + offs[i];
            const int c = offs[i + 1] - offs[i];
            int j = lane;
            for (; j + 96 < c; j += 128) {
                const unsigned long long a0 = l[j], a1 = l[j + 32], a2 = l[j + 64], a3 = l[j + 96];
                dst[j] = a0;
                dst[j + 32] = a1;
                dst[j + 64] = a2;
                dst[j + 96] = a3;
            }
            for (; j < c; j += 32) dst[j] = l[j];
        }
    }
    __syncthreads();
    unsigned long long kth = VRQ_KEY_NONE;
    if (total > k) {
        kth = radix_select_kth<MERGE_THREADS>([&](int i) { return buf[i]; }, total, k, tid, &sc, 0);
    }
    if (tid == 0) sc.counter = 0;
    if (final_out) {
        const int n2 = next_pow2(k);
        for (int i = tid; i < n2; i += MERGE_THREADS) sel[i] = VRQ_KEY_NONE;
        __syncthreads();
        for (int i = tid; i < total; i += MERGE_THREADS) {
            const unsigned long long key = buf[i];
            if (key <= kth) sel[atomicAdd(&sc.counter, 1)] = key;
        }
        bitonic_sort<MERGE_THREADS, false>(sel, nullptr, n2, tid, 0);
        for (int i = tid; i < k; i += MERGE_THREADS) final_out[(size_t)q * k + i] = sel[i];
        if (tau_out && tid == 0) {
            // threshold for a following pass over rows with HIGHER positions: strict '<' against the k-th best distance
            const unsigned long long last = sel[k - 1];
            tau_out[q] = (last == VRQ_KEY_NONE) ? 0x7fffffff : (int)(last >> VRQ_KEY_POS_BITS);
        }
    } else {
        __syncthreads();
        uint64_t* o = out_lists + ((size_t)g * nq + q) * k;
        for (int i = tid; i < total; i += MERGE_THREADS) {
            const unsigned long long key = buf[i];
            if (key <= kth) o[atomicAdd(&sc.counter, 1)] = key;
        }
        if (tid == 0) out_counts[(size_t)g * nq + q] = total < k ? total : k;
    }
}

// After a pass whose thresholds came from a sample: does every query hold at least `need` candidates over all strips?
// (A list that was compacted holds k >= need keys on its own.)  Sets *flag otherwise - the exact fallback pass runs.
// A query that came up short gets an infinite threshold for the fallback pass; the others keep theirs, so the fallback
// floods only the lists of the queries that need it.
__global__ void verify_counts_kernel(const int* __restrict__ counts, int strips, int nq, int need, int* flag, int* tau) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    long long total = 0;
    for (int s = 0; s < strips; s++) total += counts[(size_t)s * nq + q];
    if (total < need) {
        atomicOr(flag, 1);
        tau[q] = 0x7fffffff;
    }
}

// List-free sample pass (scan_mma.cu): tau[q] = the k'-th smallest of the distances the epilogue threads kept for query q
// (strips x 2 threads x 4 values, 0xFFFF = none) - the largest one kept when there are fewer than k', "no threshold" when
// there are none.
__global__ void __launch_bounds__(128) sample_tau_kernel(const unsigned short* __restrict__ sample, int strips, int nq, int kp, int* __restrict__ tau) {
    __shared__ int hist[4][1026];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + warp;
    for (int i = lane; i < 1026; i += 32) hist[warp][i] = 0;
    __syncwarp();
    if (q < nq) {
        for (int s = 0; s < strips; s++) {
            const unsigned short* src = sample + ((size_t)s * nq + q) * 8;
            if (lane < 8) {
                const int v = src[lane];
                if (v <= 1024) atomicAdd(&hist[warp][v], 1);
            }
        }
    }
    __syncwarp();
    if (q < nq && lane == 0) {
        int acc = 0, t = 0x7fffffff;
        for (int d = 0; d <= 1024; d++) {
            if (hist[warp][d]) {
                acc += hist[warp][d];
                t = d;  // the largest kept distance so far: the answer when fewer than kp were kept
                if (acc >= kp) break;
            }
        }
        tau[q] = t;
    }
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- host side ----------------------------------------------------------------------------------------------
struct ScanPlan {
    int cw;  // consumer warps (8 or 16)
    int csa;  // carry-save popcount (1) or plain XOR+POPC (0)
    int qtile, qtiles, strips, group_tiles, cap, stages;
    int64_t rows_per_strip;
    size_t smem;
    int tile_rows() const { return cw * 32; }
};

size_t scan_smem_bytes(bool tma, int cw, int stages, int qtile, int code_bytes, int cap) {
    size_t b = 1024;  // alignment slack
    if (tma) b += (size_t)stages * cw * 32 * CODE_BYTES;
    b += (size_t)qtile * code_bytes + 16;
    b += sizeof(int) * 2 * (size_t)qtile + 16;
    b += sizeof(unsigned long long) * (size_t)cap;
    b += sizeof(SelectScratch) + 8;
    b += sizeof(unsigned long long) * 16 + 16;
    return b;
}

int plan_scan(vrq_ctx* ctx, bool tma, int code_bytes, int64_t rows, int nq, int k, ScanPlan* pl) {
    const int sms = ctx->sm_count;
    const bool stream_regime = nq <= 8;
    // tuning knobs (defaults chosen from the measurements in profiles/): VRQ_SCAN_CW_STREAM / VRQ_SCAN_CW_BATCH
    pl->cw = tma ? (stream_regime ? env_int("VRQ_SCAN_CW_STREAM", 16) : env_int("VRQ_SCAN_CW_BATCH", 16)) : 8;
    if (pl->cw != 16) pl->cw = 8;
    pl->csa = env_int("VRQ_SCAN_CSA", 16);  // 0 = plain XOR+POPC, 14 / 16 = number of carry-save adders per pair
    if (pl->csa != 14 && pl->csa != 16 && pl->csa != 17) pl->csa = 0;  // 17 = 16 CSAs + IMAD accumulation
    const int tile_rows = pl->tile_rows();
    int max_qtile = 256;
    if ((size_t)max_qtile * code_bytes > 32 * 1024) max_qtile = (int)(32 * 1024 / code_bytes);
    if (max_qtile < 1) max_qtile = 1;
    pl->qtiles = (nq + max_qtile - 1) / max_qtile;
    pl->qtile = (nq + pl->qtiles - 1) / pl->qtiles;  // balanced tiles
    if (!stream_regime && pl->qtile < max_qtile) pl->qtile = ((pl->qtile + 7) / 8) * 8 < max_qtile ? ((pl->qtile + 7) / 8) * 8 : max_qtile;
    pl->qtiles = (nq + pl->qtile - 1) / pl->qtile;
    // tiles between overflow checks (one named barrier each): few queries per tile -> amortise over more tiles
    pl->group_tiles = env_int("VRQ_SCAN_GROUP_TILES", pl->qtile <= 8 ? 4 : (pl->qtile <= 16 ? 2 : 1));
    if (pl->group_tiles < 1) pl->group_tiles = 1;
    const int slack = k < 256 ? 256 : (k > 2048 ? 2048 : k);
    pl->cap = k + slack + pl->group_tiles * tile_rows;
    int strips = sms / pl->qtiles;
    if (strips < 1) strips = 1;
    int64_t tiles = (rows + tile_rows - 1) / tile_rows;
    if (tiles < 1) tiles = 1;
    if (strips > tiles) strips = (int)tiles;
    int64_t tps = (tiles + strips - 1) / strips;
    pl->rows_per_strip = tps * tile_rows;
    pl->strips = (int)((tiles + tps - 1) / tps);
    const size_t limit = ctx->smem_optin ? ctx->smem_optin : 227 * 1024;
    pl->stages = 0;
    if (tma) {
        const int max_stages = env_int("VRQ_SCAN_STAGES", 8);
        for (int s = max_stages > 8 ? 8 : max_stages; s >= 2; s--) {
            if (scan_smem_bytes(true, pl->cw, s, pl->qtile, code_bytes, pl->cap) <= limit) {
                pl->stages = s;
                break;
            }
        }
        if (pl->stages == 0) {
            vrq_set_error("Hamming top-k with k=%d does not fit the shared-memory plan", k);
            return VRQ_ERR_UNSUPPORTED;
        }
    }
    pl->smem = scan_smem_bytes(tma, pl->cw, pl->stages, pl->qtile, code_bytes, pl->cap);
    if (pl->smem > limit) {
        vrq_set_error("Hamming top-k with k=%d does not fit the shared-memory plan", k);
        return VRQ_ERR_UNSUPPORTED;
    }
    return 0;
}

template <bool TMA, int CW, int CSA>
int launch_scan_t(const CUtensorMap& tmap, const ScanParams& sp, const ScanPlan& pl, cudaStream_t st) {
    dim3 grid(pl.qtiles, pl.strips);
    VRQ_CUDA(cudaFuncSetAttribute(hamming_scan_kernel<TMA, CW, CSA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    hamming_scan_kernel<TMA, CW, CSA><<<grid, ScanCfg<CW>::THREADS, pl.smem, st>>>(tmap, sp);
    return 0;
}

int launch_scan(vrq_ctx* ctx, bool tma, const CUtensorMap& tmap, const ScanParams& sp, const ScanPlan& pl, cudaStream_t st) {
    if (!tma)
        VRQ_TRY((launch_scan_t<false, 8, 0>(tmap, sp, pl, st)));
    else if (pl.cw == 16 && pl.csa == 17)
        VRQ_TRY((launch_scan_t<true, 16, 17>(tmap, sp, pl, st)));
    else if (pl.cw == 8 && pl.csa == 17)
        VRQ_TRY((launch_scan_t<true, 8, 17>(tmap, sp, pl, st)));
    else if (pl.cw == 16 && pl.csa == 16)
        VRQ_TRY((launch_scan_t<true, 16, 16>(tmap, sp, pl, st)));
    else if (pl.cw == 16 && pl.csa == 14)
        VRQ_TRY((launch_scan_t<true, 16, 14>(tmap, sp, pl, st)));
    else if (pl.cw == 16)
        VRQ_TRY((launch_scan_t<true, 16, 0>(tmap, sp, pl, st)));
    else if (pl.csa == 16)
        VRQ_TRY((launch_scan_t<true, 8, 16>(tmap, sp, pl, st)));
    else if (pl.csa == 14)
        VRQ_TRY((launch_scan_t<true, 8, 14>(tmap, sp, pl, st)));
    else
        VRQ_TRY((launch_scan_t<true, 8, 0>(tmap, sp, pl, st)));
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

// Merge tree over `strips` lists per query (each holding <= k keys, row stride cap): returns sorted keys in out.
// list_max: upper bound of the keys one input list holds (k when the scan kernel compacted its lists; more after a
// threshold-only pass).
int launch_merge(vrq_ctx* ctx, const uint64_t* lists, const int* counts, int strips, int nq, int cap, int k, uint64_t* out,
                 int* tau_out, cudaStream_t st, const int* guard = nullptr, int list_max = 0) {
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    if (list_max < k) list_max = k;
    const size_t budget = 200 * 1024;
    int gs = (int)((budget - sizeof(unsigned long long) * (size_t)n2) / (sizeof(unsigned long long) * (size_t)list_max));
    if (gs > 64) gs = 64;
    if (gs < 2) {
        vrq_set_error("merge: k=%d too large for the shared-memory merge", k);
        return VRQ_ERR_UNSUPPORTED;
    }
    const uint64_t* cur_lists = lists;
    const int* cur_counts = counts;
    int cur_strips = strips, cur_cap = cap, level = 0;
    while (true) {
        const bool final_level = cur_strips <= gs;
        const int g = final_level ? 1 : (cur_strips + gs - 1) / gs;
        const int ns = final_level ? cur_strips : gs;
        const int buf_cap = ns * (level == 0 ? list_max : k);
        const size_t smem = sizeof(unsigned long long) * ((size_t)buf_cap + (final_level ? (size_t)n2 : 0));
        VRQ_CUDA(cudaFuncSetAttribute(merge_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget + 8192)));
        uint64_t* nxt_lists = nullptr;
        int* nxt_counts = nullptr;
        if (!final_level) {
            void *a, *b;
            VRQ_TRY(vrq_ws_get(ctx, (level & 1) ? VRQ_WS_MERGE_B : VRQ_WS_MERGE_A, sizeof(uint64_t) * (size_t)g * nq * k, &a));
            VRQ_TRY(vrq_ws_get(ctx, (level & 1) ? VRQ_WS_MERGE_CB : VRQ_WS_MERGE_CA, sizeof(int) * (size_t)g * nq, &b));
            nxt_lists = (uint64_t*)a;
            nxt_counts = (int*)b;
        }
        merge_group_kernel<<<dim3(g, nq), MERGE_THREADS, smem, st>>>(cur_lists, cur_counts, nq, cur_cap, cur_strips, ns, k, buf_cap,
                                                                     nxt_lists, nxt_counts, final_level ? out : nullptr,
                                                                     final_level ? tau_out : nullptr, guard);
        vrq_count_launch(ctx);
        VRQ_CUDA(cudaGetLastError());
        if (final_level) break;
        cur_lists = nxt_lists;
        cur_counts = nxt_counts;
        cur_strips = g;
        cur_cap = k;
        level++;
    }
    return 0;
}

}  // namespace

// Plan of one scan launch on either kernel: the integer-pipe kernel above or the tensor-core kernel of scan_mma.cu.
struct PassPlan {
    bool mma = false;
    ScanPlan sp{};
    MmaPlan mp{};
    int strips() const { return mma ? mp.strips : sp.strips; }
    int cap() const { return mma ? mp.cap : sp.cap; }
    int tile_rows() const { return mma ? MMA_TILE_ROWS : sp.tile_rows(); }
    int64_t rows_per_strip() const { return mma ? mp.rows_per_strip : sp.rows_per_strip; }
    void set_cap(bool tma, int code_bytes, int cap) {
        if (mma) {
            mma_plan_set_cap(&mp, cap);
        } else {
            sp.cap = cap;
            sp.smem = scan_smem_bytes(tma, sp.cw, sp.stages, sp.qtile, code_bytes, cap);
        }
    }
};

static int plan_pass(vrq_ctx* ctx, bool tma, bool mma, int code_bytes, int64_t rows, int nq, int k, PassPlan* pl, bool allow_few = true) {
    pl->mma = mma;
    if (mma) return plan_scan_mma(ctx, rows, nq, k, &pl->mp, allow_few);
    return plan_scan(ctx, tma, code_bytes, rows, nq, k, &pl->sp);
}

static int launch_pass(vrq_ctx* ctx, bool tma, const CUtensorMap& tmap, const CUtensorMap* tmap_mma, ScanParams sp, const PassPlan& pl,
                       cudaStream_t st) {
    sp.num_strips = pl.strips();
    sp.rows_per_strip = pl.rows_per_strip();
    sp.cap = pl.cap();
    if (pl.mma) {
        sp.qtile = 128;
        sp.group_tiles = pl.mp.group_tiles;
        sp.seg_cols = pl.mp.seg_cols;
        sp.seg_full = pl.mp.seg_full;
        sp.seg_tail = pl.mp.seg_tail;
        if (sp.run_stride == 0) {  // dense scan of [row_begin, row_end)
            sp.run_stride = MMA_TILE_ROWS;
            sp.run_shift = 0;
            sp.total_tiles = sp.row_end > sp.row_begin ? (sp.row_end - sp.row_begin + MMA_TILE_ROWS - 1) / MMA_TILE_ROWS : 0;
        }
        return launch_scan_mma(ctx, tmap_mma[0], tmap_mma[1], sp, pl.mp, st);
    }
    sp.qtile = pl.sp.qtile;
    sp.group_tiles = pl.sp.group_tiles;
    sp.stages = pl.sp.stages;
    return launch_scan(ctx, tma, tmap, sp, pl.sp, st);
}

// One batch of queries (nq <= 1024): optional prefix pass, main pass, merge.
// key_lo (nullable, [nq]): only keys strictly greater than key_lo[q] count - the chunk that follows `already` keys returned by
// earlier calls (vrq_hamming_topk_dev walks a top-k larger than VRQ_PASS_K in chunks).
static int topk_batch(vrq_ctx* ctx, const uint8_t* codes, int64_t n, int code_bytes, int64_t pos_base,
                      const uint8_t* q_dev, int nq, int k, uint64_t* keys_out, int32_t* dbg, cudaStream_t st,
                      const unsigned long long* key_lo = nullptr, int64_t already = 0) {
    const bool tma = (code_bytes == CODE_BYTES) && ((uintptr_t)codes % 16 == 0) && n > 0;
    // >= 4 queries per pass go to the tensor cores (scan_mma.cu): measured crossover at 100 M rows (profiles/r01/
    // scan_regime_sweep_final_100M.txt: 4 queries 2.5 ms vs 2.9 ms); fewer are HBM-bound on the integer pipes (scan.cu)
    const int mma_mode = env_int("VRQ_SCAN_MMA", 1);
    const bool mma = tma && mma_mode != 0 && (nq >= env_int("VRQ_SCAN_MMA_MIN_NQ", 3) || mma_mode == 2) && k + 2048 + 256 <= 8192;
    if (dbg && !mma) {
        vrq_set_error("the distance dump is only available on the tensor-core scan path");
        return VRQ_ERR_UNSUPPORTED;
    }
    CUtensorMap tmap, tmap_mma[2];  // tmap_mma: 128-row boxes (one CTA per tile), 64-row boxes (CTA pairs)
    memset(&tmap, 0, sizeof(tmap));
    memset(tmap_mma, 0, sizeof(tmap_mma));
    if (tma) VRQ_TRY(make_codes_tmap(codes, n, TMA_BOX_ROWS, &tmap));
    if (mma) VRQ_TRY(make_codes_tmap(codes, n, MMA_TILE_ROWS, &tmap_mma[0]));
    if (mma) VRQ_TRY(make_codes_tmap(codes, n, MMA_TILE_ROWS / 2, &tmap_mma[1]));

    if (mma) {
        // ---- thresholds from a strided sample (tensor-core path) ------------------------------------------------
        // The k'-th best distance T over a sample of m rows bounds the number of rows with d <= T in the whole database
        // at about k' n / m; m is chosen so that this is `safety` x k.  One dense pass with tau = T + 1 then collects a
        // few thousand candidates per query instead of k per strip, with no list compaction.  The result stays exact:
        // a verification kernel checks that every query collected >= min(k, n) candidates and otherwise raises a flag
        // that un-gates an exact fallback pass (tau of the short queries starts at infinity) enqueued right behind.
        const int kp = env_int("VRQ_MMA_SAMPLE_K", 32);
        // safety = expected candidates per query / k.  The count at the sampled threshold is k' x Gamma(k') / k' distributed: with
        // k' = 32 it falls below 1/4 of its expectation with probability ~1e-9, so 4 x k expected candidates never trip the
        // fallback in practice, and fewer candidates mean fewer survivor visits in the dense pass (measured: 8 -> 4 = +1.4 % QPS,
        // 16 = -4 %; profiles/r02/bench knob sweep of GPU call 31)
        const int safety = env_int("VRQ_MMA_SAFETY", 4);
        const int64_t total_tiles = (n + MMA_TILE_ROWS - 1) / MMA_TILE_ROWS;
        int64_t sample_tiles = 0;
        if (safety > 0 && kp > 0 && kp <= k * safety) {
            const int64_t want_rows = (int64_t)(((double)kp * (double)n) / ((double)safety * (double)k)) + 1;
            sample_tiles = (want_rows + MMA_TILE_ROWS - 1) / MMA_TILE_ROWS;
            if (sample_tiles * 16 > total_tiles || sample_tiles * MMA_TILE_ROWS < (int64_t)64 * kp) sample_tiles = 0;
        }
        if (sample_tiles > 0) {
            // the sample = runs of 16 consecutive tiles (one 32 KB stretch of codes) spread evenly over the database: single
            // tiles 4 MB apart cost a fresh DRAM page / TLB entry per 16 KB (measured: 2.2 ms for 0.4 % of the rows)
            const int run_shift = 4;
            const int64_t runs = (sample_tiles + 15) / 16;
            const int64_t run_stride = (total_tiles / runs) * MMA_TILE_ROWS;  // >= 16 tiles: runs never overlap
            const int64_t actual_tiles = runs * 16;
            PassPlan s_pl, m_pl;
            // the sample pass always runs on the 128-query-tile kernel in its list-free form (a few dozen microseconds); the
            // swapped-operand kernels would sample through lists that flood until the first compaction (0.2 - 0.5 ms)
            VRQ_TRY(plan_pass(ctx, tma, true, code_bytes, actual_tiles * MMA_TILE_ROWS, nq, kp, &s_pl, env_int("VRQ_MMA_SAMPLE_FEW", 0) != 0));
            VRQ_TRY(plan_pass(ctx, tma, true, code_bytes, n, nq, k, &m_pl));
            const int cap = s_pl.cap() > m_pl.cap() ? s_pl.cap() : m_pl.cap();
            s_pl.set_cap(tma, code_bytes, cap);
            m_pl.set_cap(tma, code_bytes, cap);
            const int max_strips = s_pl.strips() > m_pl.strips() ? s_pl.strips() : m_pl.strips();
            void *lists_v, *counts_v, *tau_v, *skeys_v, *flag_v;
            VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_LISTS, sizeof(uint64_t) * (size_t)max_strips * nq * cap, &lists_v));
            VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_COUNTS, sizeof(int) * (size_t)max_strips * nq, &counts_v));
            VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_TAU, sizeof(int) * (size_t)nq, &tau_v));
            VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SAMPLE_KEYS, sizeof(uint64_t) * (size_t)nq * kp, &skeys_v));
            VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_FLAG, sizeof(int) * 4, &flag_v));
            uint64_t* lists = (uint64_t*)lists_v;
            int* counts = (int*)counts_v;
            int* tau = (int*)tau_v;
            int* flag = (int*)flag_v;
            ScanParams sp{};
            sp.codes = codes;
            sp.code_bytes = code_bytes;
            sp.pos_base = pos_base;
            sp.queries = q_dev;
            sp.nq = nq;
            sp.lists = lists;
            sp.counts = counts;
            sp.dbg_stride = n;
            sp.one = 1;
            sp.key_lo = key_lo;
            sp.row_begin = 0;
            sp.row_end = n;
            vrq_timer_scope ts(ctx, VRQ_CAT_SCAN, st);
            // 1. sample pass: exact top-k' of the sampled tiles -> tau[q] = k'-th best distance
            sp.k = kp;
            sp.run_stride = run_stride;
            sp.run_shift = run_shift;
            sp.total_tiles = actual_tiles;
            sp.compact_limit = kp + 256;  // thresholds start at infinity: tighten them after the first two tiles
            sp.sample_mode = 1;           // only tau[q] matters: distance-only compaction, lists left uncompacted
            s_pl.mp.group_tiles = 2;
            // the 128-query-tile kernel keeps the k' smallest distances per epilogue thread instead of lists (no flooding, no
            // compaction: 0.83 -> ~0.1 ms at 100 M rows); the few-queries kernels (lane = database row) keep the list form
            const bool list_free = !s_pl.mp.few && env_int("VRQ_MMA_SAMPLE_LISTS", 0) == 0;
            if (list_free) {
                void* so_v;
                VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SAMPLE_D, sizeof(unsigned short) * (size_t)s_pl.strips() * nq * 8, &so_v));
                sp.sample_out = (unsigned short*)so_v;
            }
            VRQ_TRY(launch_pass(ctx, tma, tmap, tmap_mma, sp, s_pl, st));
            const int sample_list_max = sp.compact_limit + s_pl.mp.group_tiles * MMA_TILE_ROWS;
            sp.compact_limit = 0;
            sp.sample_mode = 0;
            if (list_free) {
                sample_tau_kernel<<<(nq + 3) / 4, 128, 0, st>>>(sp.sample_out, s_pl.strips(), nq, kp, tau);
                vrq_count_launch(ctx);
                VRQ_CUDA(cudaGetLastError());
                sp.sample_out = nullptr;
            } else {
                VRQ_TRY(launch_merge(ctx, lists, counts, s_pl.strips(), nq, cap, kp, (uint64_t*)skeys_v, tau, st, nullptr, sample_list_max));
            }
            // 2. dense pass with the inclusive threshold d <= T
            sp.k = k;
            sp.run_stride = 0;  // dense
            sp.tau0 = tau;
            sp.tau_bias = 1;
            sp.dbg = dbg;
            const int lock_window = env_int("VRQ_MMA_LOCKSTEP", VRQ_MMA_LOCKSTEP_DEFAULT);
            if (lock_window > 0 && m_pl.mp.pair && m_pl.mp.qtiles > 2) {
                void* prog_v;
                const size_t pbytes = sizeof(int) * (size_t)m_pl.strips() * (size_t)(m_pl.mp.qtiles / 2);
                VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_PROGRESS, pbytes, &prog_v));
                VRQ_CUDA(cudaMemsetAsync(prog_v, 0, pbytes, st));
                sp.progress = (int*)prog_v;
                sp.lock_window = lock_window;
            }
            {
                vrq_timer_scope td(ctx, VRQ_CAT_SCAN_DENSE, st);  // the dominant launch on its own (bench.py's roofline)
                VRQ_TRY(launch_pass(ctx, tma, tmap, tmap_mma, sp, m_pl, st));
            }
            sp.progress = nullptr;  // the gated fallback pass runs unthrottled (its counters would have to be reset)
            VRQ_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
            const int64_t avail = n - already > 0 ? n - already : 0;  // rows not returned by earlier chunks
            const int need = (int)((int64_t)k < avail ? (int64_t)k : avail);
            verify_counts_kernel<<<(nq + 127) / 128, 128, 0, st>>>(counts, m_pl.strips(), nq, need, flag, tau);
            vrq_count_launch(ctx);
            VRQ_TRY(launch_merge(ctx, lists, counts, m_pl.strips(), nq, cap, k, keys_out, nullptr, st));
            // 3. exact fallback, a no-op unless the verification raised the flag: the same dense pass again, with the
            //    thresholds of the queries that came up short set to infinity (the others collect the same lists again)
            sp.guard = flag;
            VRQ_TRY(launch_pass(ctx, tma, tmap, tmap_mma, sp, m_pl, st));
            VRQ_TRY(launch_merge(ctx, lists, counts, m_pl.strips(), nq, cap, k, keys_out, nullptr, st, flag));
            return 0;
        }
    }

    // prefix pass: an exact top-k of the first m rows seeds the thresholds of the main pass
    PassPlan main_pl;
    VRQ_TRY(plan_pass(ctx, tma, mma, code_bytes, n, nq, k, &main_pl));
    int64_t m = 0;
    if (key_lo == nullptr && n >= (int64_t)64 * main_pl.tile_rows() * main_pl.strips() && n >= (int64_t)16 * k) {
        m = main_pl.rows_per_strip();  // about one strip's worth of rows
        if (m < 4 * (int64_t)k) m = ((4 * (int64_t)k + main_pl.tile_rows() - 1) / main_pl.tile_rows()) * main_pl.tile_rows();
        if (m > n / 2) m = 0;
    }
    PassPlan pre_pl;
    if (m > 0) {
        VRQ_TRY(plan_pass(ctx, tma, mma, code_bytes, m, nq, k, &pre_pl));
        VRQ_TRY(plan_pass(ctx, tma, mma, code_bytes, n - m, nq, k, &main_pl));
    }
    const int cap = main_pl.cap() > (m > 0 ? pre_pl.cap() : 0) ? main_pl.cap() : pre_pl.cap();
    main_pl.set_cap(tma, code_bytes, cap);
    if (m > 0) pre_pl.set_cap(tma, code_bytes, cap);
    const int max_strips = (m > 0 && pre_pl.strips() > main_pl.strips() + 1) ? pre_pl.strips() : main_pl.strips() + 1;

    void *lists_v, *counts_v, *tau_v;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_LISTS, sizeof(uint64_t) * (size_t)max_strips * nq * cap, &lists_v));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_COUNTS, sizeof(int) * (size_t)max_strips * nq, &counts_v));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_TAU, sizeof(int) * (size_t)nq, &tau_v));
    uint64_t* lists = (uint64_t*)lists_v;
    int* counts = (int*)counts_v;
    int* tau = (int*)tau_v;

    ScanParams sp{};
    sp.codes = codes;
    sp.code_bytes = code_bytes;
    sp.pos_base = pos_base;
    sp.queries = q_dev;
    sp.nq = nq;
    sp.k = k;
    sp.lists = lists;
    sp.counts = counts;
    sp.dbg = dbg;
    sp.dbg_stride = n;
    sp.one = 1;
    sp.key_lo = key_lo;

    vrq_timer_scope ts(ctx, VRQ_CAT_SCAN, st);
    int extra = 0;
    if (m > 0) {
        sp.row_begin = 0;
        sp.row_end = m;
        sp.tau0 = nullptr;
        VRQ_TRY(launch_pass(ctx, tma, tmap, tmap_mma, sp, pre_pl, st));
        // the prefix top-k is parked in keys_out, then appended to the main pass's lists as one more "strip"
        VRQ_TRY(launch_merge(ctx, lists, counts, pre_pl.strips(), nq, cap, k, keys_out, tau, st));
        extra = 1;
    }
    sp.row_begin = m;
    sp.row_end = n;
    sp.tau0 = m > 0 ? tau : nullptr;
    if (n - m > 0) {
        VRQ_TRY(launch_pass(ctx, tma, tmap, tmap_mma, sp, main_pl, st));
    } else {
        VRQ_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)main_pl.strips() * nq, st));
    }
    if (extra) {
        uint64_t* slot = lists + (size_t)main_pl.strips() * nq * cap;
        VRQ_CUDA(cudaMemcpy2DAsync(slot, sizeof(uint64_t) * cap, keys_out, sizeof(uint64_t) * k, sizeof(uint64_t) * k, nq,
                                   cudaMemcpyDeviceToDevice, st));
        // count = number of real keys = min(k, m) == k here (m >= 4k)
        fill_int_kernel<<<(nq + 255) / 256, 256, 0, st>>>(counts + (size_t)main_pl.strips() * nq, nq, k);
        vrq_count_launch(ctx);
    }
    VRQ_TRY(launch_merge(ctx, lists, counts, main_pl.strips() + extra, nq, cap, k, keys_out, nullptr, st));
    return 0;
}

__global__ void last_keys_kernel(const uint64_t* __restrict__ keys, int64_t nq, int64_t stride, int col, unsigned long long* __restrict__ lo) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) lo[q] = keys[q * stride + col];
}

int vrq_hamming_topk_dev(vrq_ctx* ctx, const uint8_t* codes, int64_t n, int code_bytes, int64_t pos_base,
                         const uint8_t* q_dev, int64_t nq, int k, uint64_t* keys_out, cudaStream_t st, int32_t* dbg) {
    if (nq == 0) return 0;
    if (k <= 0 || k > VRQ_MAX_K) {
        vrq_set_error("Hamming top-k supports 1 <= k <= %d (got %d)", VRQ_MAX_K, k);
        return k <= 0 ? VRQ_ERR_ARG : VRQ_ERR_UNSUPPORTED;
    }
    if (code_bytes <= 0) {
        vrq_set_error("code size must be positive (got %d)", code_bytes);
        return VRQ_ERR_ARG;
    }
    if (n >= (1ll << 31)) {
        vrq_set_error("one shard holds at most 2^31 - 1 codes (TMA row coordinate); shard the database across GPUs");
        return VRQ_ERR_UNSUPPORTED;
    }
    if (pos_base < 0 || pos_base + n >= (1ll << VRQ_KEY_POS_BITS)) {
        vrq_set_error("positions beyond 2^40 are not supported");
        return VRQ_ERR_UNSUPPORTED;
    }
    VRQ_CUDA(cudaSetDevice(ctx->device));
    const int64_t QB = 1024;
    if (k <= VRQ_PASS_K) {
        for (int64_t q0 = 0; q0 < nq; q0 += QB) {
            const int nb = (int)(nq - q0 < QB ? nq - q0 : QB);
            VRQ_TRY(topk_batch(ctx, codes, n, code_bytes, pos_base, q_dev + q0 * code_bytes, nb, k, keys_out + q0 * k,
                               dbg ? dbg + q0 * n : nullptr, st));
        }
        return 0;
    }
    // k beyond what one pass holds per (strip, query) list: the ranking is produced in chunks of VRQ_PASS_K keys, each chunk
    // one more exact scan that only accepts keys above the last key of the chunk before it.  Keys are unique, so the
    // concatenation is exactly the top-k (faiss has no limit on k; CohereEnhancedVectorDB.py:267 with k=1000 asks for 10000).
    if (dbg) {
        vrq_set_error("the distance dump supports k <= %d", VRQ_PASS_K);
        return VRQ_ERR_UNSUPPORTED;
    }
    void *chunk_v, *lo_v;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_CHUNK_KEYS, sizeof(uint64_t) * (size_t)QB * VRQ_PASS_K, &chunk_v));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_CHUNK_LO, sizeof(uint64_t) * (size_t)QB, &lo_v));
    for (int64_t q0 = 0; q0 < nq; q0 += QB) {
        const int nb = (int)(nq - q0 < QB ? nq - q0 : QB);
        for (int done = 0; done < k; done += VRQ_PASS_K) {
            const int kk = k - done < VRQ_PASS_K ? k - done : VRQ_PASS_K;
            VRQ_TRY(topk_batch(ctx, codes, n, code_bytes, pos_base, q_dev + q0 * code_bytes, nb, kk, (uint64_t*)chunk_v, nullptr, st,
                               done ? (const unsigned long long*)lo_v : nullptr, done));
            VRQ_CUDA(cudaMemcpy2DAsync(keys_out + q0 * k + done, sizeof(uint64_t) * (size_t)k, chunk_v, sizeof(uint64_t) * (size_t)kk,
                                       sizeof(uint64_t) * (size_t)kk, (size_t)nb, cudaMemcpyDeviceToDevice, st));
            if (done + kk < k) {
                last_keys_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((const uint64_t*)chunk_v, nb, kk, kk - 1, (unsigned long long*)lo_v);
                vrq_count_launch(ctx);
                VRQ_CUDA(cudaGetLastError());
            }
        }
    }
    return 0;
}
