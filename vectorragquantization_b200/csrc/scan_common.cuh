// Pieces shared by the two Phase-I scan kernels (scan.cu: XOR + POPC on the integer pipes; scan_mma.cu: tcgen05 int8
// tensor-core contraction): launch parameters, PTX wrappers, list compaction and the TMA descriptor of the code array.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "topk_utils.cuh"
#include "vrq_internal.cuh"

namespace vrq {

constexpr int CODE_BYTES = 128;
constexpr int TMA_BOX_ROWS = 256;  // rows per TMA box (hardware limit per box dimension)
constexpr int BAR_CONSUMERS = 1;
constexpr int TAU_INF = 1 << 28;   // "no threshold yet": above any distance
// CW = consumer warps per CTA: 8 for query batches (long inner loops keep the POPC pipe busy), 16 for <= 8 queries
// per pass (short per-tile work: the second warp group hides the first one's shared-memory / barrier latency).
template <int CW>
struct ScanCfg {
    static constexpr int CONSUMER_THREADS = CW * 32;
    static constexpr int THREADS = CONSUMER_THREADS + 32;
    static constexpr int TILE_ROWS = CW * 32;
    static constexpr int STAGE_BYTES = TILE_ROWS * CODE_BYTES;
};

struct ScanParams {
    const uint8_t* codes;  // local row 0
    int code_bytes;
    int64_t row_begin, row_end;  // local rows scanned by this launch
    int64_t pos_base;            // global position of local row 0
    const uint8_t* queries;      // [nq][code_bytes]
    int nq, k;
    int num_strips;
    int64_t rows_per_strip;  // multiple of TILE_ROWS
    int qtile;               // queries per CTA
    int cap;                 // capacity of one list
    int group_tiles;         // tiles between overflow checks
    int stages;
    uint64_t* lists;  // [list_strips][nq][cap]
    int* counts;      // [list_strips][nq]
    const int* tau0;  // [nq] or null
    const int* guard;  // nullable: the launch does nothing when *guard == 0 (fallback passes enqueued ahead of knowing they are needed)
    int sample_mode;   // tensor-core kernel: only the k-th DISTANCE per query matters (threshold-only pass): lists are compacted on
                       // the distance bits alone and leave the kernel uncompacted (<= compact_limit + group_tiles * 128 keys)
    int compact_limit; // tensor-core kernel: compact a list once it holds more than this many keys (0 = cap - group_tiles * 128)
    int tau_bias;      // added to tau0 (1 turns a k'-th distance T of a sample into the inclusive bound "d <= T")
    // tensor-core kernel: tile i of the launch starts at row_begin + (i >> run_shift) * run_stride + (i & run_mask) * 128:
    // dense scan = runs of 1 tile, stride 128; strided sample = runs of 2^run_shift consecutive tiles, run_stride rows apart
    int64_t run_stride;
    int run_shift;
    int64_t total_tiles;
    int32_t* dbg;     // tests only: every (query, row) Hamming distance of the launch, [nq][dbg_stride] (tensor-core kernel)
    int64_t dbg_stride;
    const unsigned long long* key_lo;  // [nq] or null: only keys STRICTLY GREATER than key_lo[q] are candidates (the next chunk of a
                                       // top-k larger than one pass can hold: vrq_hamming_topk_dev walks the ranking in chunks)
    unsigned short* sample_out;  // tensor-core 128-query-tile kernel, list-free sample pass: [strip][nq][2][4] smallest distances
                                 // seen by each epilogue thread (0xFFFF = none); no lists are written (scan_mma.cu)
    int seg_cols, seg_full, seg_tail;  // tensor-core pair kernel, 1-D grid of clusters: query-tile pairs, full strips per pair column,
                                       // clusters that share the tail strip (scan_mma.cu); seg_cols == 0: classic (query tile, strip) grid
    int* progress;    // tensor-core pair kernel, dense pass: [strip][query-tile pair] tile counters (lockstep throttle), or null
    int lock_window;  // a pair's TMA producer stays within this many tiles of the slowest pair of its strip
    int one;          // == 1, opaque to the compiler: multiplier that keeps the popcount accumulation on the FMA pipe (IMAD)
};

// ---- PTX helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// Same wait for roles that run AHEAD of their consumer (producers waiting for a free slot): between polls the thread
// sleeps, so the polling does not compete with the data path for issue slots and shared-memory cycles.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    while (!done) {
        __nanosleep(sleep_ns);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr) {
    int r;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// ---- list compaction: keep the k smallest keys of one (strip, query) list --------------------------------
template <int CONSUMER_THREADS>
__device__ void compact_list(uint64_t* glist, int n, int k, unsigned long long* scratch, SelectScratch* sc, int tid,
                             int* cnt_q, int* tau_q) {
    for (int i = tid; i < n; i += CONSUMER_THREADS) scratch[i] = glist[i];
    group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
    const unsigned long long kth =
        radix_select_kth<CONSUMER_THREADS>([&](int i) { return scratch[i]; }, n, k, tid, sc, BAR_CONSUMERS);
    if (tid == 0) sc->counter = 0;
    group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
    for (int i = tid; i < n; i += CONSUMER_THREADS) {
        unsigned long long key = scratch[i];
        if (key <= kth) glist[atomicAdd(&sc->counter, 1)] = key;
    }
    group_sync<CONSUMER_THREADS>(BAR_CONSUMERS);
    if (tid == 0) {
        *cnt_q = k;
        *tau_q = (int)(kth >> VRQ_KEY_POS_BITS);
    }
}

// One WARP keeps the k smallest keys of a list, in place in global memory (the tensor-core kernel compacts 8 lists at a
// time, one per epilogue warp; a block-wide select per list serialises 128 lists behind ~15 barriers each).
// bar_id: a named barrier private to the calling warp.  Keys are unique, so "key <= k-th smallest" keeps exactly k.
// ties_ok > 0 (threshold-only passes): select on the distance bits alone (2 radix passes instead of 7) and keep every key
// whose distance ties with the k-th one - unless that would leave more than ties_ok keys, then fall back to the exact k.
__device__ __forceinline__ void compact_list_warp(uint64_t* glist, int n, int k, SelectScratch* sc, int lane, int bar_id, int* cnt_q,
                                                  int* tau_q, int ties_ok = 0) {
    unsigned long long kth = 0;
    bool exact = ties_ok <= 0;
    if (!exact) {
        kth = radix_select_kth<32>([&](int i) { return (unsigned long long)glist[i]; }, n, k, lane, sc, bar_id, VRQ_KEY_POS_BITS);
        int kept = 0;
        for (int base = 0; base < n; base += 128) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + 32 * u + lane;
                kept += (i < n && (unsigned long long)glist[i] <= kth) ? 1 : 0;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
        exact = kept > ties_ok;
    }
    if (exact) kth = radix_select_kth<32>([&](int i) { return (unsigned long long)glist[i]; }, n, k, lane, sc, bar_id);
    // in-place compaction, 128 keys per step (4 independent loads per lane): a step writes below the end of the chunk it
    // has just read
    int out = 0;
    for (int base = 0; base < n; base += 128) {
        unsigned long long key[4];
        bool keep[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + 32 * u + lane;
            key[u] = i < n ? (unsigned long long)glist[i] : ~0ull;
            keep[u] = i < n && key[u] <= kth;
        }
        __syncwarp();  // the whole 128-key chunk is in registers before anything is written below its end
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned m = __ballot_sync(0xffffffffu, keep[u]);
            if (keep[u]) glist[out + __popc(m & ((1u << lane) - 1u))] = key[u];
            out += __popc(m);
        }
        __syncwarp();
    }
    if (lane == 0) {
        *cnt_q = out;
        *tau_q = (int)(kth >> VRQ_KEY_POS_BITS);
    }
    __syncwarp();
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_codes_tmap(const uint8_t* codes, int64_t nrows, int box_rows, CUtensorMap* out) {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !f) {
            vrq_set_error("cuTensorMapEncodeTiled is not available from the driver");
            return VRQ_ERR_UNSUPPORTED;
        }
        fn = (PFN_tmapEncodeTiled)f;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)CODE_BYTES, (cuuint64_t)nrows};
    cuuint64_t gstride[1] = {(cuuint64_t)CODE_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)CODE_BYTES, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)codes, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vrq_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return VRQ_ERR_UNSUPPORTED;
    }
    return 0;
}

// ---- scan_mma.cu: tensor-core variant of the scan kernel (same lists / counts / thresholds contract) ----------
struct MmaPlan {
    int qtiles, strips, group_tiles, cap, raw_stages, b_stages;
    int64_t rows_per_strip;
    size_t smem, smem_limit;
    bool f4;    // packed e2m1 operands (kind::mxf4) instead of int8
    bool mid;   // 33 .. 64 queries on the swapped-operand kernel (four accumulator column groups instead of two)
    bool w128;  // 65 .. 128 queries on the swapped-operand kernel (one A buffer handed over per K half, 8 column groups)
    bool few;   // <= 128 queries: swapped-operand kernel (database rows = M, expanded straight into tensor memory, bias column)
    bool pair;  // CTA pairs (tcgen05 cta_group::2): two query tiles share every tile of database rows
    int seg_cols, seg_full, seg_tail;  // pair scheduler (see ScanParams); seg_cols == 0: classic grid
};
constexpr int MMA_TILE_ROWS = 128;
int plan_scan_mma(vrq_ctx* ctx, int64_t rows, int nq, int k, MmaPlan* pl, bool allow_few = true);
void mma_plan_set_cap(MmaPlan* pl, int cap);
int launch_scan_mma(vrq_ctx* ctx, const CUtensorMap& tmap128, const CUtensorMap& tmap64, const ScanParams& sp, const MmaPlan& pl,
                    cudaStream_t st);

inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}


}  // namespace vrq
