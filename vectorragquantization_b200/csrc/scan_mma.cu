// Phase I for query batches on the 5th-generation tensor cores (tcgen05, sm_100a): the same exact Hamming top-k as
// scan.cu (faiss IndexBinaryFlat::search, CohereEnhancedVectorDB.py:268 / VectorDBInt8.py:218), with the
// (query, code) bit contraction done by tcgen05.mma instead of XOR + POPC on the integer pipes.
//
// Arithmetic (exact):  with database bits c_k in {0, 1} and query bits mapped to s_k in {+1, -1}
//     dot(q, c) = sum_k s_k c_k = #(c=1, q=1) - #(c=1, q=0)      and      hamming(q, c) = popc(q) - dot(q, c)
// so "hamming < tau" is "dot > popc(q) - tau", one compare per accumulator element.  Two operand kinds: packed e2m1
// under kind::mxf4 with unit block scales (default; f32 accumulation of +-1 products, |sum| <= 1024: exact) and int8
// under kind::i8 (s32 accumulation).
//
// Kernel (one CTA per SM, grid = query tiles x row strips; a CTA walks its strip in ascending row order; with the e2m1
// kind two neighbouring query tiles form a CTA pair - cta_group::2 - and share every tile of database rows):
//   * A operand = the CTA's 128 queries, written ONCE into tensor memory with tcgen05.st and read from there by every
//     MMA (the ".ts" form): shared-memory bandwidth is left to the B operand.
//   * B operand = 128 database rows per tile, never materialised in HBM: a producer warp TMA-loads the raw 128-byte codes
//     (the CTA's half of the tile when paired), eight expander warps turn each K-block of code bits into one 128-byte
//     operand row (w & mask planes - the bit -> element permutation inside a K-block is arbitrary as long as the query
//     side uses the same one) and store it with the 128-byte swizzle the UMMA shared-memory descriptor expects, into an
//     8-stage ring of K-block stages.
//   * Two issuer warps (one elected lane each; the leader CTA of a pair) issue 4 MMAs per K-block into one of two
//     128-column TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator (multicast to both CTAs).
//   * Eight epilogue warps (lane = query, columns = database rows).  In the dense pass they run at 128 registers (setmaxnreg;
//     everyone else at 72): all four tcgen05.ld.x16 of a tile at once, the accumulator handed back before anything is
//     examined, then a tree of 3-input maxima and ONE compare of the largest of the 64 dots with the query's threshold; only
//     a group that holds a survivor is looked at column by column and appended to the (strip, query) list.  The loop is ~70
//     instructions per tile when nothing survives - measured, the epilogue's instruction stream (not tensor memory, not
//     shared memory) was what kept the tensor pipe at 83 % (now 99.5 % under ncu).  The other instantiations (sample pass,
//     int8 kind, tests' distance dump) keep a generic loop over two register sets.  Lists, thresholds and the merge tree
//     are those of scan.cu; overflowing lists are compacted in place by one warp per list.
// Few queries per pass (3 .. 96) take hamming_scan_mma_wide_kernel further down: operand roles swapped (database rows = M,
// expanded straight into tensor memory), a bias column that gives all query columns one threshold.
// Host side (scan.cu: topk_batch): thresholds come from a strided sample pass of this kernel, one dense pass collects the
// candidates, a verification kernel un-gates an exact fallback pass if a query came up short.  DESIGN.md section 3.2.1.
#include "scan_common.cuh"

namespace vrq {
namespace {

constexpr int MQ = 128;      // queries per CTA  (MMA M)
constexpr int MROWS = 128;   // database rows per tile (MMA N)
// Operand kinds.  I8: int8 operands, 8 K-blocks of 128 elements per row (16 code bytes -> 128 bytes).  F4: packed e2m1
// (4-bit float: +1 = 0x2, -1 = 0xA, 0 = 0x0) under kind::mxf4 with every block scale = 1.0, 4 K-blocks of 256 elements
// per row (32 code bytes -> 128 bytes), twice the tensor rate and half the shared-memory bytes per row.  Either way one
// MMA consumes 32 bytes of K per operand row (8 TMEM columns of A, a 32-byte step of the B descriptor).
constexpr int KIND_I8 = 8, KIND_F4 = 4;
template <int KIND>
struct KindCfg {
    static constexpr int KBLOCKS = KIND == KIND_I8 ? 8 : 4;  // shared-memory stages per 128-row tile
};
constexpr int STAGE_BYTES_B = MROWS * 128;
constexpr int STAGE_BYTES_RAW = MROWS * CODE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EXP_WARPS = 8;
constexpr int WARP_MMA = 8, MMA_WARPS = 2, WARP_TMA = 10, WARP_EXP0 = 11;  // issuer warp 8 takes even tiles, warp 9 odd tiles
constexpr int MMA_KERNEL_THREADS = (WARP_EXP0 + EXP_WARPS) * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr uint32_t TMEM_COLS = 512, TMEM_A_COL = 0, TMEM_D_COL = 256;
constexpr uint32_t TMEM_SFA_COL = 128, TMEM_SFB_COL = 160, TMEM_SF_COLS = 32;  // F4 only: block scales, all 1.0 (UE8M0 0x7F)
constexpr int B_STAGES = 8, MAX_RAW_STAGES = 4;
constexpr int BAR_WARP0 = 2;  // named barriers 2..9: one per epilogue warp (list compaction)
constexpr int SAMPLE_KEEP = 4;  // distances each epilogue thread keeps in the list-free sample pass

// instruction descriptor (bit layout of cute::UMMA::InstrDescriptor): D = s32 (bits 4-5 = 2), A = B = signed 8 bit (bits
// 7-9, 10-12 = 1), both K-major, N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MROWS >> 3) << 17) | ((uint32_t)(MQ >> 4) << 24);
// block-scaled descriptor (cute::UMMA::InstrDescriptorBlockScaled): A = B = e2m1 (MXF4 format 1), K-major, scale format
// UE8M0 (bit 23), scale-factor ids 0, N = 128, M = 128, K = 64 (bit 31 = 0); D is always f32
constexpr uint32_t IDESC_F4 = (1u << 7) | (1u << 10) | ((uint32_t)(MROWS >> 3) << 17) | (1u << 23) | ((uint32_t)(MQ >> 4) << 24);
// the same for a CTA pair: M = 256 (128 queries in each CTA's TMEM), N = 128 (64 database rows in each CTA's shared memory)
constexpr uint32_t IDESC_F4_PAIR = (1u << 7) | (1u << 10) | ((uint32_t)(MROWS >> 3) << 17) | (1u << 23) | ((uint32_t)(256 >> 4) << 24);

// ---- tcgen05 wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
        : "memory");
}
// ---- CTA-pair (cta_group::2) helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(rank));
    return ra;
}
// Remote arrive with the default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id) does.  What the
// arrive publishes here is never read by the other CTA's THREADS: B stages are read by the tensor core through the async
// proxy (made visible by fence.proxy.async before the arrive) and accumulator hand-backs carry no memory at all.  The
// .release.cluster form costs MEMBAR.ALL.GPU + ERRBAR per arrive (measured: the pair kernel ran 35 % slower with it).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_f4_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// atomicAdd on a shared-memory counter through its 32-bit shared address: ATOMS.  (atomicAdd(&sm->cnt_s[q], ...) on the
// struct carved out of the dynamic shared-memory block compiles to a GENERIC atomic, ATOM.E.ADD.STRONG.GPU.)
__device__ __forceinline__ int atoms_add(uint32_t saddr, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
// named barrier over the epilogue threads that also ORs a predicate
__device__ __forceinline__ int epi_sync_or(int pred) {
    int out;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.s32 q, %1, 0;\n"
        "bar.red.or.pred p, %2, %3, q;\n"
        "selp.s32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(out)
        : "r"(pred), "n"(BAR_CONSUMERS), "n"(EPI_THREADS)
        : "memory");
    return out;
}

struct MmaSmem {
    unsigned long long raw_full[MAX_RAW_STAGES], raw_empty[MAX_RAW_STAGES];
    unsigned long long b_full[B_STAGES], b_empty[B_STAGES];
    unsigned long long acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    int tau_s[MQ];
    int cnt_s[MQ];
    SelectScratch sc[EPI_WARPS];  // one radix-select scratch per epilogue warp
};

// CG = 1: one CTA per 128-query tile.  CG = 2 (e2m1 only): a CTA pair shares every tile of database rows - each CTA
// holds its own 128 queries in TMEM, loads and expands HALF of the tile's rows (64), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 128) which reads both halves; per SM the shared-memory traffic per tile halves.
// SAMP = true: the list-free sample pass (thresholds only) - a separate instantiation, so that none of its code sits in the
// dense pass's epilogue.
// VAR = 0: generic epilogue loop at 96 registers (every kind, the sample pass, strided runs, the tests' distance dump).
// VAR = 4: the dense pass of the CTA-pair kernel.  The epilogue warps run at 128 registers and everyone else at 72
// (setmaxnreg; the block carries one idle pad warp so that every warpgroup is whole), which lets them load all four
// accumulator column groups at once and hand the accumulator back before anything is examined; see the lean loop below.
template <int KIND, int CG, bool SAMP, int VAR = 0>
__global__ void __launch_bounds__(VAR == 4 ? MMA_KERNEL_THREADS + 32 : MMA_KERNEL_THREADS, 1)
hamming_scan_mma_kernel(const __grid_constant__ CUtensorMap tmap, ScanParams p, int raw_stages) {
    constexpr bool F4 = KIND == KIND_F4;
    constexpr int KBLOCKS = KindCfg<KIND>::KBLOCKS;
    static_assert(CG == 1 || F4, "CTA pairs are built for the e2m1 kind only");
    constexpr int MY_ROWS = MROWS / CG;                    // database rows this CTA loads and expands per tile
    constexpr int STAGE_BYTES_RAW = MY_ROWS * CODE_BYTES;  // shadows the namespace constant: per-CTA bytes
    constexpr int STAGE_BYTES_B = MY_ROWS * 128;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    // The B ring has exactly 8 stages of one K-block (16 KB) each, so a tile's stages are compile-time offsets from one
    // per-tile base: I8 tiles use all 8 (stage = K-block, parity = tile & 1); F4 tiles alternate between stages 0-3 and 4-7.
    constexpr int GROUPS = B_STAGES / KBLOCKS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* raw_mem = base;
    uint8_t* b_mem = raw_mem + (size_t)raw_stages * STAGE_BYTES_RAW;
    MmaSmem* sm = (MmaSmem*)(b_mem + (size_t)B_STAGES * STAGE_BYTES_B);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // A CTA (pair) works through one or more SEGMENTS = (query tile, range of tiles, list strip).  Classic 2-D grid
    // (seg_cols == 0): one segment, blockIdx = (query tile, strip).  Pair scheduler (seg_cols > 0, 1-D grid of clusters):
    // cluster c < seg_cols * seg_full owns strip c / seg_cols of query-tile pair c % seg_cols - the pairs of one strip walk the
    // same rows at the same time, so the re-reads hit L2; the remaining seg_tail clusters share the TAIL strip (the rows
    // after the seg_full full strips), each doing it for seg_cols / seg_tail query-tile pairs one after the other.  With
    // 148 SMs and 8 query tiles: 72 pairs x 1/18.5 of the rows + 2 pairs x (2 x 1/37): every SM is busy.
    const int tiles_per_strip = (int)(p.rows_per_strip / MROWS);
    int q0 = 0, qt = 0, strip = 0, ntiles = 0;
    int64_t tile0 = 0;
    auto segment = [&](int sgi) -> bool {
        int col;
        if (p.seg_cols == 0) {
            if (sgi > 0) return false;
            col = blockIdx.x;
            strip = blockIdx.y;
            tile0 = (int64_t)strip * tiles_per_strip;
            ntiles = (int)max((int64_t)0, min((int64_t)tiles_per_strip, p.total_tiles - tile0));
            q0 = col * MQ;
        } else {
            const int c = (int)(blockIdx.x / CG), nfull = p.seg_cols * p.seg_full;
            if (c < nfull) {
                if (sgi > 0) return false;
                col = c % p.seg_cols;
                strip = c / p.seg_cols;
                tile0 = (int64_t)strip * tiles_per_strip;
                ntiles = (int)max((int64_t)0, min((int64_t)tiles_per_strip, p.total_tiles - tile0));
            } else {
                const int per = p.seg_cols / p.seg_tail;
                if (sgi >= per) return false;
                col = (c - nfull) * per + sgi;
                strip = p.seg_full;
                tile0 = (int64_t)p.seg_full * tiles_per_strip;
                ntiles = (int)max((int64_t)0, p.total_tiles - tile0);
            }
            q0 = (col * CG + (int)rank) * MQ;
        }
        qt = max(0, min(MQ, p.nq - q0));
        return true;
    };
    const int64_t run_mask = ((int64_t)1 << p.run_shift) - 1;
    auto tile_row = [&](int t) -> int64_t {
        const int64_t i = tile0 + t;
        return p.row_begin + (i >> p.run_shift) * p.run_stride + (i & run_mask) * MROWS;
    };
    const int64_t s_end = p.row_end;
    if (p.guard && *p.guard == 0) return;

    if (tid == 0) {
        for (int s = 0; s < raw_stages; s++) {
            mbar_init(smem_u32(&sm->raw_full[s]), 1);
            mbar_init(smem_u32(&sm->raw_empty[s]), EXP_WARPS);
        }
        for (int s = 0; s < B_STAGES; s++) {
            mbar_init(smem_u32(&sm->b_full[s]), EXP_WARPS / 2);
            mbar_init(smem_u32(&sm->b_empty[s]), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(smem_u32(&sm->acc_full[s]), 1);
            mbar_init(smem_u32(&sm->acc_empty[s]), EPI_WARPS * CG);  // pair: both CTAs' epilogues release the leader's MMA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers exist before anyone arrives on them remotely
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    // running state of the rings across segments (a CTA's tiles are numbered tbase + t over all its segments: stages,
    // accumulators and barrier phases continue where the previous segment left them)
    uint32_t prod_s = 0, prod_ph = 0;  // TMA producer: raw ring
    uint32_t rs = 0, rph = 0;          // expanders: raw ring
    int tbase = 0;
    // The segment loop is written once per ROLE CLASS (epilogue warps / everyone else) so that VAR = 4 can run the two
    // classes under different register budgets: ptxas applies a setmaxnreg to the code it dominates.
    auto seg_begin_sync = [&]() {
        tc_fence_before();
        asm volatile("bar.sync 0;" ::: "memory");
        tc_fence_after();
    };
    auto seg_end_sync = [&]() {
        // end of the segment: every role is done with its tiles (the last accumulators were read, so every MMA has completed)
        // before the next segment rewrites the query operand in tensor memory - in both CTAs of a pair
        tc_fence_before();
        asm volatile("bar.sync 0;" ::: "memory");
        if constexpr (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and TMEM until the very end
        tc_fence_after();
    };
    auto epilogue_role = [&](const int sgi) {
    // ---- the query tile becomes the A operand in tensor memory (lane = query); warps 0-3 write it, every epilogue
    //      thread keeps popc(query) of the query it filters
    int pcq = 0;
    {
        const int q = tid & (MQ - 1);
        const bool qvalid = q < qt;
        const bool writer = warp < 4;
        const uint32_t* qrow = reinterpret_cast<const uint32_t*>(p.queries + (size_t)(q0 + (qvalid ? q : 0)) * CODE_BYTES);
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + TMEM_A_COL;
#pragma unroll 2
        for (int W = 0; W < 32; W += 2) {
            const uint32_t w = qvalid ? __ldg(qrow + W) : 0u;
            const uint32_t w2 = qvalid ? __ldg(qrow + W + 1) : 0u;
            pcq += __popc(w) + __popc(w2);
            if (writer) {
                uint32_t v[8];
                if constexpr (!F4) {
                    // columns 8 W + t: byte b = bit (t + 8 b) of word W as +1 (0x01) / -1 (0xFF)
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const uint32_t m = (w >> t) & 0x01010101u;
                        v[t] = qvalid ? (m | ((m ^ 0x01010101u) * 0xFFu)) : 0u;
                    }
                    tmem_st8(lane_base + 8 * W, v);
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const uint32_t m = (w2 >> t) & 0x01010101u;
                        v[t] = qvalid ? (m | ((m ^ 0x01010101u) * 0xFFu)) : 0u;
                    }
                    tmem_st8(lane_base + 8 * W + 8, v);
                } else {
                    // columns 4 W + t: nibble j = bit (t + 4 j) of word W as e2m1 +-(1 / b_t), b_t = the value the expanders
                    // give a set database bit of plane t (0.5, 1, 2, 2): +-2.0 = 0x4/0xC, +-1.0 = 0x2/0xA, +-0.5 = 0x1/0x9
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t mag = t == 0 ? 0x44444444u : (t == 1 ? 0x22222222u : 0x11111111u);
                        v[t] = qvalid ? ((mag | 0x88888888u) ^ (((w >> t) & 0x11111111u) << 3)) : 0u;
                        v[4 + t] = qvalid ? ((mag | 0x88888888u) ^ (((w2 >> t) & 0x11111111u) << 3)) : 0u;
                    }
                    tmem_st8(lane_base + 4 * W, v);
                }
            }
        }
        if (writer) {
            if (F4 && sgi == 0) {
                // every block scale (UE8M0) = 0x7F = 2^0: with one constant the scale-factor layout does not matter
                uint32_t one[8];
#pragma unroll
                for (int t = 0; t < 8; t++) one[t] = 0x7F7F7F7Fu;
#pragma unroll
                for (int c = 0; c < (int)(2 * TMEM_SF_COLS); c += 8) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + TMEM_SFA_COL + c, one);
            }
            tmem_wait_st();
            sm->tau_s[q] = qvalid ? (p.tau0 ? min(p.tau0[q0 + q], TAU_INF - 1) + p.tau_bias : TAU_INF) : 0;
            sm->cnt_s[q] = 0;
        }
    }
    seg_begin_sync();
    {
        // 8 warps: warp & 3 = TMEM lane quadrant (32 queries), warp >> 2 = which 64 of the tile's 128 columns.  The two
        // threads that share a query append to the same list through a shared-memory counter.
        const int q = tid & (MQ - 1);
        const int half = warp >> 2;
        const bool qvalid = q < qt;
        uint64_t* my_list = p.lists + ((size_t)strip * p.nq + q0 + (qvalid ? q : 0)) * p.cap;
        const bool has_lo = p.key_lo != nullptr;
        const unsigned long long lo_q = (has_lo && qvalid) ? p.key_lo[q0 + q] : 0ull;
        int thr = qvalid ? pcq - sm->tau_s[q] : 0x7fffffff;  // survivor <=> dot > thr <=> hamming < tau
        float thr_f = (float)thr;
        constexpr bool samp = SAMP;  // list-free sample pass (thresholds only)
        int sbest[SAMPLE_KEEP];  // ascending
#pragma unroll
        for (int i = 0; i < SAMPLE_KEEP; i++) sbest[i] = 0x7fff;
        if (samp && qvalid) {
            thr = pcq - 0x7fff;
            thr_f = (float)thr;
        }
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + TMEM_D_COL + (uint32_t)(half * 64);
        const int limit = p.compact_limit > 0 ? min(p.compact_limit, p.cap - p.group_tiles * MROWS) : p.cap - p.group_tiles * MROWS;
        int until_check = p.group_tiles;
        auto release_acc = [&](const int as) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2)
                    mbar_arrive_cluster(mapa_rank(smem_u32(&sm->acc_empty[as]), 0));
                else
                    mbar_arrive(smem_u32(&sm->acc_empty[as]));
            }
        };
        const uint32_t cnt_addr = smem_u32(&sm->cnt_s[q]);
        // largest of 16 f32 accumulator columns: a depth-3 tree of 3-input maxima (FMNMX3) instead of a chain of eight
        auto max16 = [](const int(&w)[16]) -> float {
            auto f = [&](int j) -> float { return __int_as_float(w[j]); };
            auto mx3 = [](float x, float y, float z) -> float { return fmaxf(fmaxf(x, y), z); };
            const float a0 = mx3(f(0), f(1), f(2)), a1 = mx3(f(3), f(4), f(5)), a2 = mx3(f(6), f(7), f(8));
            const float a3 = mx3(f(9), f(10), f(11)), a4 = mx3(f(12), f(13), f(14));
            return fmaxf(mx3(a0, a1, a2), mx3(a3, a4, f(15)));
        };
        // the survivors of one group of 16 columns go to the (strip, query) list; m = the group's largest dot
        auto collect = [&](const int(&w)[16], const int g, const int nv, const int64_t lrow0, const int m) {
            auto dot_of = [&](int j) -> int { return F4 ? (int)__int_as_float(w[j]) : w[j]; };
            // Some column of this lane survives.  bit (15 - j) of mask <=> w[j] > thr: the sign of thr - w[j] is
            // shifted in with one funnel shift per column (2 instructions per column, no branches).
            const unsigned long long pos0 = (unsigned long long)(p.pos_base + lrow0 + 16 * g);
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t sgn = F4 ? __float_as_uint(thr_f - __int_as_float(w[j])) : (uint32_t)(thr - w[j]);
                mask = __funnelshift_l(sgn, mask, 1);
            }
            if (nv < 16) mask = nv <= 0 ? 0u : (mask & ~(0xFFFFu >> nv));
            if (has_lo) {
                // a later chunk of a large top-k: keys at or below the chunk's lower bound were returned already
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if ((mask >> (15 - j)) & 1u) {
                        const unsigned long long key = ((unsigned long long)(pcq - dot_of(j)) << VRQ_KEY_POS_BITS) | (pos0 + j);
                        if (key > lo_q) {
                            const int slot = atoms_add(cnt_addr, 1);
                            if (slot >= p.cap) __trap();
                            my_list[slot] = key;
                        }
                    }
                }
            } else if (__popc(mask) == 1 && nv >= 16) {
                // the usual case once tau has converged: the single survivor is the maximum itself
                const int slot = atoms_add(cnt_addr, 1);
                if (slot >= p.cap) __trap();  // cannot happen (overflow check every group_tiles tiles); never write past a list
                my_list[slot] = ((unsigned long long)(pcq - m) << VRQ_KEY_POS_BITS) | (pos0 + (__clz((int)mask) - 16));
            } else if (mask) {
                int slot = atoms_add(cnt_addr, __popc(mask));
                if (slot + __popc(mask) > p.cap) __trap();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if ((mask >> (15 - j)) & 1u) {
                        my_list[slot] = ((unsigned long long)(pcq - dot_of(j)) << VRQ_KEY_POS_BITS) | (pos0 + j);
                        slot++;
                    }
                }
            }
        };
        auto examine = [&](const int(&w)[16], const int g, const int nvalid, const int64_t lrow0) {
            const int nv = nvalid - 16 * g;  // valid columns in this group (>= 16: all)
            // F4 accumulates in f32: the dots are integers of magnitude <= 1024, exact in binary32
            auto dot_of = [&](int j) -> int { return F4 ? (int)__int_as_float(w[j]) : w[j]; };
            bool any;
            int m;
            if constexpr (F4) {
                float mf;
                mf = max16(w);
                any = mf > thr_f;
                m = (int)mf;
            } else {
                m = w[0];
#pragma unroll
                for (int j = 1; j < 16; j++) m = max(m, w[j]);
                any = m > thr;
            }
            if (samp) {
                // List-free sample pass: only (an upper bound of) the k'-th smallest DISTANCE of the sample matters.  Every
                // epilogue thread keeps the SAMPLE_KEEP smallest distances it has seen in registers (a branch-free sorted
                // insert, 7 min / max per element); its own threshold is the largest of them, so after the first tiles
                // almost no group gets here (the chance that row r of a thread is among its best so far is 4 / r).  The
                // k'-th smallest of the union over the threads of a query is >= the k'-th smallest of the whole sample and
                // equal to it unless one thread holds more than SAMPLE_KEEP of the k' best - a slightly looser threshold at
                // worst, never a wrong result (the dense pass is verified, scan.cu).
                if (any) {
                    const unsigned long long pos0 = (unsigned long long)(p.pos_base + lrow0 + 16 * g);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        int x = pcq - dot_of(j);
                        const bool ok = j < nv && !(has_lo && ((((unsigned long long)x) << VRQ_KEY_POS_BITS) | (pos0 + j)) <= lo_q);
                        x = ok ? x : 0x7fff;
#pragma unroll
                        for (int i = 0; i < SAMPLE_KEEP - 1; i++) {
                            const int lo_ = min(sbest[i], x);
                            x = max(sbest[i], x);
                            sbest[i] = lo_;
                        }
                        sbest[SAMPLE_KEEP - 1] = min(sbest[SAMPLE_KEEP - 1], x);
                    }
                    thr = pcq - sbest[SAMPLE_KEEP - 1];  // survivor <=> hamming < the largest distance this thread keeps
                    thr_f = (float)thr;
                }
            } else if (any) {
                collect(w, g, nv, lrow0, m);
            }
        };
        auto overflow_check = [&](const int t) {
        // ---- overflow check every group_tiles tiles: no list may exceed cap during the next group ----
        if (!samp && --until_check == 0 && t + 1 < ntiles) {
            until_check = p.group_tiles;
            group_sync<EPI_THREADS>(BAR_CONSUMERS);  // every append of this group of tiles is in its list
            if (epi_sync_or(sm->cnt_s[q] > limit)) {
                for (int qq = warp; qq < qt; qq += EPI_WARPS) {  // one list per warp, 8 lists at a time
                    const int n = sm->cnt_s[qq];
                    if (n > limit)
                        compact_list_warp(p.lists + ((size_t)strip * p.nq + q0 + qq) * p.cap, n, p.k, &sm->sc[warp], lane, BAR_WARP0 + warp,
                                          &sm->cnt_s[qq], &sm->tau_s[qq], p.sample_mode ? limit : 0);
                }
                group_sync<EPI_THREADS>(BAR_CONSUMERS);
                if (qvalid) thr = pcq - sm->tau_s[q];
                thr_f = (float)thr;
            }
        }
        };
        if constexpr (VAR == 4) {
            // Lean tile loop of the dense pass (tile t of the segment = rows row_begin + (tile0 + t) * 128 ...).  Measured with
            // the generic loop below: 151 instructions per tile and warp on the no-survivor path kept the epilogue warps busy
            // for ~950 of the 1024 cycles a tile's MMAs take, so every survivor made them late for the next accumulator
            // (tensor pipe 83 % busy; 99.95 % with an epilogue that only loads).  Here: barrier addresses and the number of
            // full tiles are loop invariants, all four column groups are loaded at once (64 registers - the epilogue warps
            // run at 128 registers, setmaxnreg), the accumulator goes back before anything is examined, and ONE test on the
            // maximum of all 64 columns decides whether any group needs a closer look.
            const int64_t row0 = p.row_begin + tile0 * MROWS + half * 64;  // first row of this warp's columns in tile 0
            const int64_t left0 = s_end - row0;
            const int full_tiles = left0 >= 64 ? (int)min((int64_t)ntiles, (left0 - 64) / MROWS + 1) : 0;
            uint32_t full_bar = smem_u32(&sm->acc_full[0]);
            uint32_t empty_bar = CG == 2 ? mapa_rank(smem_u32(&sm->acc_empty[0]), 0) : smem_u32(&sm->acc_empty[0]);
            uint32_t acc0 = lane_base;
            // opaque to the compiler, or it rematerialises the ~20 instructions of address arithmetic in every tile
            asm volatile("" : "+r"(full_bar), "+r"(empty_bar), "+r"(acc0));
            uint32_t T = (uint32_t)tbase;
            for (int t = 0; t < ntiles; t++, T++) {
                const uint32_t as = T & 1u;
                mbar_wait(full_bar + 8u * as, (T >> 1) & 1u);
                tc_fence_after();
                const uint32_t acc = acc0 + as * MROWS;
                int v[4][16];
                __syncwarp();
                tmem_ld16(acc, v[0]);
                tmem_ld16(acc + 16, v[1]);
                tmem_ld16(acc + 32, v[2]);
                tmem_ld16(acc + 48, v[3]);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2)
                        mbar_arrive_cluster(empty_bar + 8u * as);
                    else
                        mbar_arrive(empty_bar + 8u * as);
                }
                float gm[4];
#pragma unroll
                for (int g = 0; g < 4; g++) gm[g] = max16(v[g]);
                if (t >= full_tiles) {
                    // the last tile(s) of the database: some columns are past its end
                    const int64_t lrow0 = row0 + (int64_t)t * MROWS;
                    const int64_t left = s_end - lrow0;
                    const int nvalid = left >= 64 ? 64 : (int)max(left, (int64_t)0);
#pragma unroll
                    for (int g = 0; g < 4; g++) examine(v[g], g, nvalid, lrow0);
                } else if (fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3])) > thr_f) {
                    const int64_t lrow0 = row0 + (int64_t)t * MROWS;
#pragma unroll
                    for (int g = 0; g < 4; g++)
                        if (gm[g] > thr_f) collect(v[g], g, 16, lrow0, (int)gm[g]);
                }
                overflow_check(t);
            }
        } else {
        int64_t cur_row = tile_row(0) + half * 64;
        int in_run = (int)(tile0 & run_mask);
        for (int t = 0; t < ntiles; t++) {
            const uint32_t T = (uint32_t)(tbase + t);
            const int as = (int)(T & 1u);
            mbar_wait(smem_u32(&sm->acc_full[as]), (T >> 1) & 1u);
            tc_fence_after();
            const int64_t lrow0 = cur_row;  // first database row of this warp's 64 columns
            const int64_t left = s_end - lrow0;
            const int nvalid = left >= 64 ? 64 : (int)max(left, (int64_t)0);
            // advance the tile cursor (dense scan: +128 rows; strided sample: +128 inside a run, a jump at its end)
            if (++in_run > (int)run_mask) {
                in_run = 0;
                cur_row += p.run_stride - run_mask * MROWS;
            } else {
                cur_row += MROWS;
            }
            const uint32_t acc = lane_base + (uint32_t)as * MROWS;
            if (p.dbg) {
                // tests only: every distance of the tile (a second, unpipelined read of the accumulator)
                for (int g = 0; g < 4; g++) {
                    int w[16];
                    __syncwarp();
                    tmem_ld16(acc + 16 * g, w);
                    tmem_wait_ld();
                    if (qvalid)
                        for (int j = 0; j < 16; j++)
                            if (16 * g + j < nvalid)
                                p.dbg[(size_t)(q0 + q) * p.dbg_stride + (lrow0 + 16 * g + j)] = pcq - (F4 ? (int)__int_as_float(w[j]) : w[j]);
                }
            }
            // 64 columns in 4 groups of 16, software-pipelined over two register sets: the tcgen05.ld of group g + 1 is in
            // flight while group g is examined; the accumulator goes back to the MMA issuer as soon as the last group has landed
            int v[2][16];
            __syncwarp();
            tmem_ld16(acc, v[0]);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                tmem_wait_ld();
                if (g < 3)
                    tmem_ld16(acc + 16 * (g + 1), v[(g + 1) & 1]);
                else
                    release_acc(as);
                examine(v[g & 1], g, nvalid, lrow0);
            }
            overflow_check(t);
        }
        }
        if (samp && qvalid) {
            unsigned short* o_ = p.sample_out + (((size_t)strip * p.nq + q0 + q) * 2 + half) * SAMPLE_KEEP;
#pragma unroll
            for (int i = 0; i < SAMPLE_KEEP; i++) o_[i] = sbest[i] >= 0x7fff ? (unsigned short)0xFFFF : (unsigned short)sbest[i];
        }
        group_sync<EPI_THREADS>(BAR_CONSUMERS);
        // final compaction: every list leaves the kernel with at most k keys (bounds the merge's working set)
        for (int qq = warp; qq < qt && !p.sample_mode; qq += EPI_WARPS) {
            const int n = sm->cnt_s[qq];
            if (n > p.k)
                compact_list_warp(p.lists + ((size_t)strip * p.nq + q0 + qq) * p.cap, n, p.k, &sm->sc[warp], lane, BAR_WARP0 + warp,
                                  &sm->cnt_s[qq], &sm->tau_s[qq]);
        }
        group_sync<EPI_THREADS>(BAR_CONSUMERS);
        if (qvalid && half == 0) p.counts[(size_t)strip * p.nq + q0 + q] = sm->cnt_s[q];
    }
    };
    auto other_role = [&]() {
    seg_begin_sync();
    if (warp == WARP_TMA) {
        // ===================== raw-code producer =====================
        if (lane == 0) {
            uint32_t s = prod_s, ph = prod_ph;
            // Lockstep throttle (dense pass, CTA pairs): the pairs of one strip read the same rows, and only the first read
            // of a row comes from HBM as long as the others follow within the L2's reach.  Left alone the pairs drift apart
            // (ncu: DRAM reads 2x the code array); here the leader's producer publishes its tile counter every 16 tiles and
            // does not run more than lock_window tiles ahead of the slowest pair of its strip - which costs nothing, since
            // the launch ends with the slowest pair anyway.
            const int ncols = p.seg_cols > 0 ? p.seg_cols : (int)(gridDim.x / CG);
            const bool full_strip = p.seg_cols == 0 || (int)(blockIdx.x / CG) < p.seg_cols * p.seg_full;
            volatile int* prog = (p.progress && rank == 0 && ncols > 1 && full_strip) ? p.progress + (size_t)strip * ncols : nullptr;
            const int mycol = p.seg_cols > 0 ? (int)(blockIdx.x / CG) % p.seg_cols : (int)(blockIdx.x / CG);
            for (int t = 0; t < ntiles; t++) {
                if (prog && (t & 15) == 0) {
                    prog[mycol] = t;
                    const int lim = t - p.lock_window;
                    if (lim > 0)
                        for (int c = 0; c < ncols; c++)
                            while (prog[c] < lim) __nanosleep(256);
                }
                mbar_wait_relaxed(smem_u32(&sm->raw_empty[s]), ph ^ 1u, 256);
                mbar_expect_tx(smem_u32(&sm->raw_full[s]), STAGE_BYTES_RAW);
                tma_load_2d(smem_u32(raw_mem) + s * (uint32_t)STAGE_BYTES_RAW, &tmap, 0,
                            (int)(tile_row(t) + (int64_t)rank * MY_ROWS), smem_u32(&sm->raw_full[s]));
                if (++s == (uint32_t)raw_stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            if (prog) prog[mycol] = 0x7fffffff;
            prod_s = s;
            prod_ph = ph;
        }
    } else if (warp >= WARP_MMA && warp < WARP_MMA + MMA_WARPS) {
        // ===================== MMA issuers =====================
        // Two issuer warps alternate tiles (the single-thread issue path - waits, descriptor arithmetic, 4 MMAs and a
        // commit per K-block - costs about as many cycles per tile as the e2m1 MMAs themselves take to execute).  Tiles
        // of different issuers use different accumulators and barrier-guarded B stages, so their order is free.
        // The whole warp walks the loop (everything stays warp-uniform); one elected lane issues.  Per K-block: one wait,
        // 4 MMAs (32 bytes of K each) whose descriptors are immediates off the tile's base, one commit.
        const uint64_t desc0 = umma_desc_sw128(smem_u32(b_mem));
        const uint32_t full0 = smem_u32(&sm->b_full[0]), empty0 = smem_u32(&sm->b_empty[0]);
        // (int8 tiles all share the same 8 stages: a second issuer would wait on a barrier phase two ahead of the completed
        // one, which a parity wait cannot express - so the int8 kind keeps a single issuer.)
        constexpr int ISSUERS = F4 ? MMA_WARPS : 1;
        for (int t = 0; t < ntiles && warp - WARP_MMA < ISSUERS && rank == 0; t++) {
            const uint32_t T = (uint32_t)(tbase + t);  // tile number over all segments of this CTA
            if ((int)(T % (uint32_t)ISSUERS) != warp - WARP_MMA) continue;
            const int as = (int)(T & 1u);
            const uint32_t grp = T & (uint32_t)(GROUPS - 1);
            const uint32_t ph = (T / (uint32_t)GROUPS) & 1u;
            const uint64_t desc_t = desc0 + (uint64_t)(grp * (uint32_t)(KBLOCKS * (STAGE_BYTES_B >> 4)));
            const uint32_t full_t = full0 + grp * (uint32_t)(KBLOCKS * 8), empty_t = empty0 + grp * (uint32_t)(KBLOCKS * 8);
            mbar_wait(smem_u32(&sm->acc_empty[as]), ((T >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem + TMEM_D_COL + (uint32_t)as * MROWS;
#pragma unroll
            for (int kb = 0; kb < KBLOCKS; kb++) {
                mbar_wait(full_t + kb * 8, ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        const uint64_t desc = desc_t + (uint64_t)(kb * (STAGE_BYTES_B >> 4) + k4 * 2);
                        const uint32_t a_tmem = tmem + TMEM_A_COL + (uint32_t)(kb * 4 + k4) * 8;
                        if constexpr (CG == 2)
                            umma_f4_ts_pair(d_tmem, a_tmem, desc, IDESC_F4_PAIR, tmem + TMEM_SFA_COL, tmem + TMEM_SFB_COL, (kb | k4) != 0);
                        else if constexpr (F4)
                            umma_f4_ts(d_tmem, a_tmem, desc, IDESC_F4, tmem + TMEM_SFA_COL, tmem + TMEM_SFB_COL, (kb | k4) != 0);
                        else
                            umma_i8_ts(d_tmem, a_tmem, desc, IDESC_I8, (kb | k4) != 0);
                    }
                    if constexpr (CG == 2) {
                        tc_commit_pair(empty_t + kb * 8);  // frees the stage in BOTH CTAs
                        if (kb == KBLOCKS - 1) tc_commit_pair(smem_u32(&sm->acc_full[as]));
                    } else {
                        tc_commit(empty_t + kb * 8);
                        if (kb == KBLOCKS - 1) tc_commit(smem_u32(&sm->acc_full[as]));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= WARP_EXP0 + EXP_WARPS) {
        // pad warp (VAR = 4 only): completes the last warpgroup for setmaxnreg, no work
    } else if (warp >= WARP_EXP0) {
        // ===================== expanders: code bits -> one 128-byte K-block row of the B operand =====================
        const int et = tid - WARP_EXP0 * 32;
        constexpr int KB_STEP = (EXP_WARPS * 32) / MY_ROWS;  // K-blocks in flight across the 256 expander threads (2; pair: 4)
        constexpr int UNITS = KBLOCKS / KB_STEP;             // K-blocks per thread per tile
        const int row = et & (MY_ROWS - 1);
        const int par = et / MY_ROWS;  // this thread expands K-blocks par, par + KB_STEP, ... of its row
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t row_off = (uint32_t)row * 128u;
        const uint32_t empty0 = smem_u32(&sm->b_empty[0]);
        // the "stage full" barriers the MMA issuer waits on live in the leader CTA
        const uint32_t full0 = CG == 2 ? mapa_rank(smem_u32(&sm->b_full[0]), 0) : smem_u32(&sm->b_full[0]);
        const uint32_t bmem0 = smem_u32(b_mem) + row_off;
        for (int t = 0; t < ntiles; t++) {
            const uint32_t T = (uint32_t)(tbase + t);
            const uint32_t grp = T & (uint32_t)(GROUPS - 1);
            const uint32_t ph = (T / (uint32_t)GROUPS) & 1u;
            const uint32_t stage0 = grp * (uint32_t)KBLOCKS + (uint32_t)par;
            mbar_wait_relaxed(smem_u32(&sm->raw_full[rs]), rph, 64);
            const uint32_t raddr = smem_u32(raw_mem) + rs * (uint32_t)STAGE_BYTES_RAW + row_off;
            uint4 c[4];
            // I8: K-block kb = raw chunk kb (16 bytes);  F4: K-block kb = raw chunks 2 kb, 2 kb + 1 (32 bytes)
#pragma unroll
            for (int j = 0; j < (F4 ? 2 * UNITS : UNITS); j++) {
                const uint32_t chunk = F4 ? (uint32_t)(2 * (par + KB_STEP * (j >> 1)) + (j & 1)) : (uint32_t)(par + KB_STEP * j);
                c[j] = lds128(raddr + ((chunk ^ sw) << 4));
            }
#pragma unroll
            for (int j = 0; j < UNITS; j++) {
                const uint32_t s = stage0 + KB_STEP * j;
                mbar_wait_relaxed(empty0 + s * 8, ph ^ 1u, 64);
                const uint32_t baddr = bmem0 + s * (uint32_t)STAGE_BYTES_B;
                if constexpr (!F4) {
                    const uint32_t w[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const uint32_t ch = (uint32_t)(2 * i + h);
                            sts128(baddr + ((ch ^ sw) << 4), (w[i] >> (4 * h)) & 0x01010101u, (w[i] >> (4 * h + 1)) & 0x01010101u,
                                   (w[i] >> (4 * h + 2)) & 0x01010101u, (w[i] >> (4 * h + 3)) & 0x01010101u);
                        }
                    }
                } else {
                    // word i of the 32 code bytes -> 16-byte chunk i: plane t holds bit (t + 4 j) of the word in nibble j.  Planes
                    // 0-2 keep the bit where it is - as e2m1 that reads 0.5 (0x1), 1.0 (0x2), 2.0 (0x4) - and the query side
                    // carries +-2, +-1, +-0.5 so every product is +-1; only plane 3 (bit 3 = the sign bit) needs a shift.
                    const uint32_t w[8] = {c[2 * j].x, c[2 * j].y, c[2 * j].z, c[2 * j].w, c[2 * j + 1].x, c[2 * j + 1].y, c[2 * j + 1].z,
                                           c[2 * j + 1].w};
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        sts128(baddr + (((uint32_t)i ^ sw) << 4), w[i] & 0x11111111u, w[i] & 0x22222222u, w[i] & 0x44444444u,
                               (w[i] >> 1) & 0x44444444u);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2)
                        mbar_arrive_cluster(full0 + s * 8);
                    else
                        mbar_arrive(full0 + s * 8);
                }
            }
            // the raw tile goes back to the producer only now: every c[j] has been consumed by real instructions, so the
            // shared-memory reads above are known to have completed (an arrive right after the ld.shared can overtake them)
            if (lane == 0) mbar_arrive(smem_u32(&sm->raw_empty[rs]));
            if (++rs == (uint32_t)raw_stages) {
                rs = 0;
                rph ^= 1u;
            }
        }
    }
    };
    if constexpr (VAR == 4) {
        // The CTA's register pool is what it was launched with: 20 warps x 96.  8 epilogue warps x 128 + 12 others x 72 = 1888
        // <= 1920 per lane; per scheduler partition 2 x 128 + 3 x 72 = 472 <= 480.  (136 / 80 exceeds the pool: the inc never
        // returns.)
        if (warp < EPI_WARPS) {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
            for (int sgi = 0; segment(sgi); sgi++) {
                epilogue_role(sgi);
                tbase += ntiles;
                seg_end_sync();
            }
        } else {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
            for (int sgi = 0; segment(sgi); sgi++) {
                other_role();
                tbase += ntiles;
                seg_end_sync();
            }
        }
    } else {
        for (int sgi = 0; segment(sgi); sgi++) {
            if (warp < EPI_WARPS)
                epilogue_role(sgi);
            else
                other_role();
            tbase += ntiles;
            seg_end_sync();
        }
    }

    if (warp == WARP_MMA) {
        tc_fence_after();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Few queries per pass (3 .. 64): the operand roles are swapped.  The MMA of the kernel above costs 8 cycles per database
// row whatever the number of queries, which caps a single 128-query tile at 4.5 TB/s of codes (measured 3.1 TB/s).  Here
// the DATABASE rows are the M dimension and the queries the N dimension (N = nq rounded up to 16), so one tile costs
// 8 N cycles of tensor time:
//   * B operand = the queries (+-2, +-1, +-0.5 e2m1), expanded once into shared memory (N x 512 bytes);
//   * A operand = 128 database rows, expanded by the expander threads straight into TENSOR MEMORY with tcgen05.st
//     (thread = row = TMEM lane; two 128-column buffers), so the expansion never touches shared memory;
//   * D = 128 rows x N queries (two buffers); epilogue lane = database row, column = query.
// Same lists / counts / thresholds contract as the kernels above, e2m1 kind only.
constexpr int FEW_MAXQ = 32;   // <= 32 queries: two 16-column groups of the accumulator
constexpr int FEW_WARP_MMA = 4, FEW_WARP_TMA = 5, FEW_WARP_EXP0 = 6, FEW_EXP_WARPS = 8;
constexpr int FEW_THREADS = (FEW_WARP_EXP0 + FEW_EXP_WARPS) * 32;
constexpr int FEW_EPI_WARPS = 4, FEW_EPI_THREADS = FEW_EPI_WARPS * 32;
// tensor memory: A = 2 x 128 columns, D = 2 buffers of 64 columns, 8 columns of bias operand, block scales = the last 64 columns
constexpr uint32_t FEW_A_COL = 0, FEW_D_COL = 256, FEW_SF_COL = 448;
constexpr int FEW_MAX_RAW = 8;

// What makes it scale to 64 query columns (the first generation, with one threshold per column, was slower than the
// 128-query-tile kernel beyond 32 queries: 8.1 ms per 100 M codes at 64):
//   * ONE threshold for every query column.  A 17th MMA per tile adds a per-query bias to the accumulator: the A operand of
//     that MMA is a constant (62 elements 4.0, 2 elements 1.0, written to tensor memory once), its B operand is a fifth
//     K-block of the query rows in shared memory that encodes bias = -(popc(q) - tau) as a sum of e2m1 products
//     (4 x {6, 4, 3, 2, 1.5, 1, 0.5} = 24, 16, 12, 8, 6, 4, 2, plus 1 x 1).  The accumulator then holds dot - thr(q): a row has
//     a survivor <=> the MAXIMUM over its columns is > 0 - a tree of 3-input maxima and one compare per row instead of a
//     subtract and a maximum per column (which is what made the per-column form slower than linear in N).  Thresholds that
//     tighten later (compaction) make the bias conservative, never wrong: the exact test is redone on the survivors.
//   * The expanders split a tile by (row, half of K): all eight warps work on every tile (4 x LDS.128, 64 ALU, 2 x tcgen05.st.x32
//     per thread).
//   * The accumulator goes back to the issuer before anything is examined; barrier addresses are loop invariants.
//   * Survivors go through a mask and ONE out-of-line append (wide_append below).
constexpr int WIDE_MAXQ = 64, WIDE128_MAXQ = 128;
constexpr uint32_t WIDE_BIAS_COL = 384;  // 8 columns: the constant A operand of the bias MMA
constexpr int WIDE_BIAS_MAX = 1440;      // 60 x 24: beyond every |dot| <= 1024, i.e. "always" / "never"

// One survivor of the wide kernel goes to its query's list.  NOT inlined: the body (exact re-test against the current
// threshold, key, slot atomic, store) exists once in the kernel instead of once per accumulator column - unrolled into every
// column the survivor path was ~40 KB of code that ran out of the instruction cache (ncu: stall_no_inst, ~5600 cycles per
// visit, which is what made the per-column form of this kernel slower than linear in the number of queries).
__device__ __noinline__ void wide_append(const float* bias_s, const float* thr_s, const int* pcq_s, int* cnt_s, uint64_t* lists0,
                                         const unsigned long long* key_lo, int cap, unsigned long long pos, int q, float f) {
    const float dot = f - bias_s[q];  // exact: integers below 2^24
    if (dot > thr_s[q]) {             // the CURRENT threshold (the bias may be older)
        const unsigned long long key = ((unsigned long long)(pcq_s[q] - (int)dot) << VRQ_KEY_POS_BITS) | pos;
        if (key_lo == nullptr || key > key_lo[q]) {
            const int slot = atoms_add(smem_u32(&cnt_s[q]), 1);
            if (slot >= cap) __trap();
            lists0[(size_t)q * cap + slot] = key;
        }
    }
}
// element j (runtime) of 16 registers: a tree of selects
__device__ __forceinline__ int sel16(const int (&w)[16], int j) {
    int a[8], b[4], c[2];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (j & 1) ? w[2 * i + 1] : w[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 2; i++) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
    return (j & 8) ? c[1] : c[0];
}

template <int MAXQ>
struct WideSmem {
    unsigned long long raw_full[FEW_MAX_RAW], raw_empty[FEW_MAX_RAW];
    unsigned long long a_full[4], a_empty[4], acc_full[2], acc_empty[2];  // a_*: per A buffer (2 used) or per K quarter (HALVES)
    uint32_t tmem_base;
    int tau_s[MAXQ];
    int cnt_s[MAXQ];
    int pcq_s[MAXQ];
    __align__(16) float thr_s[MAXQ];   // current exact threshold of the query: survivor <=> dot > thr_s
    __align__(16) float bias_s[MAXQ];  // what the bias MMA added to the query's column
    SelectScratch sc[8];  // one per epilogue warp (4 or 8)
};

// MAXQ = 32 / 64: two A buffers of 128 columns, handed over whole; 4 epilogue warps (one per TMEM lane quadrant).
// MAXQ = 128: tensor memory has room for ONE A buffer next to two 128-column accumulators, so it is handed over per K HALF
// (64 columns): one half is refilled while the MMAs of the other run.  8 epilogue warps: warps w and w + 4 share a lane
// quadrant and split the query columns (64 each, as many as an epilogue warp of the 128-query-tile kernel examines) - with 4
// warps holding 128 columns each (184 registers via setmaxnreg) the epilogue was the bound (128 queries: 5.0 ms per 100 M
// codes).  4 expander warps (one per lane quadrant, whole rows) keep the block at 14 warps = 4 per scheduler partition;
// with 8 (18 warps, 96 registers) everything was slower.  Per K quarter instead of per half: the same within 3 %.
template <int MAXQ>
struct WideCfg {
    static constexpr int EPIW = MAXQ > 64 ? 8 : 4;  // epilogue warps
    static constexpr int WARP_MMA = EPIW, WARP_TMA = EPIW + 1, WARP_EXP0 = EPIW + 2;
    static constexpr int EXPW = MAXQ > 64 ? 4 : 8;   // expander warps (MAXQ = 128: one per lane quadrant, whole rows)
    static constexpr int THREADS = (WARP_EXP0 + EXPW) * 32;  // 14 warps either way: at most 4 per scheduler partition, 144 registers
    static constexpr int PARTS = 2;                  // MAXQ = 128: hand-overs of the single A buffer per tile (K halves)
};
template <int MAXQ>
__global__ void __launch_bounds__(WideCfg<MAXQ>::THREADS, 1)
hamming_scan_mma_wide_kernel(const __grid_constant__ CUtensorMap tmap, ScanParams p, int raw_stages, int npad) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* raw_mem = base;                                          // [raw_stages][128 rows][128 B], TMA SWIZZLE_128B
    uint8_t* q_mem = raw_mem + (size_t)raw_stages * STAGE_BYTES_RAW;  // [5 K-blocks][npad queries][128 B], same swizzle
    WideSmem<MAXQ>* sm = (WideSmem<MAXQ>*)(q_mem + (size_t)5 * MAXQ * 128);
    constexpr bool HALVES = MAXQ > 64;
    constexpr int EPIW = WideCfg<MAXQ>::EPIW, EPI_THREADS_ = EPIW * 32;
    constexpr int W_MMA = WideCfg<MAXQ>::WARP_MMA, W_TMA = WideCfg<MAXQ>::WARP_TMA, W_EXP0 = WideCfg<MAXQ>::WARP_EXP0;
    constexpr int NGW = (MAXQ / 16) / (EPIW / 4);  // 16-column groups per epilogue warp
    constexpr uint32_t D_COL = HALVES ? 128 : FEW_D_COL, D_STRIDE = HALVES ? 128 : 64, A_STRIDE = HALVES ? 0 : 128;
    constexpr int NG = MAXQ / 16;  // 16-column groups of the accumulator
    // barrier a_full / a_empty [i]: i = A buffer (tile parity), or with HALVES i = K half; their phase flips every second tile
    // resp. every tile
    auto a_phase = [](uint32_t t) -> uint32_t { return HALVES ? (t & 1u) : ((t >> 1) & 1u); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nq = p.nq;  // <= MAXQ, one query tile
    const int strip = blockIdx.y;
    const int tiles_per_strip = (int)(p.rows_per_strip / MROWS);
    const int64_t tile0 = (int64_t)strip * tiles_per_strip;
    const int ntiles = (int)max((int64_t)0, min((int64_t)tiles_per_strip, p.total_tiles - tile0));
    const int64_t run_mask = ((int64_t)1 << p.run_shift) - 1;
    auto tile_row = [&](int t) -> int64_t {
        const int64_t i = tile0 + t;
        return p.row_begin + (i >> p.run_shift) * p.run_stride + (i & run_mask) * MROWS;
    };
    const int64_t s_end = p.row_end;
    if (p.guard && *p.guard == 0) return;

    if (tid == 0) {
        for (int s = 0; s < raw_stages; s++) {
            mbar_init(smem_u32(&sm->raw_full[s]), 1);
            mbar_init(smem_u32(&sm->raw_empty[s]), WideCfg<MAXQ>::EXPW);
        }
        for (int s = 0; s < 4; s++) {
            mbar_init(smem_u32(&sm->a_full[s]), WideCfg<MAXQ>::EXPW);  // every expander warp contributes to every hand-over
            mbar_init(smem_u32(&sm->a_empty[s]), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(smem_u32(&sm->acc_full[s]), 1);
            mbar_init(smem_u32(&sm->acc_empty[s]), EPIW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < MAXQ) {
        sm->pcq_s[tid] = 0;
        sm->cnt_s[tid] = 0;
        sm->tau_s[tid] = tid < nq ? (p.tau0 ? min(p.tau0[tid], TAU_INF - 1) + p.tau_bias : TAU_INF) : 0;
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    // ---- the queries become the B operand in shared memory: row = query, K-block kb = code words 8 kb .. 8 kb + 7, chunk i
    //      of the 128-byte row = word 8 kb + i as four planes (nibble j of plane t = bit t + 4 j) of +-(1 / plane value)
    for (int idx = tid; idx < npad * 32; idx += WideCfg<MAXQ>::THREADS) {
        const int q = idx >> 5, W = idx & 31;
        const bool qvalid = q < nq;
        const uint32_t w = qvalid ? __ldg(reinterpret_cast<const uint32_t*>(p.queries + (size_t)q * CODE_BYTES) + W) : 0u;
        if (qvalid) atomicAdd(&sm->pcq_s[q], __popc(w));
        uint32_t v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t mag = t == 0 ? 0x44444444u : (t == 1 ? 0x22222222u : 0x11111111u);
            v[t] = qvalid ? ((mag | 0x88888888u) ^ (((w >> t) & 0x11111111u) << 3)) : 0u;
        }
        const int kb = W >> 3, i = W & 7;
        sts128(smem_u32(q_mem) + (uint32_t)(kb * npad * 128 + q * 128 + ((i ^ (q & 7)) << 4)), v[0], v[1], v[2], v[3]);
    }
    if (warp < 4) {
        // every block scale (UE8M0) = 0x7F = 2^0; the constant A operand of the bias MMA: K positions 0 .. 61 = 4.0 (0x6),
        // 62, 63 = 1.0 (0x2)
        uint32_t one[8];
#pragma unroll
        for (int t = 0; t < 8; t++) one[t] = 0x7F7F7F7Fu;
#pragma unroll
        for (int c = 0; c < 64; c += 8) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + FEW_SF_COL + c, one);
        uint32_t ca[8];
#pragma unroll
        for (int t = 0; t < 7; t++) ca[t] = 0x66666666u;
        ca[7] = 0x22666666u;
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + WIDE_BIAS_COL, ca);
        tmem_wait_st();
    }
    __syncthreads();  // pcq_s complete
    if (tid < npad) {
        // bias(q) = -(popc(q) - tau(q)) clamped to "always" / "never"; padding columns never survive
        const int q = tid;
        const int tau = sm->tau_s[q];
        int bias = -WIDE_BIAS_MAX;
        float thr = 3.0e9f;
        if (q < nq) {
            const int t = sm->pcq_s[q] - tau;  // survivor <=> dot > t
            thr = (float)t;
            bias = tau >= TAU_INF ? WIDE_BIAS_MAX : max(-WIDE_BIAS_MAX, min(WIDE_BIAS_MAX, -t));
        }
        sm->thr_s[q] = thr;
        sm->bias_s[q] = (float)bias;
        // |bias| = 24 n24 + rem (even, < 24: at most two more products) + odd; nibble p of the 64: products of the constant
        // operand (4.0 at p < 62, 1.0 at p = 62) with e2m1 codes 7 = 6.0, 6 = 4.0, 5 = 3.0, 4 = 2.0, 3 = 1.5, 2 = 1.0, 1 = 0.5
        const uint32_t sgn = bias < 0 ? 0x8u : 0x0u;
        const int mag = bias < 0 ? -bias : bias;
        const int odd = mag & 1, n24 = (mag - odd) / 24, rem = (mag - odd) % 24;
        uint32_t c1 = 0, c2 = 0;  // e2m1 codes of the (up to two) products that make up rem
        switch (rem) {
            case 2: c1 = 1; break;
            case 4: c1 = 2; break;
            case 6: c1 = 3; break;
            case 8: c1 = 4; break;
            case 10: c1 = 4, c2 = 1; break;
            case 12: c1 = 5; break;
            case 14: c1 = 5, c2 = 1; break;
            case 16: c1 = 6; break;
            case 18: c1 = 6, c2 = 1; break;
            case 20: c1 = 6, c2 = 2; break;
            case 22: c1 = 6, c2 = 3; break;
            default: break;
        }
        uint32_t wd[8];
#pragma unroll
        for (int wi = 0; wi < 8; wi++) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int pos = 8 * wi + j;
                uint32_t code = 0;
                if (pos < n24) code = 7;
                else if (pos == n24) code = c1;
                else if (pos == n24 + 1) code = c2;
                if (pos == 62) code = odd ? 2u : 0u;
                if (pos == 63) code = 0;
                if (code) code |= sgn;
                x |= code << (4 * j);
            }
            wd[wi] = x;
        }
        const uint32_t rowa = smem_u32(q_mem) + (uint32_t)(4 * npad * 128 + q * 128);
        sts128(rowa + (((uint32_t)0 ^ (uint32_t)(q & 7)) << 4), wd[0], wd[1], wd[2], wd[3]);
        sts128(rowa + (((uint32_t)1 ^ (uint32_t)(q & 7)) << 4), wd[4], wd[5], wd[6], wd[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == W_TMA) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int t = 0; t < ntiles; t++) {
                mbar_wait_relaxed(smem_u32(&sm->raw_empty[s]), ph ^ 1u, 128);
                mbar_expect_tx(smem_u32(&sm->raw_full[s]), STAGE_BYTES_RAW);
                tma_load_2d(smem_u32(raw_mem) + s * (uint32_t)STAGE_BYTES_RAW, &tmap, 0, (int)tile_row(t), smem_u32(&sm->raw_full[s]));
                if (++s == (uint32_t)raw_stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: 16 + 1 MMAs (M = 128 rows, N = npad queries, K = 64) per tile ===================
        const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | (1u << 23) | ((uint32_t)(MQ >> 4) << 24);
        const uint64_t qdesc0 = umma_desc_sw128(smem_u32(q_mem));
        const uint32_t kb_step = (uint32_t)(npad * 128) >> 4;
        uint32_t bar_a_full = smem_u32(&sm->a_full[0]), bar_a_empty = smem_u32(&sm->a_empty[0]);
        uint32_t bar_acc_full = smem_u32(&sm->acc_full[0]), bar_acc_empty = smem_u32(&sm->acc_empty[0]);
        asm volatile("" : "+r"(bar_a_full), "+r"(bar_a_empty), "+r"(bar_acc_full), "+r"(bar_acc_empty));
        for (int t = 0; t < ntiles; t++) {
            const uint32_t ab = (uint32_t)t & 1u, ph = ((uint32_t)t >> 1) & 1u, aph = a_phase((uint32_t)t);
            const uint32_t d_tmem = tmem + D_COL + ab * D_STRIDE, a0 = tmem + FEW_A_COL + ab * A_STRIDE;
            if constexpr (!HALVES) {
                mbar_wait(bar_a_full + 8u * ab, aph);  // the tile is in tensor memory
                mbar_wait(bar_acc_empty + 8u * ab, ph ^ 1u);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int s = 0; s < 16; s++)
                        umma_f4_ts(d_tmem, a0 + 8 * s, qdesc0 + (uint64_t)((s >> 2) * kb_step + (s & 3) * 2), idesc, tmem + FEW_SF_COL,
                                   tmem + FEW_SF_COL + 32, s != 0);
                    umma_f4_ts(d_tmem, tmem + WIDE_BIAS_COL, qdesc0 + (uint64_t)(4 * kb_step), idesc, tmem + FEW_SF_COL, tmem + FEW_SF_COL + 32, 1);
                    tc_commit(bar_a_empty + 8u * ab);
                    tc_commit(bar_acc_full + 8u * ab);
                }
                __syncwarp();
            } else {
                mbar_wait(bar_acc_empty + 8u * ab, ph ^ 1u);
                constexpr int PARTS = WideCfg<MAXQ>::PARTS, SPP = 16 / PARTS;  // MMAs per piece
#pragma unroll
                for (int part = 0; part < PARTS; part++) {
                    mbar_wait(bar_a_full + 8u * part, aph);  // piece `part` of the tile's A operand is in tensor memory
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int s = SPP * part; s < SPP * part + SPP; s++)
                            umma_f4_ts(d_tmem, a0 + 8 * s, qdesc0 + (uint64_t)((s >> 2) * kb_step + (s & 3) * 2), idesc, tmem + FEW_SF_COL,
                                       tmem + FEW_SF_COL + 32, s != 0);
                        if (part == PARTS - 1)
                            umma_f4_ts(d_tmem, tmem + WIDE_BIAS_COL, qdesc0 + (uint64_t)(4 * kb_step), idesc, tmem + FEW_SF_COL, tmem + FEW_SF_COL + 32,
                                       1);
                        tc_commit(bar_a_empty + 8u * part);
                        if (part == PARTS - 1) tc_commit(bar_acc_full + 8u * ab);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= W_EXP0) {
        // ===================== expanders: thread = (database row = TMEM lane, half of K); all 8 warps on every tile ========
        const int kh = (warp - W_EXP0) >> 2;  // which 64 code bytes (chunks 4 kh .. 4 kh + 3) -> which 64 columns
        const int row = (warp & 3) * 32 + lane;      // a warp may only touch the TMEM lane quadrant warp_id % 4
        const uint32_t sw = (uint32_t)(row & 7);
        uint32_t a_buf0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + FEW_A_COL + (uint32_t)kh * 64;
        uint32_t a_quad0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + FEW_A_COL;
        asm volatile("" : "+r"(a_quad0));
        uint32_t raw0 = smem_u32(raw_mem) + (uint32_t)row * 128u;
        uint32_t bar_raw_full = smem_u32(&sm->raw_full[0]), bar_raw_empty = smem_u32(&sm->raw_empty[0]);
        uint32_t bar_a_full = smem_u32(&sm->a_full[0]), bar_a_empty = smem_u32(&sm->a_empty[0]);
        asm volatile("" : "+r"(a_buf0), "+r"(raw0), "+r"(bar_raw_full), "+r"(bar_raw_empty), "+r"(bar_a_full), "+r"(bar_a_empty));
        uint32_t rs = 0, rph = 0;
        for (int t = 0; t < ntiles; t++) {
            const uint32_t ab = (uint32_t)t & 1u, aph = a_phase((uint32_t)t);
            mbar_wait_relaxed(bar_raw_full + 8u * rs, rph, 32);
            const uint32_t raddr = raw0 + rs * (uint32_t)STAGE_BYTES_RAW;
            if constexpr (HALVES) {
                // one expander warp per lane quadrant: the whole 128-byte row, handed over in PARTS pieces of 128 / PARTS columns;
                // the planes of piece i + 1 are computed while the tcgen05.st of piece i are in flight
                constexpr int PARTS = WideCfg<MAXQ>::PARTS, QPP = 4 / PARTS;  // 32-column quarters per piece
                uint4 c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) c[j] = lds128(raddr + (((uint32_t)j ^ sw) << 4));
                uint32_t v[2][QPP][32];
                auto planes = [&](const int quarter, uint32_t(&o)[32]) {
                    const uint32_t w[8] = {c[2 * quarter].x, c[2 * quarter].y, c[2 * quarter].z, c[2 * quarter].w,
                                           c[2 * quarter + 1].x, c[2 * quarter + 1].y, c[2 * quarter + 1].z, c[2 * quarter + 1].w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        o[4 * i + 0] = w[i] & 0x11111111u;
                        o[4 * i + 1] = w[i] & 0x22222222u;
                        o[4 * i + 2] = w[i] & 0x44444444u;
                        o[4 * i + 3] = (w[i] >> 1) & 0x44444444u;
                    }
                };
#pragma unroll
                for (int qq = 0; qq < QPP; qq++) planes(qq, v[0][qq]);
#pragma unroll
                for (int part = 0; part < PARTS; part++) {
                    mbar_wait(bar_a_empty + 8u * part, aph ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int qq = 0; qq < QPP; qq++) tmem_st32(a_quad0 + 32 * (part * QPP + qq), v[part & 1][qq]);
                    if (part + 1 < PARTS) {
#pragma unroll
                        for (int qq = 0; qq < QPP; qq++) planes((part + 1) * QPP + qq, v[(part + 1) & 1][qq]);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_a_full + 8u * part);
                }
            } else {
                uint4 c[4];
#pragma unroll
                for (int j = 0; j < 4; j++) c[j] = lds128(raddr + (((uint32_t)(4 * kh + j) ^ sw) << 4));
                mbar_wait(bar_a_empty + 8u * ab, aph ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int part = 0; part < 2; part++) {
                    // words 8 (2 kh + part) .. + 7 of the code -> columns 64 kh + 32 part .. + 31 (column 4 W + t = plane t of word W)
                    const uint32_t w[8] = {c[2 * part].x, c[2 * part].y, c[2 * part].z, c[2 * part].w,
                                           c[2 * part + 1].x, c[2 * part + 1].y, c[2 * part + 1].z, c[2 * part + 1].w};
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        v[4 * i + 0] = w[i] & 0x11111111u;
                        v[4 * i + 1] = w[i] & 0x22222222u;
                        v[4 * i + 2] = w[i] & 0x44444444u;
                        v[4 * i + 3] = (w[i] >> 1) & 0x44444444u;
                    }
                    tmem_st32(a_buf0 + ab * A_STRIDE + 32 * part, v);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_full + 8u * ab);
            }
            if (lane == 0) mbar_arrive(bar_raw_empty + 8u * rs);  // every c[j] has been consumed by real instructions
            if (++rs == (uint32_t)raw_stages) {
                rs = 0;
                rph ^= 1u;
            }
        }
    } else {
        // ===================== epilogue: lane = database row, column = query ==========================================
        // warp & 3 = TMEM lane quadrant (32 database rows); warp >> 2 (MAXQ = 128 only) = which half of the query columns
        const int quad = warp & 3, g0 = EPIW == 8 ? (warp >> 2) * NGW : 0;  // first 16-column group of this warp
        uint32_t acc0 = tmem + ((uint32_t)(quad * 32) << 16) + D_COL + 16u * (uint32_t)g0;
        uint32_t bar_acc_full = smem_u32(&sm->acc_full[0]), bar_acc_empty = smem_u32(&sm->acc_empty[0]);
        asm volatile("" : "+r"(acc0), "+r"(bar_acc_full), "+r"(bar_acc_empty));
        const int limit = p.compact_limit > 0 ? min(p.compact_limit, p.cap - p.group_tiles * MROWS) : p.cap - p.group_tiles * MROWS;
        uint64_t* const lists0 = p.lists + (size_t)strip * p.nq * p.cap;
        int until_check = p.group_tiles;
        const int ng = npad >> 4;  // column groups in use (npad is a multiple of 16)
        const bool has_dbg = p.dbg != nullptr;
        auto max16 = [](const int(&w)[16]) -> float {
            auto f = [&](int j) -> float { return __int_as_float(w[j]); };
            auto mx3 = [](float x, float y, float z) -> float { return fmaxf(fmaxf(x, y), z); };
            const float a0 = mx3(f(0), f(1), f(2)), a1 = mx3(f(3), f(4), f(5)), a2 = mx3(f(6), f(7), f(8));
            const float a3 = mx3(f(9), f(10), f(11)), a4 = mx3(f(12), f(13), f(14));
            return fmaxf(mx3(a0, a1, a2), mx3(a3, a4, f(15)));
        };
        for (int t = 0; t < ntiles; t++) {
            const uint32_t ab = (uint32_t)t & 1u;
            mbar_wait(bar_acc_full + 8u * ab, ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            int v[NGW][16];
            __syncwarp();
#pragma unroll
            for (int g = 0; g < NGW; g++)
                if (g0 + g < ng) tmem_ld16(acc0 + ab * D_STRIDE + 16 * g, v[g]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8u * ab);  // the accumulator goes back before anything is examined
            float gm[NGW];
            float mx = -3.0e9f;
#pragma unroll
            for (int g = 0; g < NGW; g++) {
                gm[g] = g0 + g < ng ? max16(v[g]) : -3.0e9f;
                mx = fmaxf(mx, gm[g]);
            }
            if (mx > 0.0f || has_dbg) {
                const int64_t lrow = tile_row(t) + quad * 32 + lane;
                if (lrow < s_end) {
                    if (has_dbg) {
                        // tests only: every distance of the tile
#pragma unroll
                        for (int g = 0; g < NGW; g++)
                            if (g0 + g < ng)
                                for (int j = 0; j < 16; j++) {
                                    const int q = 16 * (g0 + g) + j;
                                    if (q < nq) p.dbg[(size_t)q * p.dbg_stride + lrow] = sm->pcq_s[q] - (int)(__int_as_float(sel16(v[g], j)) - sm->bias_s[q]);
                                }
                    }
                    const unsigned long long pos = (unsigned long long)(p.pos_base + lrow);
#pragma unroll
                    for (int g = 0; g < NGW; g++) {
                        if (g0 + g < ng && gm[g] > 0.0f) {
                            // bit (15 - j) <=> column 16 g + j is > 0: the sign of 0 - f, one add + one funnel shift per column
                            uint32_t mask = 0;
#pragma unroll
                            for (int j = 0; j < 16; j++) mask = __funnelshift_l(__float_as_uint(0.0f - __int_as_float(v[g][j])), mask, 1);
                            mask &= 0xFFFFu;
                            while (mask) {
                                const int b = 31 - __clz((int)mask);
                                mask &= ~(1u << b);
                                const int j = 15 - b;
                                wide_append(sm->bias_s, sm->thr_s, sm->pcq_s, sm->cnt_s, lists0, p.key_lo, p.cap, pos, 16 * (g0 + g) + j,
                                            __int_as_float(sel16(v[g], j)));
                            }
                        }
                    }
                }
            }
            if (--until_check == 0 && t + 1 < ntiles) {
                until_check = p.group_tiles;
                group_sync<EPI_THREADS_>(BAR_CONSUMERS);
                const int over = (tid < nq && sm->cnt_s[tid] > limit) ? 1 : 0;
                int any;
                asm volatile(
                    "{\n"
                    ".reg .pred p, q;\n"
                    "setp.ne.s32 q, %1, 0;\n"
                    "bar.red.or.pred p, %2, %3, q;\n"
                    "selp.s32 %0, 1, 0, p;\n"
                    "}\n"
                    : "=r"(any)
                    : "r"(over), "n"(BAR_CONSUMERS), "n"(EPI_THREADS_)
                    : "memory");
                if (any) {
                    for (int qq = warp; qq < nq; qq += EPIW) {
                        const int n = sm->cnt_s[qq];
                        if (n > limit)
                            compact_list_warp(lists0 + (size_t)qq * p.cap, n, p.k, &sm->sc[warp], lane, BAR_WARP0 + warp, &sm->cnt_s[qq],
                                              &sm->tau_s[qq], p.sample_mode ? limit : 0);
                    }
                    group_sync<EPI_THREADS_>(BAR_CONSUMERS);
                    if (tid < nq) sm->thr_s[tid] = (float)(sm->pcq_s[tid] - sm->tau_s[tid]);
                    group_sync<EPI_THREADS_>(BAR_CONSUMERS);
                }
            }
        }
        group_sync<EPI_THREADS_>(BAR_CONSUMERS);
        for (int qq = warp; qq < nq && !p.sample_mode; qq += EPIW) {
            const int n = sm->cnt_s[qq];
            if (n > p.k)
                compact_list_warp(lists0 + (size_t)qq * p.cap, n, p.k, &sm->sc[warp], lane, BAR_WARP0 + warp, &sm->cnt_s[qq], &sm->tau_s[qq]);
        }
        group_sync<EPI_THREADS_>(BAR_CONSUMERS);
        if (tid < nq) p.counts[(size_t)strip * p.nq + tid] = sm->cnt_s[tid];
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

size_t wide_smem_bytes(int raw_stages, int maxq) {
    return 1024 + (size_t)raw_stages * STAGE_BYTES_RAW + (size_t)5 * maxq * 128 +
           (maxq <= 32 ? sizeof(WideSmem<FEW_MAXQ>) : (maxq <= 64 ? sizeof(WideSmem<WIDE_MAXQ>) : sizeof(WideSmem<WIDE128_MAXQ>))) + 16;
}

size_t mma_smem_bytes(int raw_stages, int cap) {
    (void)cap;  // lists are compacted in place in global memory: no shared-memory copy
    return 1024 + (size_t)raw_stages * STAGE_BYTES_RAW + (size_t)B_STAGES * STAGE_BYTES_B + sizeof(MmaSmem) + 16;
}

}  // namespace

int plan_scan_mma(vrq_ctx* ctx, int64_t rows, int nq, int k, MmaPlan* pl, bool allow_few) {
    const int sms = ctx->sm_count;
    pl->f4 = env_int("VRQ_MMA_KIND", 4) != 8;  // 4 (default): packed e2m1 operands, 8: int8 operands
    pl->qtiles = (nq + MQ - 1) / MQ;
    // <= 32 queries: the swapped-operand kernel (database rows = M) instead of 8 tensor cycles per row whatever the batch
    pl->few = allow_few && pl->f4 && nq <= FEW_MAXQ && env_int("VRQ_MMA_FEW", 1) != 0;
    // 33 .. 64 queries: the same kernel with four column groups (VRQ_MMA_MID=0: the 128-query-tile kernel)
    pl->mid = allow_few && pl->f4 && !pl->few && nq <= WIDE_MAXQ && env_int("VRQ_MMA_FEW", 1) != 0 && env_int("VRQ_MMA_MID", 1) != 0;
    // 65 .. 96 queries: the same kernel with one A buffer handed over per K half, 8 epilogue warps and up to 128 accumulator
    // columns.  Measured per 100 M codes: 65 queries 3.45 ms (128-query-tile kernel: 3.83), 96: 3.77 (3.98), 112: 4.03 (4.03),
    // 128: 4.28 (4.08) - so it takes over up to 96 (VRQ_MMA_W128 = upper limit, 0 = off)
    pl->w128 = allow_few && pl->f4 && !pl->few && !pl->mid && nq <= WIDE128_MAXQ && nq <= env_int("VRQ_MMA_W128", 96) && env_int("VRQ_MMA_FEW", 1) != 0;
    if (pl->mid || pl->w128) pl->few = true;
    // CTA pairs need an even number of query tiles (a pair = two neighbouring tiles); VRQ_MMA_PAIR=0 switches them off
    pl->pair = pl->f4 && pl->qtiles % 2 == 0 && env_int("VRQ_MMA_PAIR", 1) != 0;
    pl->group_tiles = env_int("VRQ_MMA_GROUP_TILES", 32);  // tiles between two overflow checks of the lists (measured: 8 -> 32 = +1.3 %)
    if (pl->group_tiles < 1) pl->group_tiles = 1;
    const int slack = k < 256 ? 256 : (k > 2048 ? 2048 : k);
    pl->cap = k + slack + pl->group_tiles * MROWS;
    int64_t tiles = (rows + MROWS - 1) / MROWS;
    if (tiles < 1) tiles = 1;
    pl->seg_cols = pl->seg_full = pl->seg_tail = 0;
    if (pl->pair && !pl->few && env_int("VRQ_MMA_TAIL", 1) != 0) {
        // pair scheduler: P = sms / 2 clusters, cols query-tile pairs; F full strips per column, the E clusters left over share a
        // tail strip (E must divide cols: each of them walks cols / E columns one after the other).  Equal work per cluster:
        // a full strip holds T1 = tiles * cols / (E + F * cols) tiles, the tail strip the rest.
        const int P = sms / 2, cols = pl->qtiles / 2;
        int F = P / cols, E = P % cols;
        while (E > 0 && cols % E != 0) E--;
        if (F < 1) F = 1, E = 0;
        int64_t T1 = (tiles * cols + (E + (int64_t)F * cols) - 1) / (E + (int64_t)F * cols);
        if (T1 < 1) T1 = 1;
        if (T1 * F >= tiles) {  // nothing left for a tail strip (small inputs)
            E = 0;
            F = (int)((tiles + T1 - 1) / T1);
        }
        pl->seg_cols = cols;
        pl->seg_full = F;
        pl->seg_tail = E;
        pl->rows_per_strip = T1 * MROWS;
        pl->strips = F + (E > 0 ? 1 : 0);
    } else {
        int strips = sms / pl->qtiles;
        if (strips < 1) strips = 1;
        if (strips > tiles) strips = (int)tiles;
        const int64_t tps = (tiles + strips - 1) / strips;
        pl->rows_per_strip = tps * MROWS;
        pl->strips = (int)((tiles + tps - 1) / tps);
    }
    const size_t limit = ctx->smem_optin ? ctx->smem_optin : 227 * 1024;
    pl->smem_limit = limit;
    pl->raw_stages = env_int("VRQ_MMA_RAW_STAGES", pl->few ? FEW_MAX_RAW : 4);
    if (pl->raw_stages < 1) pl->raw_stages = 1;
    if (pl->raw_stages > (pl->few ? FEW_MAX_RAW : MAX_RAW_STAGES)) pl->raw_stages = pl->few ? FEW_MAX_RAW : MAX_RAW_STAGES;
    pl->b_stages = B_STAGES;
    mma_plan_set_cap(pl, pl->cap);
    if (pl->smem > limit) {
        vrq_set_error("tensor-core Hamming top-k with k=%d does not fit the shared-memory plan", k);
        return VRQ_ERR_UNSUPPORTED;
    }
    return 0;
}

void mma_plan_set_cap(MmaPlan* pl, int cap) {
    pl->cap = cap;
    if (pl->few) {
        pl->smem = wide_smem_bytes(pl->raw_stages, pl->w128 ? WIDE128_MAXQ : (pl->mid ? WIDE_MAXQ : FEW_MAXQ));
        return;
    }
    while (pl->raw_stages > 1 && mma_smem_bytes(pl->raw_stages, cap) > pl->smem_limit) pl->raw_stages--;
    pl->smem = mma_smem_bytes(pl->raw_stages, cap);
}

int launch_scan_mma(vrq_ctx* ctx, const CUtensorMap& tmap128, const CUtensorMap& tmap64, const ScanParams& sp, const MmaPlan& pl,
                    cudaStream_t st) {
    const size_t limit = ctx->smem_optin ? ctx->smem_optin : 227 * 1024;
    if (pl.smem > limit) {
        vrq_set_error("tensor-core Hamming top-k: shared-memory plan of %zu bytes exceeds %zu", pl.smem, limit);
        return VRQ_ERR_UNSUPPORTED;
    }
    dim3 grid(pl.qtiles, pl.strips);
    if (pl.seg_cols > 0) grid = dim3(2 * (pl.seg_cols * pl.seg_full + pl.seg_tail), 1);  // 1-D grid of CTA pairs
    if (pl.few) {
        const int npad = ((sp.nq + 15) / 16) * 16;
        if (pl.w128) {
            VRQ_CUDA(cudaFuncSetAttribute(hamming_scan_mma_wide_kernel<WIDE128_MAXQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            hamming_scan_mma_wide_kernel<WIDE128_MAXQ><<<grid, WideCfg<WIDE128_MAXQ>::THREADS, pl.smem, st>>>(tmap128, sp, pl.raw_stages, npad);
        } else if (pl.mid) {
            VRQ_CUDA(cudaFuncSetAttribute(hamming_scan_mma_wide_kernel<WIDE_MAXQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            hamming_scan_mma_wide_kernel<WIDE_MAXQ><<<grid, FEW_THREADS, pl.smem, st>>>(tmap128, sp, pl.raw_stages, npad);
        } else {
            VRQ_CUDA(cudaFuncSetAttribute(hamming_scan_mma_wide_kernel<FEW_MAXQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            hamming_scan_mma_wide_kernel<FEW_MAXQ><<<grid, FEW_THREADS, pl.smem, st>>>(tmap128, sp, pl.raw_stages, npad);
        }
    } else if (pl.f4 && pl.pair) {
        // CTA pairs: clusters of 2 along x = two neighbouring query tiles of the same strip
        auto kern = sp.sample_out ? hamming_scan_mma_kernel<KIND_F4, 2, true> : hamming_scan_mma_kernel<KIND_F4, 2, false>;
        // the lean epilogue loop (VAR 4) is written for the dense scan (runs of one tile) without the tests' distance dump;
        // VRQ_MMA_VAR=0 keeps the generic loop
        int var = sp.sample_out ? 0 : env_int("VRQ_MMA_VAR", 4);
        if (var != 4 || !(sp.run_shift == 0 && sp.run_stride == MROWS && sp.dbg == nullptr)) var = 0;
        if (var == 4) kern = hamming_scan_mma_kernel<KIND_F4, 2, false, 4>;
        VRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(var == 4 ? MMA_KERNEL_THREADS + 32 : MMA_KERNEL_THREADS);
        cfg.dynamicSmemBytes = pl.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        VRQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap64, sp, pl.raw_stages));
    } else if (pl.f4) {
        auto kern = sp.sample_out ? hamming_scan_mma_kernel<KIND_F4, 1, true> : hamming_scan_mma_kernel<KIND_F4, 1, false>;
        int var = sp.sample_out ? 0 : env_int("VRQ_MMA_VAR", 4);  // the lean dense-pass epilogue, as for the CTA pairs
        if (var != 4 || !(sp.run_shift == 0 && sp.run_stride == MROWS && sp.dbg == nullptr)) var = 0;
        if (var == 4) kern = hamming_scan_mma_kernel<KIND_F4, 1, false, 4>;
        VRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        kern<<<grid, var == 4 ? MMA_KERNEL_THREADS + 32 : MMA_KERNEL_THREADS, pl.smem, st>>>(tmap128, sp, pl.raw_stages);
    } else {
        auto kern = sp.sample_out ? hamming_scan_mma_kernel<KIND_I8, 1, true> : hamming_scan_mma_kernel<KIND_I8, 1, false>;
        VRQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        kern<<<grid, MMA_KERNEL_THREADS, pl.smem, st>>>(tmap128, sp, pl.raw_stages);
    }
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vrq
