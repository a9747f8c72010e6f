// Internal declarations shared by the translation units of libvrq.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/vrq.h"

#define VRQ_SM_FALLBACK 148

void vrq_set_error(const char* fmt, ...);

#define VRQ_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            vrq_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

#define VRQ_CHECK_ARG(cond, msg)                                     \
    do {                                                             \
        if (!(cond)) {                                               \
            vrq_set_error("%s: %s", __func__, msg);                  \
            return VRQ_ERR_ARG;                                      \
        }                                                            \
    } while (0)

#define VRQ_TRY(expr)          \
    do {                       \
        int _r = (expr);       \
        if (_r != 0) return _r; \
    } while (0)

struct vrq_timed {
    int cat;
    cudaEvent_t e0, e1;
};

enum { VRQ_CAT_SCAN = 0, VRQ_CAT_ENCODE = 1, VRQ_CAT_RESCORE = 2, VRQ_CAT_MERGE = 3, VRQ_CAT_SCAN_DENSE = 4, VRQ_CAT_COUNT = 5 };

// Growable device scratch.  Several named slots so that nested users do not trample each other.
struct vrq_buf {
    void* p = nullptr;
    size_t bytes = 0;
};
enum { VRQ_WS_STAGE_IN0 = 0, VRQ_WS_STAGE_IN1, VRQ_WS_STAGE_OUT0, VRQ_WS_STAGE_OUT1, VRQ_WS_LISTS, VRQ_WS_COUNTS,
       VRQ_WS_TOPK, VRQ_WS_SEARCH_A, VRQ_WS_SEARCH_B, VRQ_WS_SEARCH_C, VRQ_WS_SEARCH_D, VRQ_WS_SEARCH_E,
       VRQ_WS_QUERY_A, VRQ_WS_QUERY_B, VRQ_WS_OUT_A, VRQ_WS_OUT_B, VRQ_WS_OUT_C, VRQ_WS_OUT_D, VRQ_WS_OUT_E,
       VRQ_WS_MISC, VRQ_WS_TAU, VRQ_WS_MERGE_A, VRQ_WS_MERGE_B, VRQ_WS_MERGE_CA, VRQ_WS_MERGE_CB, VRQ_WS_SAMPLE_KEYS, VRQ_WS_FLAG, VRQ_WS_PROGRESS, VRQ_WS_SAMPLE_D, VRQ_WS_CHUNK_KEYS, VRQ_WS_CHUNK_LO, VRQ_WS_M3_SRC, VRQ_WS_M3_R2, VRQ_WS_IP_SCORES, VRQ_WS_SHARD_PACKED, VRQ_WS_SHARD_GATHER, VRQ_WS_SLOTS };

struct vrq_ctx {
    int device = 0;
    int sm_count = VRQ_SM_FALLBACK;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // where device-pointer calls are enqueued
    cudaStream_t pipe[2] = {nullptr, nullptr};
    cudaEvent_t pipe_ev[2] = {nullptr, nullptr};
    int64_t launches = 0;
    vrq_buf ws[VRQ_WS_SLOTS];
    void* nccl_comm = nullptr;  // ncclComm_t of this GPU's rank (vrq_ctx_set_nccl), for the sharded searches of sharded.cu
    int nccl_rank = 0, nccl_world = 1;
    bool timing = false;
    std::vector<vrq_timed> timed;
    std::vector<cudaEvent_t> ev_pool;
};

int vrq_ws_get(vrq_ctx* ctx, int slot, size_t bytes, void** out);

// true = device (or managed) memory, false = host.  Null pointers are "unknown" and skipped by callers.
int vrq_is_device_ptr(const void* p, bool* is_dev);
// All non-null pointers must agree; result in *is_dev.
int vrq_space_of(const void* const* ptrs, int n, bool* is_dev);

struct vrq_timer_scope {
    vrq_ctx* ctx;
    cudaStream_t st;
    int idx = -1;
    vrq_timer_scope(vrq_ctx* c, int cat, cudaStream_t s);
    ~vrq_timer_scope();
};

static inline void vrq_count_launch(vrq_ctx* ctx, int n = 1) { ctx->launches += n; }

// ---- encode.cu --------------------------------------------------------------------------------------
enum { VRQ_CODEC_NONE = 0, VRQ_CODEC_INT8_PERDOC, VRQ_CODEC_INT8_GLOBAL, VRQ_CODEC_INT16_GLOBAL, VRQ_CODEC_INT4 };
struct vrq_encode_args {
    const float* x;
    int64_t n;
    int d;
    int codec;
    float limit_f32;  // np.float32(limit)
    float scale_f32;  // np.float32(qmax / limit)
    void* q;          // int8 / int16 / packed int4
    void* mn;         // f32 (INT8_PERDOC) or f64 (INT4), nullable
    void* mx;
    uint8_t* ubin;  // nullable
    int ge;
};
int vrq_launch_encode(vrq_ctx* ctx, const vrq_encode_args& a, cudaStream_t st);
int vrq_launch_to_binary_int(vrq_ctx* ctx, const void* x, int elem_bytes, int64_t n, int d, int ge, uint8_t* ubin,
                             cudaStream_t st);
struct vrq_dequant_args {
    int kind;  // VRQ_PAYLOAD_*
    const void* q;
    int64_t n;
    int d;
    const void* mn;
    const void* mx;
    double limit;
    float* out;
};
int vrq_launch_dequant(vrq_ctx* ctx, const vrq_dequant_args& a, cudaStream_t st);

// ---- scan.cu ----------------------------------------------------------------------------------------
// keys out: [nq, k] sorted ascending, key = (hamming << 40) | (pos_base + row); missing = ~0ull.
int vrq_hamming_topk_dev(vrq_ctx* ctx, const uint8_t* codes, int64_t n, int code_bytes, int64_t pos_base,
                         const uint8_t* q_dev, int64_t nq, int k, uint64_t* keys_out, cudaStream_t st,
                         int32_t* dbg_dist = nullptr);  // dbg_dist: tests only, [nq][n] every distance (tensor-core path)
#define VRQ_KEY_POS_BITS 40
#define VRQ_KEY_POS_MASK ((1ull << VRQ_KEY_POS_BITS) - 1ull)
#define VRQ_KEY_NONE (~0ull)
#define VRQ_MAX_K 16384   // largest k of a Hamming top-k / k * binary_oversample of the fused searches
#define VRQ_PASS_K 4096  // keys one scan pass keeps per query; larger k are produced in chunks (scan.cu)

// ---- rescore.cu -------------------------------------------------------------------------------------
int vrq_launch_rescore_binary(vrq_ctx* ctx, const uint8_t* codes, int d, const uint64_t* keys, const int64_t* pos,
                              int64_t pos_base, int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
int vrq_launch_rescore_int8cos(vrq_ctx* ctx, const int8_t* rows, int d, const uint64_t* keys, const int64_t* pos,
                               int64_t pos_base, int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
// rescore_mma.cu (d == 1024): nibble-table Phase II, tensor-core (mma.sync s8) Phase III, Phase III over regenerated rows
int vrq_launch_rescore_binary_lut(vrq_ctx* ctx, const uint8_t* codes, const uint64_t* keys, const int64_t* pos, int64_t pos_base,
                                  int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
int vrq_launch_rescore_binary_imma(vrq_ctx* ctx, const uint8_t* codes, const uint64_t* keys, const int64_t* pos, int64_t pos_base,
                                   int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
int vrq_launch_rescore_int8cos_imma(vrq_ctx* ctx, const int8_t* rows, const uint64_t* keys, const int64_t* pos, int64_t pos_base,
                                    int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
int vrq_launch_rescore_int8cos_synth(vrq_ctx* ctx, uint64_t seed, int64_t synth_row0, const uint64_t* keys, const int64_t* pos,
                                     int64_t pos_base, int64_t nq, int m, const float* qf, double* score, cudaStream_t st);
struct vrq_rescore2_args {
    int kind;  // payload kind
    const void* payload;
    const void* aux;
    double limit;
    int d;
    const uint64_t* keys;  // [nq, m]
    int64_t pos_base;
    int64_t nq;
    int m;
    const float* qf;
    float* score;  // [nq, m] float32
};
int vrq_launch_rescore_payload_dot(vrq_ctx* ctx, const vrq_rescore2_args& a, cudaStream_t st);
int vrq_launch_merge3(vrq_ctx* ctx, int world, int64_t nq, int bk, int64_t rank_stride, const uint64_t* keys, const int64_t* labels,
                      const double* sbin, const double* scos, int k, int k2, int64_t* out_labels, int32_t* out_ham,
                      double* out_sbin, double* out_scos, int32_t* out_count, cudaStream_t st);
int vrq_launch_select2(vrq_ctx* ctx, int64_t nq, int m, const uint64_t* keys, const int64_t* labels, const float* score,
                       int k, int64_t* out_labels, float* out_score, int32_t* out_count, cudaStream_t st);
int vrq_launch_keys_to_dist_labels(vrq_ctx* ctx, const uint64_t* keys, int64_t count, int64_t pos_base,
                                   const int64_t* id_map, int64_t id0, int32_t* dist, int64_t* labels, cudaStream_t st);

// ---- float_ip.cu: brute-force float32 inner-product top-k (CohereVectorDBFloat) ------------------------------------
int vrq_launch_ip_topk(vrq_ctx* ctx, const float* rows, int64_t n, int d, const float* q, int64_t nq, int k, const int64_t* id_map,
                       int64_t id0, float* out_scores, int64_t* out_labels, cudaStream_t st);

// ---- synth.cu ---------------------------------------------------------------------------------------
int vrq_launch_synth_f32(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale, float* out,
                         cudaStream_t st);
int vrq_launch_synth_codes_int8(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, uint8_t* codes,
                                int8_t* i8, cudaStream_t st);
int vrq_launch_iota_i64(vrq_ctx* ctx, int64_t* out, int64_t n, int64_t start, cudaStream_t st);
