// Row-sharded 3-phase search INSIDE the C ABI: per-shard candidates -> ncclAllGather over NVLink -> device merge, as one
// call, so a host without torch (the reference is plain Python; a maintainer binds libvrq with ctypes) can run the
// multi-GPU path through libvrq.so alone.  The reference has no distributed path; this is the B200
// scaling of CohereEnhancedVectorDB.search (:227-322), identical in result to the single-index search3 over the
// concatenated shards (scores are pure functions of (query, document), so computing them before the exchange is exact).
//
// NCCL is loaded with dlopen at first use (libnccl.so.2: the process's already-loaded copy when torch is imported, the
// system library otherwise) - libvrq.so itself keeps linking against libcudart only.  Two ways to use it:
//   * one process per GPU (torchrun-style): vrq_nccl_unique_id on rank 0, ship the 128 bytes to the other ranks by any
//     means, vrq_nccl_init_rank everywhere, vrq_ctx_set_nccl, then vrq_index_search3_sharded on every rank;
//   * one process driving all GPUs of the box: vrq_nccl_init_all, vrq_ctx_set_nccl per context, then
//     vrq_search3_sharded_group with the per-GPU index handles (the collective is issued for every rank inside one
//     ncclGroupStart / ncclGroupEnd).
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vrq_internal.cuh"

namespace {

typedef int (*PFN_AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*PFN_CommInitAll)(void**, int, const int*);
typedef int (*PFN_GetUniqueId)(void*);
typedef struct {
    char internal[128];
} NcclId;
typedef int (*PFN_CommInitRank)(void**, int, NcclId, int);
typedef int (*PFN_Void)(void);
typedef int (*PFN_CommDestroy)(void*);
typedef const char* (*PFN_ErrStr)(int);

struct Nccl {
    void* so = nullptr;
    PFN_AllGather all_gather = nullptr;
    PFN_CommInitAll init_all = nullptr;
    PFN_GetUniqueId unique_id = nullptr;
    PFN_CommInitRank init_rank = nullptr;
    PFN_Void group_start = nullptr, group_end = nullptr;
    PFN_CommDestroy destroy = nullptr;
    PFN_ErrStr err = nullptr;
};
Nccl g_nccl;
constexpr int NCCL_INT64 = 4;  // ncclDataType_t

int load_nccl() {
    if (g_nccl.so) return 0;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) {
        vrq_set_error("NCCL is not available: %s", dlerror());
        return VRQ_ERR_UNSUPPORTED;
    }
    Nccl n;
    n.so = so;
    n.all_gather = (PFN_AllGather)dlsym(so, "ncclAllGather");
    n.init_all = (PFN_CommInitAll)dlsym(so, "ncclCommInitAll");
    n.unique_id = (PFN_GetUniqueId)dlsym(so, "ncclGetUniqueId");
    n.init_rank = (PFN_CommInitRank)dlsym(so, "ncclCommInitRank");
    n.group_start = (PFN_Void)dlsym(so, "ncclGroupStart");
    n.group_end = (PFN_Void)dlsym(so, "ncclGroupEnd");
    n.destroy = (PFN_CommDestroy)dlsym(so, "ncclCommDestroy");
    n.err = (PFN_ErrStr)dlsym(so, "ncclGetErrorString");
    if (!n.all_gather || !n.init_all || !n.unique_id || !n.init_rank || !n.group_start || !n.group_end || !n.destroy) {
        vrq_set_error("libnccl.so.2 lacks an expected entry point");
        return VRQ_ERR_UNSUPPORTED;
    }
    g_nccl = n;
    return 0;
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return 0;
    vrq_set_error("%s failed: %s", what, g_nccl.err ? g_nccl.err(rc) : "NCCL error");
    return VRQ_ERR_STATE;
}

struct ShardCall {
    vrq_index* ix;
    vrq_ctx* ctx;
    int64_t nq;
    int k, bk, k2;
    void *packed, *gathered;
};

}  // namespace

// index.cu
int vrq_index_ctx(vrq_index* ix, vrq_ctx** out);

extern "C" int vrq_nccl_unique_id(char* id128) {
    VRQ_CHECK_ARG(id128 != nullptr, "null argument");
    VRQ_TRY(load_nccl());
    NcclId id;
    VRQ_TRY(nccl_check(g_nccl.unique_id(&id), "ncclGetUniqueId"));
    memcpy(id128, id.internal, 128);
    return 0;
}

extern "C" int vrq_nccl_init_rank(int device, int world, const char* id128, int rank, void** comm_out) {
    VRQ_CHECK_ARG(id128 && comm_out && world > 0 && rank >= 0 && rank < world, "bad argument");
    VRQ_TRY(load_nccl());
    VRQ_CUDA(cudaSetDevice(device));
    NcclId id;
    memcpy(id.internal, id128, 128);
    return nccl_check(g_nccl.init_rank(comm_out, world, id, rank), "ncclCommInitRank");
}

extern "C" int vrq_nccl_init_all(int ndev, const int* devices, void** comms_out) {
    VRQ_CHECK_ARG(ndev > 0 && comms_out, "bad argument");
    VRQ_TRY(load_nccl());
    return nccl_check(g_nccl.init_all(comms_out, ndev, devices), "ncclCommInitAll");
}

extern "C" int vrq_nccl_destroy(void* comm) {
    if (!comm) return 0;
    VRQ_TRY(load_nccl());
    return nccl_check(g_nccl.destroy(comm), "ncclCommDestroy");
}

extern "C" int vrq_ctx_set_nccl(vrq_ctx* ctx, void* nccl_comm, int rank, int world) {
    VRQ_CHECK_ARG(ctx != nullptr, "ctx is null");
    VRQ_CHECK_ARG((nccl_comm == nullptr) || (world > 0 && rank >= 0 && rank < world), "bad rank / world");
    ctx->nccl_comm = nccl_comm;
    ctx->nccl_rank = nccl_comm ? rank : 0;
    ctx->nccl_world = nccl_comm ? world : 1;
    return 0;
}

// The three steps of one rank, split so that the group form can interleave them over the ranks of one process.
static int shard_local(ShardCall& c, const float* q_float, const uint8_t* q_ubin, int64_t pos_base) {
    VRQ_CUDA(cudaSetDevice(c.ctx->device));
    const int world = c.ctx->nccl_world;
    const size_t block = (size_t)4 * c.nq * c.bk;  // int64 elements per rank: keys, labels, score_binary, score_cosine
    VRQ_TRY(vrq_ws_get(c.ctx, VRQ_WS_SHARD_PACKED, 8 * block, &c.packed));
    VRQ_TRY(vrq_ws_get(c.ctx, VRQ_WS_SHARD_GATHER, 8 * block * (size_t)world, &c.gathered));
    uint64_t* pk = (uint64_t*)c.packed;
    const size_t cnt = (size_t)c.nq * c.bk;
    return vrq_index_search3_local(c.ix, c.nq, q_float, q_ubin, c.bk, pos_base, pk, (int64_t*)(pk + cnt), (double*)(pk + 2 * cnt),
                                   (double*)(pk + 3 * cnt));
}
static int shard_gather(ShardCall& c) {
    const size_t block = (size_t)4 * c.nq * c.bk;
    if (c.ctx->nccl_world == 1) {
        VRQ_CUDA(cudaSetDevice(c.ctx->device));
        VRQ_CUDA(cudaMemcpyAsync(c.gathered, c.packed, 8 * block, cudaMemcpyDeviceToDevice, c.ctx->stream));
        return 0;
    }
    return nccl_check(g_nccl.all_gather(c.packed, c.gathered, block, NCCL_INT64, c.ctx->nccl_comm, c.ctx->stream), "ncclAllGather");
}
static int shard_merge(ShardCall& c, int64_t* labels, int32_t* hamming, double* sb, double* sc, int32_t* count) {
    VRQ_CUDA(cudaSetDevice(c.ctx->device));
    const uint64_t* g = (const uint64_t*)c.gathered;
    const size_t cnt = (size_t)c.nq * c.bk;
    return vrq_launch_merge3(c.ctx, c.ctx->nccl_world, c.nq, c.bk, (int64_t)(4 * cnt), g, (const int64_t*)(g + cnt), (const double*)(g + 2 * cnt),
                             (const double*)(g + 3 * cnt), c.k, c.k2, labels, hamming, sb, sc, count, c.ctx->stream);
}

static int shard_prepare(ShardCall& c, vrq_index* ix, int64_t nq, int k, int bo, int io, int64_t ntotal_global) {
    VRQ_CHECK_ARG(ix != nullptr && nq > 0 && k > 0 && bo > 0 && io > 0 && ntotal_global > 0, "bad argument");
    c.ix = ix;
    VRQ_TRY(vrq_index_ctx(ix, &c.ctx));
    if (c.ctx->nccl_world > 1 && !c.ctx->nccl_comm) {
        vrq_set_error("no NCCL communicator on this context (vrq_ctx_set_nccl)");
        return VRQ_ERR_STATE;
    }
    const int64_t bk64 = std::min<int64_t>((int64_t)k * bo, ntotal_global);  // binary_k = min(k * oversample, GLOBAL ntotal) (:267)
    if (bk64 > VRQ_MAX_K) {
        vrq_set_error("k * binary_oversample = %lld exceeds the supported %d", (long long)bk64, VRQ_MAX_K);
        return VRQ_ERR_UNSUPPORTED;
    }
    c.nq = nq;
    c.k = k;
    c.bk = (int)bk64;
    c.k2 = k * io;
    return 0;
}

// One rank of a row-sharded search (every rank of the communicator must call it with the same queries).  All pointers are
// DEVICE pointers on this rank's GPU; the work is enqueued on the context's stream.  pos_base = global position of this
// shard's row 0; ntotal_global = rows over all shards.
extern "C" int vrq_index_search3_sharded(vrq_index* ix, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k, int binary_oversample,
                                         int int8_oversample, int64_t pos_base, int64_t ntotal_global, int64_t* labels, int32_t* hamming,
                                         double* score_binary, double* score_cosine, int32_t* out_count) {
    VRQ_CHECK_ARG(q_float && q_ubin && labels && hamming && score_binary && score_cosine && out_count, "null argument");
    ShardCall c{};
    VRQ_TRY(shard_prepare(c, ix, nq, k, binary_oversample, int8_oversample, ntotal_global));
    if (c.ctx->nccl_world > 1) VRQ_TRY(load_nccl());
    VRQ_TRY(shard_local(c, q_float, q_ubin, pos_base));
    VRQ_TRY(shard_gather(c));
    return shard_merge(c, labels, hamming, score_binary, score_cosine, out_count);
}

// All ranks from ONE process: ixs[r] is the shard on GPU r (its context carries the communicator of rank r), q_float[r] /
// q_ubin[r] / outputs[r] are device pointers on that GPU.  pos_base[r] as above.
extern "C" int vrq_search3_sharded_group(int world, vrq_index* const* ixs, int64_t nq, const float* const* q_float, const uint8_t* const* q_ubin,
                                         int k, int binary_oversample, int int8_oversample, const int64_t* pos_base, int64_t ntotal_global,
                                         int64_t* const* labels, int32_t* const* hamming, double* const* score_binary,
                                         double* const* score_cosine, int32_t* const* out_count) {
    VRQ_CHECK_ARG(world > 0 && ixs && q_float && q_ubin && pos_base && labels && hamming && score_binary && score_cosine && out_count,
                  "null argument");
    std::vector<ShardCall> calls((size_t)world);
    for (int r = 0; r < world; r++) {
        VRQ_TRY(shard_prepare(calls[r], ixs[r], nq, k, binary_oversample, int8_oversample, ntotal_global));
        if (calls[r].ctx->nccl_world != world || calls[r].ctx->nccl_rank != r) {
            vrq_set_error("context of shard %d is not rank %d of a %d-rank communicator", r, r, world);
            return VRQ_ERR_STATE;
        }
    }
    if (world > 1) VRQ_TRY(load_nccl());
    for (int r = 0; r < world; r++) VRQ_TRY(shard_local(calls[r], q_float[r], q_ubin[r], pos_base[r]));
    if (world > 1) VRQ_TRY(nccl_check(g_nccl.group_start(), "ncclGroupStart"));
    for (int r = 0; r < world; r++) {
        const int rc = shard_gather(calls[r]);
        if (rc != 0) {
            if (world > 1) g_nccl.group_end();
            return rc;
        }
    }
    if (world > 1) VRQ_TRY(nccl_check(g_nccl.group_end(), "ncclGroupEnd"));
    for (int r = 0; r < world; r++) VRQ_TRY(shard_merge(calls[r], labels[r], hamming[r], score_binary[r], score_cosine[r], out_count[r]));
    return 0;
}
