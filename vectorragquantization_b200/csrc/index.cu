// Device-resident binary index: the faiss.IndexBinaryIDMap2(faiss.IndexBinaryFlat(d)) surface the reference
// classes call, plus an optional per-position payload matrix that replaces the RocksDB
// point-gets inside the reference's rescoring loops, and the fused multi-phase searches.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <iterator>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vrq_internal.cuh"

struct vrq_index {
    vrq_ctx* ctx = nullptr;
    int d = 0;
    int code_bytes = 0;
    int64_t ntotal = 0, capacity = 0;
    uint8_t* codes = nullptr;
    int64_t* ids = nullptr;  // null while ids are implicit (id = id0 + position)
    bool implicit_ids = true;
    int64_t id0 = 0;
    int payload_kind = VRQ_PAYLOAD_NONE;
    double limit = 0.0;
    size_t payload_row = 0, aux_row = 0;
    uint8_t* payload = nullptr;
    uint8_t* aux = nullptr;
    std::vector<int64_t> host_ids;
    bool host_ids_valid = false;
    std::unordered_map<int64_t, int64_t> rev;  // id -> last position
    bool rev_valid = false;
    // remove_ids is lazy: removed positions (sorted, unique, in the coordinates of the arrays as they are) wait here and
    // the arrays are compacted once, in place, before the next call that reads them.  `ntotal` counts the rows in the
    // arrays, ntotal - dead.size() is what the caller sees.
    std::vector<int64_t> dead;
    // Benchmark-only "virtual" INT8_RAW payload: no rows in HBM, Phase III regenerates row p from the counter-based
    // generator (seed, synth_row0 + p).  For the 1-billion-row legs on fewer than 8 GPUs.
    bool synth_payload = false;
    uint64_t synth_seed = 0;
    int64_t synth_row0 = 0;
};

namespace {

size_t payload_row_bytes(int kind, int d) {
    switch (kind) {
        case VRQ_PAYLOAD_INT8_RAW:
        case VRQ_PAYLOAD_INT8_PERDOC:
        case VRQ_PAYLOAD_INT8_GLOBAL:
            return (size_t)d;
        case VRQ_PAYLOAD_INT16_GLOBAL:
            return (size_t)d * 2;
        case VRQ_PAYLOAD_INT4_PERDOC:
        case VRQ_PAYLOAD_INT4_GLOBAL:
            return (size_t)d / 2;
        case VRQ_PAYLOAD_F32:
            return (size_t)d * 4;
    }
    return 0;
}
size_t aux_row_bytes(int kind) {
    if (kind == VRQ_PAYLOAD_INT8_PERDOC) return 2 * sizeof(float);
    if (kind == VRQ_PAYLOAD_INT4_PERDOC) return 2 * sizeof(double);
    return 0;
}

int grow_one(vrq_index* ix, void** buf, size_t row_bytes, int64_t new_cap) {
    if (row_bytes == 0) return 0;
    void* nb = nullptr;
    cudaError_t e = cudaMalloc(&nb, row_bytes * (size_t)new_cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        vrq_set_error("cudaMalloc of %zu bytes for the index failed: %s", row_bytes * (size_t)new_cap, cudaGetErrorString(e));
        return VRQ_ERR_NOMEM;
    }
    if (*buf && ix->ntotal > 0)
        VRQ_CUDA(cudaMemcpyAsync(nb, *buf, row_bytes * (size_t)ix->ntotal, cudaMemcpyDeviceToDevice, ix->ctx->stream));
    if (*buf) {
        VRQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));
        VRQ_CUDA(cudaFree(*buf));
    }
    *buf = nb;
    return 0;
}

int ensure_capacity(vrq_index* ix, int64_t want) {
    if (want <= ix->capacity) return 0;
    int64_t nc = std::max<int64_t>(want, std::max<int64_t>(1024, ix->capacity + ix->capacity / 2));
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    VRQ_TRY(grow_one(ix, (void**)&ix->codes, (size_t)ix->code_bytes, nc));
    if (!ix->implicit_ids) VRQ_TRY(grow_one(ix, (void**)&ix->ids, sizeof(int64_t), nc));
    VRQ_TRY(grow_one(ix, (void**)&ix->payload, ix->payload_row, nc));
    VRQ_TRY(grow_one(ix, (void**)&ix->aux, ix->aux_row, nc));
    ix->capacity = nc;
    return 0;
}

int materialise_ids(vrq_index* ix) {
    if (!ix->implicit_ids) return 0;
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    if (ix->capacity > 0) {
        VRQ_CUDA(cudaMalloc((void**)&ix->ids, sizeof(int64_t) * (size_t)ix->capacity));
        VRQ_TRY(vrq_launch_iota_i64(ix->ctx, ix->ids, ix->ntotal, ix->id0, ix->ctx->stream));
    }
    ix->implicit_ids = false;
    return 0;
}

int load_host_ids(vrq_index* ix) {
    if (ix->host_ids_valid) return 0;
    try {
        ix->host_ids.resize((size_t)ix->ntotal);
    } catch (const std::bad_alloc&) {  // never let a C++ exception cross the extern "C" boundary
        vrq_set_error("out of host memory for %lld ids", (long long)ix->ntotal);
        return VRQ_ERR_NOMEM;
    }
    if (ix->implicit_ids) {
        for (int64_t i = 0; i < ix->ntotal; i++) ix->host_ids[(size_t)i] = ix->id0 + i;
    } else if (ix->ntotal > 0) {
        VRQ_CUDA(cudaMemcpyAsync(ix->host_ids.data(), ix->ids, sizeof(int64_t) * (size_t)ix->ntotal, cudaMemcpyDeviceToHost,
                                 ix->ctx->stream));
        VRQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));
    }
    ix->host_ids_valid = true;
    return 0;
}

int build_rev(vrq_index* ix) {
    if (ix->rev_valid) return 0;
    VRQ_TRY(load_host_ids(ix));
    try {
        ix->rev.clear();
        ix->rev.reserve((size_t)ix->ntotal * 2);
        for (int64_t i = 0; i < ix->ntotal; i++) ix->rev[ix->host_ids[(size_t)i]] = i;  // last added wins (IDMap2)
    } catch (const std::bad_alloc&) {
        vrq_set_error("out of host memory for the id -> position map of %lld ids", (long long)ix->ntotal);
        return VRQ_ERR_NOMEM;
    }
    ix->rev_valid = true;
    return 0;
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ positions, int64_t m,
                                   int row_bytes, uint8_t* __restrict__ dst) {
    // one warp per row, byte granularity in 4-byte words when possible
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < m; r += (int64_t)gridDim.x * 8) {
        const int64_t p = positions[r];
        const uint8_t* s = src + (size_t)p * row_bytes;
        uint8_t* o = dst + (size_t)r * row_bytes;
        if ((row_bytes & 3) == 0) {
            for (int w = lane; w < row_bytes / 4; w += 32)
                reinterpret_cast<uint32_t*>(o)[w] = reinterpret_cast<const uint32_t*>(s)[w];
        } else {
            for (int b = lane; b < row_bytes; b += 32) o[b] = s[b];
        }
    }
}

int gather_rows(vrq_ctx* ctx, const uint8_t* src, const int64_t* pos_dev, int64_t m, size_t row_bytes, uint8_t* dst,
                cudaStream_t st) {
    if (m == 0 || row_bytes == 0) return 0;
    int64_t blocks = (m + 7) / 8, cap = (int64_t)ctx->sm_count * 16;
    gather_rows_kernel<<<(unsigned)std::min(blocks, cap), 256, 0, st>>>(src, pos_dev, m, (int)row_bytes, dst);
    vrq_count_launch(ctx);
    VRQ_CUDA(cudaGetLastError());
    return 0;
}

// dst row (i - #dead below i) of the compacted array <- row i, for the kept rows i of [a, b); written to a bounce buffer
// whose row 0 is compacted row `out` (an in-place forward shift cannot be done by an unordered grid).
__global__ void compact_chunk_kernel(const uint8_t* __restrict__ src, int64_t a, int64_t b, const int64_t* __restrict__ dead, int64_t m,
                                     int row_bytes, int64_t out, uint8_t* __restrict__ bounce) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = a + (int64_t)blockIdx.x * 8 + warp; i < b; i += (int64_t)gridDim.x * 8) {
        int64_t lo = 0, hi = m;  // lower_bound(dead, i)
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (dead[mid] < i) lo = mid + 1; else hi = mid;
        }
        if (lo < m && dead[lo] == i) continue;
        const uint8_t* s = src + (size_t)i * row_bytes;
        uint8_t* o = bounce + (size_t)(i - lo - out) * row_bytes;
        if ((row_bytes & 15) == 0) {
            for (int w = lane; w < row_bytes / 16; w += 32) reinterpret_cast<uint4*>(o)[w] = reinterpret_cast<const uint4*>(s)[w];
        } else if ((row_bytes & 3) == 0) {
            for (int w = lane; w < row_bytes / 4; w += 32) reinterpret_cast<uint32_t*>(o)[w] = reinterpret_cast<const uint32_t*>(s)[w];
        } else {
            for (int w = lane; w < row_bytes; w += 32) o[w] = s[w];
        }
    }
}

// Apply the pending removals: order-preserving, in place (faiss shifts the tail down; survivors keep their relative
// order, so the (distance, position) tie order of every later search is the one faiss would give).  Rows before the first
// removed position do not move; the rest goes through a 64 MB bounce buffer chunk by chunk in ascending order, so no
// second copy of a 100 GB payload is ever needed.
int flush_dead(vrq_index* ix) {
    if (ix->dead.empty()) return 0;
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    VRQ_TRY(materialise_ids(ix));
    const int64_t m = (int64_t)ix->dead.size();
    void* dead_dev = nullptr;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_B, sizeof(int64_t) * (size_t)m, &dead_dev));
    VRQ_CUDA(cudaMemcpyAsync(dead_dev, ix->dead.data(), sizeof(int64_t) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    auto compact = [&](uint8_t* buf, size_t row) -> int {
        if (row == 0 || !buf) return 0;
        const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)((64u << 20) / row));
        void* bounce = nullptr;
        VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_STAGE_OUT0, (size_t)chunk_rows * row, &bounce));
        for (int64_t a = ix->dead[0]; a < ix->ntotal; a += chunk_rows) {
            const int64_t b = std::min(ix->ntotal, a + chunk_rows);
            const int64_t dead_below_a = std::lower_bound(ix->dead.begin(), ix->dead.end(), a) - ix->dead.begin();
            const int64_t dead_below_b = std::lower_bound(ix->dead.begin(), ix->dead.end(), b) - ix->dead.begin();
            const int64_t out = a - dead_below_a, kept = (b - a) - (dead_below_b - dead_below_a);
            if (kept == 0) continue;
            const int64_t blocks = std::min<int64_t>((b - a + 7) / 8, (int64_t)ctx->sm_count * 16);
            compact_chunk_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(buf, a, b, (const int64_t*)dead_dev, m, (int)row, out,
                                                                            (uint8_t*)bounce);
            vrq_count_launch(ctx);
            VRQ_CUDA(cudaGetLastError());
            VRQ_CUDA(cudaMemcpyAsync(buf + (size_t)out * row, bounce, (size_t)kept * row, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return 0;
    };
    VRQ_TRY(compact(ix->codes, (size_t)ix->code_bytes));
    VRQ_TRY(compact((uint8_t*)ix->ids, sizeof(int64_t)));
    VRQ_TRY(compact(ix->payload, ix->payload_row));
    VRQ_TRY(compact(ix->aux, ix->aux_row));
    VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
    ix->ntotal -= m;
    ix->dead.clear();
    ix->host_ids_valid = false;
    ix->rev_valid = false;
    return 0;
}

struct DevIO {
    // Stages host arguments of one index call into scratch, remembers host outputs to copy back.
    vrq_ctx* ctx;
    bool host;
    std::vector<std::pair<void*, std::pair<void*, size_t>>> outs;
    int in(const void* p, size_t bytes, int slot, const void** dev) {
        if (!host || !p) {
            *dev = p;
            return 0;
        }
        void* d;
        VRQ_TRY(vrq_ws_get(ctx, slot, bytes, &d));
        VRQ_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        *dev = d;
        return 0;
    }
    int out(void* p, size_t bytes, int slot, void** dev) {
        if (!host || !p) {
            *dev = p;
            return 0;
        }
        void* d;
        VRQ_TRY(vrq_ws_get(ctx, slot, bytes, &d));
        outs.push_back({p, {d, bytes}});
        *dev = d;
        return 0;
    }
    int finish() {
        if (!host) return 0;
        for (auto& o : outs)
            VRQ_CUDA(cudaMemcpyAsync(o.first, o.second.first, o.second.second, cudaMemcpyDeviceToHost, ctx->stream));
        VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
};

}  // namespace

extern "C" int vrq_index_create(vrq_ctx* ctx, int d, vrq_index** out) {
    VRQ_CHECK_ARG(ctx != nullptr && out != nullptr, "null argument");
    VRQ_CHECK_ARG(d > 0 && d % 8 == 0, "d must be a positive multiple of 8 (faiss binary indexes require it)");
    vrq_index* ix = new vrq_index();
    ix->ctx = ctx;
    ix->d = d;
    ix->code_bytes = d / 8;
    *out = ix;
    return 0;
}

int vrq_index_ctx(vrq_index* ix, vrq_ctx** out) {
    VRQ_CHECK_ARG(ix != nullptr && out != nullptr, "null argument");
    *out = ix->ctx;
    return 0;
}

extern "C" int vrq_index_free(vrq_index* ix) {
    if (!ix) return 0;
    cudaSetDevice(ix->ctx->device);
    cudaStreamSynchronize(ix->ctx->stream);
    if (ix->codes) cudaFree(ix->codes);
    if (ix->ids) cudaFree(ix->ids);
    if (ix->payload) cudaFree(ix->payload);
    if (ix->aux) cudaFree(ix->aux);
    delete ix;
    return 0;
}

extern "C" int64_t vrq_index_ntotal(const vrq_index* ix) { return ix ? ix->ntotal - (int64_t)ix->dead.size() : 0; }
extern "C" int vrq_index_d(const vrq_index* ix) { return ix ? ix->d : 0; }
extern "C" int vrq_index_payload_kind(const vrq_index* ix) { return ix ? ix->payload_kind : 0; }

extern "C" int vrq_index_device_ptrs(vrq_index* ix, void** codes, void** ids, void** payload, void** aux) {
    VRQ_CHECK_ARG(ix != nullptr, "index is null");
    VRQ_TRY(flush_dead(ix));
    if (codes) *codes = ix->codes;
    if (ids) *ids = ix->ids;
    if (payload) *payload = ix->payload;
    if (aux) *aux = ix->aux;
    return 0;
}

extern "C" int vrq_index_reserve(vrq_index* ix, int64_t cap) {
    VRQ_CHECK_ARG(ix != nullptr && cap >= 0, "bad argument");
    if (cap <= ix->capacity) return 0;
    // exact reservation (no geometric slack): 100 M x (128 + 1024) B is most of the HBM
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    VRQ_TRY(grow_one(ix, (void**)&ix->codes, (size_t)ix->code_bytes, cap));
    if (!ix->implicit_ids) VRQ_TRY(grow_one(ix, (void**)&ix->ids, sizeof(int64_t), cap));
    VRQ_TRY(grow_one(ix, (void**)&ix->payload, ix->payload_row, cap));
    VRQ_TRY(grow_one(ix, (void**)&ix->aux, ix->aux_row, cap));
    ix->capacity = cap;
    return 0;
}

extern "C" int vrq_index_set_payload(vrq_index* ix, int kind, double global_limit) {
    VRQ_CHECK_ARG(ix != nullptr, "index is null");
    VRQ_CHECK_ARG(kind >= VRQ_PAYLOAD_NONE && kind <= VRQ_PAYLOAD_CODES_PM1, "unknown payload kind");
    if (ix->ntotal != 0 || ix->capacity != 0) {
        vrq_set_error("payload kind can only be set on an empty, unreserved index");
        return VRQ_ERR_STATE;
    }
    ix->payload_kind = kind;
    ix->limit = global_limit;
    ix->payload_row = payload_row_bytes(kind, ix->d);
    ix->aux_row = aux_row_bytes(kind);
    return 0;
}

extern "C" int vrq_index_set_synthetic_payload(vrq_index* ix, uint64_t seed, int64_t row0) {
    VRQ_CHECK_ARG(ix != nullptr, "index is null");
    if (ix->ntotal != 0 || ix->capacity != 0 || ix->payload_kind != VRQ_PAYLOAD_INT8_RAW) {
        vrq_set_error("a synthetic payload can only replace the INT8_RAW payload of an empty, unreserved index");
        return VRQ_ERR_STATE;
    }
    ix->synth_payload = true;
    ix->synth_seed = seed;
    ix->synth_row0 = row0;
    ix->payload_row = 0;  // nothing is stored
    return 0;
}

extern "C" int vrq_index_add_with_ids(vrq_index* ix, int64_t n, const uint8_t* codes, const int64_t* ids, const void* payload,
                                      const void* aux) {
    VRQ_CHECK_ARG(ix != nullptr && n >= 0, "bad argument");
    if (ix && ix->synth_payload) {
        vrq_set_error("an index with a synthetic payload only accepts vrq_index_add_synthetic");
        return VRQ_ERR_STATE;
    }
    if (ix) VRQ_TRY(flush_dead(ix));
    if (n == 0) return 0;
    VRQ_CHECK_ARG(ids != nullptr, "ids are null");
    // codes may be NULL only for a float32-payload index that is searched by inner product alone (CohereVectorDBFloat): zero codes
    VRQ_CHECK_ARG(codes != nullptr || ix->payload_kind == VRQ_PAYLOAD_F32, "codes are null");
    VRQ_CHECK_ARG((ix->payload_row == 0) == (payload == nullptr), "payload must be given exactly when a payload kind is set");
    VRQ_CHECK_ARG((ix->aux_row == 0) == (aux == nullptr), "aux (min,max pairs) must be given exactly for the per-document kinds");
    const void* all[4] = {codes, ids, payload, aux};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 4, &is_dev));
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    VRQ_TRY(materialise_ids(ix));
    VRQ_TRY(ensure_capacity(ix, ix->ntotal + n));
    if (!ix->ids) VRQ_CUDA(cudaMalloc((void**)&ix->ids, sizeof(int64_t) * (size_t)ix->capacity));
    cudaStream_t st = ix->ctx->stream;
    const cudaMemcpyKind kind = is_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (codes)
        VRQ_CUDA(cudaMemcpyAsync(ix->codes + (size_t)ix->ntotal * ix->code_bytes, codes, (size_t)n * ix->code_bytes, kind, st));
    else
        VRQ_CUDA(cudaMemsetAsync(ix->codes + (size_t)ix->ntotal * ix->code_bytes, 0, (size_t)n * ix->code_bytes, st));
    VRQ_CUDA(cudaMemcpyAsync(ix->ids + ix->ntotal, ids, sizeof(int64_t) * (size_t)n, kind, st));
    if (payload) VRQ_CUDA(cudaMemcpyAsync(ix->payload + (size_t)ix->ntotal * ix->payload_row, payload, (size_t)n * ix->payload_row, kind, st));
    if (aux) VRQ_CUDA(cudaMemcpyAsync(ix->aux + (size_t)ix->ntotal * ix->aux_row, aux, (size_t)n * ix->aux_row, kind, st));
    if (!is_dev) VRQ_CUDA(cudaStreamSynchronize(st));
    ix->ntotal += n;
    ix->host_ids_valid = false;
    ix->rev_valid = false;
    return 0;
}

extern "C" int vrq_index_add_synthetic(vrq_index* ix, uint64_t seed, int64_t row0, int64_t nrows, int64_t id0) {
    VRQ_CHECK_ARG(ix != nullptr && nrows >= 0, "bad argument");
    if (ix) VRQ_TRY(flush_dead(ix));
    if (ix->payload_kind != VRQ_PAYLOAD_NONE && ix->payload_kind != VRQ_PAYLOAD_INT8_RAW) {
        vrq_set_error("add_synthetic fills codes (+ INT8_RAW payload) only");
        return VRQ_ERR_STATE;
    }
    if (nrows == 0) return 0;
    if (ix->synth_payload && (seed != ix->synth_seed || row0 != ix->synth_row0 + ix->ntotal)) {
        vrq_set_error("add_synthetic: rows must continue the (seed, row0) sequence the synthetic payload was declared with");
        return VRQ_ERR_ARG;
    }
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    const bool contiguous = ix->implicit_ids && (ix->ntotal == 0 || ix->id0 + ix->ntotal == id0);
    if (ix->ntotal == 0 && ix->implicit_ids) ix->id0 = id0;
    if (!contiguous) VRQ_TRY(materialise_ids(ix));
    VRQ_TRY(ensure_capacity(ix, ix->ntotal + nrows));
    cudaStream_t st = ix->ctx->stream;
    if (!ix->implicit_ids) {
        if (!ix->ids) VRQ_CUDA(cudaMalloc((void**)&ix->ids, sizeof(int64_t) * (size_t)ix->capacity));
        VRQ_TRY(vrq_launch_iota_i64(ix->ctx, ix->ids + ix->ntotal, nrows, id0, st));
    }
    VRQ_TRY(vrq_launch_synth_codes_int8(ix->ctx, seed, row0, nrows, ix->d, ix->codes + (size_t)ix->ntotal * ix->code_bytes,
                                        ix->payload ? (int8_t*)(ix->payload + (size_t)ix->ntotal * ix->payload_row) : nullptr, st));
    ix->ntotal += nrows;
    ix->host_ids_valid = false;
    ix->rev_valid = false;
    return 0;
}

extern "C" int vrq_index_search(vrq_index* ix, int64_t nq, const uint8_t* q, int k, int32_t* dist, int64_t* labels) {
    VRQ_CHECK_ARG(ix != nullptr && nq >= 0, "bad argument");
    VRQ_TRY(flush_dead(ix));
    if (nq == 0) return 0;
    VRQ_CHECK_ARG(q != nullptr && dist != nullptr && labels != nullptr, "null pointer");
    VRQ_CHECK_ARG(k > 0, "k must be > 0");
    const void* all[3] = {q, dist, labels};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 3, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    DevIO io{ctx, !is_dev, {}};
    const void* dq;
    void *dd, *dl, *keys;
    VRQ_TRY(io.in(q, (size_t)nq * ix->code_bytes, VRQ_WS_QUERY_A, &dq));
    VRQ_TRY(io.out(dist, sizeof(int32_t) * (size_t)nq * k, VRQ_WS_OUT_A, &dd));
    VRQ_TRY(io.out(labels, sizeof(int64_t) * (size_t)nq * k, VRQ_WS_OUT_B, &dl));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_TOPK, sizeof(uint64_t) * (size_t)nq * k, &keys));
    VRQ_TRY(vrq_hamming_topk_dev(ctx, ix->codes, ix->ntotal, ix->code_bytes, 0, (const uint8_t*)dq, nq, k, (uint64_t*)keys, ctx->stream));
    VRQ_TRY(vrq_launch_keys_to_dist_labels(ctx, (const uint64_t*)keys, nq * (int64_t)k, 0, ix->implicit_ids ? nullptr : ix->ids,
                                           ix->id0, (int32_t*)dd, (int64_t*)dl, ctx->stream));
    return io.finish();
}

extern "C" int vrq_index_distances(vrq_index* ix, int64_t nq, const uint8_t* q, int32_t* dist) {
    VRQ_CHECK_ARG(ix != nullptr && nq >= 0, "bad argument");
    VRQ_TRY(flush_dead(ix));
    if (nq == 0 || ix->ntotal == 0) return 0;
    VRQ_CHECK_ARG(q != nullptr && dist != nullptr, "null pointer");
    const void* all[2] = {q, dist};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 2, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    DevIO io{ctx, !is_dev, {}};
    const void* dq;
    void *dd, *keys;
    VRQ_TRY(io.in(q, (size_t)nq * ix->code_bytes, VRQ_WS_QUERY_A, &dq));
    VRQ_TRY(io.out(dist, sizeof(int32_t) * (size_t)nq * ix->ntotal, VRQ_WS_OUT_A, &dd));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_TOPK, sizeof(uint64_t) * (size_t)nq, &keys));
    VRQ_TRY(vrq_hamming_topk_dev(ctx, ix->codes, ix->ntotal, ix->code_bytes, 0, (const uint8_t*)dq, nq, 1, (uint64_t*)keys, ctx->stream,
                                 (int32_t*)dd));
    return io.finish();
}

extern "C" int64_t vrq_index_position_of(vrq_index* ix, int64_t id) {
    if (!ix) return -1;
    if (flush_dead(ix) != 0) return -1;
    if (ix->implicit_ids) return (id >= ix->id0 && id < ix->id0 + ix->ntotal) ? id - ix->id0 : -1;
    if (build_rev(ix) != 0) return -1;
    auto it = ix->rev.find(id);
    return it == ix->rev.end() ? -1 : it->second;
}

extern "C" int vrq_index_reconstruct(vrq_index* ix, int64_t id, uint8_t* code_out) {
    VRQ_CHECK_ARG(ix != nullptr && code_out != nullptr, "null argument");
    const int64_t p = vrq_index_position_of(ix, id);
    if (p < 0) {
        vrq_set_error("reconstruct: id %lld not found", (long long)id);
        return VRQ_ERR_ARG;
    }
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    VRQ_CUDA(cudaMemcpyAsync(code_out, ix->codes + (size_t)p * ix->code_bytes, ix->code_bytes, cudaMemcpyDefault, ix->ctx->stream));
    VRQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));
    return 0;
}

extern "C" int vrq_index_get_payload(vrq_index* ix, int64_t m, const int64_t* positions, void* payload_out, void* aux_out) {
    VRQ_CHECK_ARG(ix != nullptr && m >= 0, "bad argument");
    if (ix) VRQ_TRY(flush_dead(ix));
    if (m == 0) return 0;
    VRQ_CHECK_ARG(positions != nullptr, "positions is null");
    if (ix->payload_kind == VRQ_PAYLOAD_NONE || ix->synth_payload) {
        vrq_set_error("index has no stored payload");
        return VRQ_ERR_STATE;
    }
    const void* all[3] = {positions, payload_out, aux_out};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 3, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    if (!is_dev)
        for (int64_t i = 0; i < m; i++) VRQ_CHECK_ARG(positions[i] >= 0 && positions[i] < ix->ntotal, "position out of range");
    DevIO io{ctx, !is_dev, {}};
    const void* dp;
    void *po, *ao;
    VRQ_TRY(io.in(positions, sizeof(int64_t) * (size_t)m, VRQ_WS_SEARCH_B, &dp));
    VRQ_TRY(io.out(payload_out, ix->payload_row * (size_t)m, VRQ_WS_OUT_A, &po));
    VRQ_TRY(io.out(aux_out, ix->aux_row * (size_t)m, VRQ_WS_OUT_B, &ao));
    if (po) VRQ_TRY(gather_rows(ctx, ix->payload, (const int64_t*)dp, m, ix->payload_row, (uint8_t*)po, ctx->stream));
    if (ao && ix->aux_row) VRQ_TRY(gather_rows(ctx, ix->aux, (const int64_t*)dp, m, ix->aux_row, (uint8_t*)ao, ctx->stream));
    return io.finish();
}

extern "C" int64_t vrq_index_remove_ids(vrq_index* ix, int64_t n, const int64_t* ids) {
    if (!ix || n < 0 || (n > 0 && !ids)) {
        vrq_set_error("remove_ids: bad argument");
        return VRQ_ERR_ARG;
    }
    if (n == 0 || ix->ntotal == 0) return 0;
    if (ix->synth_payload) {
        vrq_set_error("remove_ids: rows of an index with a synthetic payload cannot be removed (position = generator row)");
        return VRQ_ERR_STATE;
    }
    vrq_ctx* ctx = ix->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return VRQ_ERR_STATE;
    bool is_dev;
    vrq_is_device_ptr(ids, &is_dev);
    std::vector<int64_t> hids((size_t)n);
    if (is_dev) {
        if (cudaMemcpy(hids.data(), ids, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return VRQ_ERR_STATE;
    } else {
        memcpy(hids.data(), ids, sizeof(int64_t) * (size_t)n);
    }
    // positions (in the coordinates of the arrays as they are now) of every row whose id is listed
    std::vector<int64_t> pos;
    if (ix->implicit_ids) {
        for (int64_t id : hids)
            if (id >= ix->id0 && id < ix->id0 + ix->ntotal) pos.push_back(id - ix->id0);
    } else {
        // explicit ids may repeat (IDMap2 keeps duplicates and removes all of them): scan the id column on the host
        if (load_host_ids(ix) != 0) return VRQ_ERR_STATE;
        std::unordered_set<int64_t> kill(hids.begin(), hids.end());
        for (int64_t i = 0; i < ix->ntotal; i++)
            if (kill.count(ix->host_ids[(size_t)i])) pos.push_back(i);
    }
    std::sort(pos.begin(), pos.end());
    pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
    std::vector<int64_t> merged;
    merged.reserve(ix->dead.size() + pos.size());
    std::set_union(ix->dead.begin(), ix->dead.end(), pos.begin(), pos.end(), std::back_inserter(merged));
    const int64_t removed = (int64_t)merged.size() - (int64_t)ix->dead.size();
    ix->dead.swap(merged);
    ix->rev_valid = false;  // position_of / reconstruct must not find the removed rows
    return removed;
}

// ---- row access, attachable payloads -------------------------------------------------------------------------------------
namespace {
int rows_of(vrq_index* ix, int which, uint8_t** base, size_t* row) {
    switch (which) {
        case VRQ_ROWS_CODES: *base = ix->codes; *row = (size_t)ix->code_bytes; return 0;
        case VRQ_ROWS_IDS: *base = (uint8_t*)ix->ids; *row = sizeof(int64_t); return 0;
        case VRQ_ROWS_PAYLOAD: *base = ix->payload; *row = ix->payload_row; return 0;
        case VRQ_ROWS_AUX: *base = ix->aux; *row = ix->aux_row; return 0;
    }
    vrq_set_error("unknown row array %d", which);
    return VRQ_ERR_ARG;
}
}  // namespace

extern "C" int vrq_index_read_rows(vrq_index* ix, int which, int64_t offset, int64_t count, void* out) {
    VRQ_CHECK_ARG(ix != nullptr && offset >= 0 && count >= 0, "bad argument");
    VRQ_TRY(flush_dead(ix));
    if (count == 0) return 0;
    VRQ_CHECK_ARG(out != nullptr, "out is null");
    VRQ_CHECK_ARG(offset + count <= ix->ntotal, "row range beyond ntotal");
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    if (which == VRQ_ROWS_IDS) VRQ_TRY(materialise_ids(ix));
    uint8_t* base;
    size_t row;
    VRQ_TRY(rows_of(ix, which, &base, &row));
    if (row == 0 || base == nullptr) {
        vrq_set_error("the index holds no such rows");
        return VRQ_ERR_STATE;
    }
    VRQ_CUDA(cudaMemcpyAsync(out, base + (size_t)offset * row, (size_t)count * row, cudaMemcpyDefault, ix->ctx->stream));
    bool is_dev;
    vrq_is_device_ptr(out, &is_dev);
    if (!is_dev) VRQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));
    return 0;
}

extern "C" int vrq_index_write_rows(vrq_index* ix, int which, int64_t offset, int64_t count, const void* src) {
    VRQ_CHECK_ARG(ix != nullptr && offset >= 0 && count >= 0, "bad argument");
    VRQ_CHECK_ARG(which == VRQ_ROWS_PAYLOAD || which == VRQ_ROWS_AUX || which == VRQ_ROWS_CODES, "ids cannot be overwritten in place");
    VRQ_TRY(flush_dead(ix));
    if (count == 0) return 0;
    VRQ_CHECK_ARG(src != nullptr, "src is null");
    VRQ_CHECK_ARG(offset + count <= ix->ntotal, "row range beyond ntotal");
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    uint8_t* base;
    size_t row;
    VRQ_TRY(rows_of(ix, which, &base, &row));
    if (row == 0 || base == nullptr) {
        vrq_set_error("the index holds no such rows");
        return VRQ_ERR_STATE;
    }
    VRQ_CUDA(cudaMemcpyAsync(base + (size_t)offset * row, src, (size_t)count * row, cudaMemcpyDefault, ix->ctx->stream));
    bool is_dev;
    vrq_is_device_ptr(src, &is_dev);
    if (!is_dev) VRQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));
    return 0;
}

extern "C" int vrq_index_attach_payload(vrq_index* ix, int kind, double global_limit) {
    VRQ_CHECK_ARG(ix != nullptr, "index is null");
    VRQ_CHECK_ARG(kind > VRQ_PAYLOAD_NONE && kind <= VRQ_PAYLOAD_CODES_PM1, "unknown payload kind");
    if (ix->payload_kind != VRQ_PAYLOAD_NONE) {
        vrq_set_error("the index already has a payload");
        return VRQ_ERR_STATE;
    }
    VRQ_TRY(flush_dead(ix));
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    const size_t prow = payload_row_bytes(kind, ix->d), arow = aux_row_bytes(kind);
    void *pp = nullptr, *ap = nullptr;
    if (ix->capacity > 0 && prow) {
        cudaError_t e = cudaMalloc(&pp, prow * (size_t)ix->capacity);
        if (e == cudaSuccess && arow) e = cudaMalloc(&ap, arow * (size_t)ix->capacity);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (pp) cudaFree(pp);
            vrq_set_error("cudaMalloc of %zu payload bytes failed: %s", prow * (size_t)ix->capacity, cudaGetErrorString(e));
            return VRQ_ERR_NOMEM;
        }
        VRQ_CUDA(cudaMemsetAsync(pp, 0, prow * (size_t)ix->capacity, ix->ctx->stream));
        if (ap) VRQ_CUDA(cudaMemsetAsync(ap, 0, arow * (size_t)ix->capacity, ix->ctx->stream));
    }
    ix->payload_kind = kind;
    ix->limit = global_limit;
    ix->payload_row = prow;
    ix->aux_row = arow;
    ix->payload = (uint8_t*)pp;
    ix->aux = (uint8_t*)ap;
    return 0;
}

// ---- files -------------------------------------------------------------------------------------------------------------
// Every file is written to "<path>.tmp", flushed, fsync'ed and renamed over <path>: a crash leaves the old file or the new
// one, never a torn one.  Arrays stream between the file and device memory through one pinned 64 MB bounce buffer.
namespace {

constexpr size_t IO_CHUNK = 64u << 20;

struct PinnedBounce {
    void* p = nullptr;
    ~PinnedBounce() {
        if (p) cudaFreeHost(p);
    }
    int get(size_t bytes) {
        if (p) return 0;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            vrq_set_error("cudaHostAlloc of the %zu-byte IO bounce buffer failed", bytes);
            return VRQ_ERR_NOMEM;
        }
        return 0;
    }
};

struct AtomicFile {
    std::string path, tmp;
    FILE* f = nullptr;
    int open(const char* p) {
        path = p;
        tmp = path + ".tmp";
        f = fopen(tmp.c_str(), "wb");
        if (!f) {
            vrq_set_error("cannot open %s for writing: %s", tmp.c_str(), strerror(errno));
            return VRQ_ERR_IO;
        }
        return 0;
    }
    bool put(const void* d, size_t n) { return n == 0 || fwrite(d, 1, n, f) == n; }
    int fail(const char* why) {
        if (f) fclose(f);
        f = nullptr;
        remove(tmp.c_str());
        vrq_set_error("%s: %s", path.c_str(), why);
        return VRQ_ERR_IO;
    }
    int commit() {
        if (fflush(f) != 0 || fsync(fileno(f)) != 0) return fail("flush failed");
        if (fclose(f) != 0) {
            f = nullptr;
            return fail("close failed");
        }
        f = nullptr;
        if (rename(tmp.c_str(), path.c_str()) != 0) return fail("rename failed");
        return 0;
    }
    ~AtomicFile() {
        if (f) {
            fclose(f);
            remove(tmp.c_str());
        }
    }
};

// device array -> file
int stream_out(vrq_index* ix, AtomicFile& af, const uint8_t* dev, size_t nbytes, PinnedBounce& pb) {
    if (nbytes == 0) return 0;
    VRQ_TRY(pb.get(IO_CHUNK));
    for (size_t off = 0; off < nbytes; off += IO_CHUNK) {
        const size_t len = std::min(IO_CHUNK, nbytes - off);
        cudaError_t e = cudaMemcpyAsync(pb.p, dev + off, len, cudaMemcpyDeviceToHost, ix->ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->ctx->stream);
        if (e != cudaSuccess) {
            af.fail(cudaGetErrorString(e));
            return (int)e;
        }
        if (!af.put(pb.p, len)) return af.fail("short write");
    }
    return 0;
}

// file -> device array
int stream_in(vrq_ctx* ctx, FILE* f, uint8_t* dev, size_t nbytes, PinnedBounce& pb, const char* path) {
    if (nbytes == 0) return 0;
    VRQ_TRY(pb.get(IO_CHUNK));
    for (size_t off = 0; off < nbytes; off += IO_CHUNK) {
        const size_t len = std::min(IO_CHUNK, nbytes - off);
        if (fread(pb.p, 1, len, f) != len) {
            vrq_set_error("%s: truncated file", path);
            return VRQ_ERR_IO;
        }
        VRQ_CUDA(cudaMemcpyAsync(dev + off, pb.p, len, cudaMemcpyHostToDevice, ctx->stream));
        VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int64_t file_size(FILE* f) {
    struct stat st;
    if (fstat(fileno(f), &st) != 0) return -1;
    return (int64_t)st.st_size;
}

}  // namespace

// faiss file format ("IBM2" wrapping "IBxF"), SURVEY App. B.1
extern "C" int vrq_index_write(vrq_index* ix, const char* path) {
    VRQ_CHECK_ARG(ix != nullptr && path != nullptr, "null argument");
    VRQ_TRY(flush_dead(ix));
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    AtomicFile af;
    VRQ_TRY(af.open(path));
    auto hdr = [&](const char* fourcc) -> bool {
        int32_t d = ix->d, cs = ix->code_bytes, metric = 1;
        int64_t nt = ix->ntotal;
        uint8_t trained = 1;
        return af.put(fourcc, 4) && af.put(&d, 4) && af.put(&cs, 4) && af.put(&nt, 8) && af.put(&trained, 1) && af.put(&metric, 4);
    };
    const uint64_t nbytes = (uint64_t)ix->ntotal * ix->code_bytes, nids = (uint64_t)ix->ntotal;
    if (!hdr("IBM2") || !hdr("IBxF") || !af.put(&nbytes, 8)) return af.fail("short write");
    PinnedBounce pb;
    VRQ_TRY(stream_out(ix, af, ix->codes, (size_t)nbytes, pb));
    if (!af.put(&nids, 8)) return af.fail("short write");
    if (ix->implicit_ids) {
        // ids are id0 + position: produced on the fly, no device or host array needed
        std::vector<int64_t> tmp(std::min<size_t>((size_t)nids, 1u << 20));
        for (uint64_t off = 0; off < nids; off += tmp.size()) {
            const size_t len = (size_t)std::min<uint64_t>(tmp.size(), nids - off);
            for (size_t i = 0; i < len; i++) tmp[i] = ix->id0 + (int64_t)(off + i);
            if (!af.put(tmp.data(), 8 * len)) return af.fail("short write");
        }
    } else {
        VRQ_TRY(stream_out(ix, af, (const uint8_t*)ix->ids, (size_t)nids * 8, pb));
    }
    return af.commit();
}

extern "C" int vrq_index_read(vrq_ctx* ctx, const char* path, vrq_index** out) {
    VRQ_CHECK_ARG(ctx != nullptr && path != nullptr && out != nullptr, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) {
        vrq_set_error("cannot open %s", path);
        return VRQ_ERR_IO;
    }
    vrq_index* ix = nullptr;
    auto fail = [&](const char* why) {
        fclose(f);
        if (ix) vrq_index_free(ix);
        vrq_set_error("%s: %s", path, why);
        return VRQ_ERR_IO;
    };
    struct H {
        char cc[4];
        int32_t d, cs;
        int64_t nt;
        uint8_t trained;
        int32_t metric;
    };
    auto rd = [&](H* h) -> bool {
        return fread(h->cc, 1, 4, f) == 4 && fread(&h->d, 4, 1, f) == 1 && fread(&h->cs, 4, 1, f) == 1 &&
               fread(&h->nt, 8, 1, f) == 1 && fread(&h->trained, 1, 1, f) == 1 && fread(&h->metric, 4, 1, f) == 1;
    };
    H h1, h2;
    if (!rd(&h1) || memcmp(h1.cc, "IBM2", 4) != 0) return fail("not an IndexBinaryIDMap2 file (fourcc IBM2 expected)");
    if (!rd(&h2) || memcmp(h2.cc, "IBxF", 4) != 0) return fail("inner index is not IndexBinaryFlat (fourcc IBxF expected)");
    if (h2.d <= 0 || h2.d % 8 != 0 || h2.cs != h2.d / 8 || h2.nt < 0) return fail("corrupt header");
    uint64_t nbytes = 0;
    if (fread(&nbytes, 8, 1, f) != 1 || nbytes != (uint64_t)h2.nt * h2.cs) return fail("code array size mismatch");
    // the header must agree with the size of the file before anything is allocated from it
    const int64_t fsz = file_size(f);
    if (fsz >= 0 && (uint64_t)fsz != 58 + nbytes + 8 + 8 * (uint64_t)h2.nt) return fail("file size does not match its header");
    int r = vrq_index_create(ctx, h2.d, &ix);
    if (r != 0) {
        fclose(f);
        return r;
    }
    r = cudaSetDevice(ctx->device) == cudaSuccess ? 0 : VRQ_ERR_STATE;
    if (r == 0 && h2.nt > 0) r = ensure_capacity(ix, h2.nt);
    PinnedBounce pb;
    if (r == 0) r = stream_in(ctx, f, ix->codes, (size_t)nbytes, pb, path);
    uint64_t nids = 0;
    if (r == 0 && (fread(&nids, 8, 1, f) != 1 || nids != (uint64_t)h2.nt)) return fail("id map size mismatch");
    if (r == 0 && h2.nt > 0) {
        ix->implicit_ids = false;
        if (cudaMalloc((void**)&ix->ids, sizeof(int64_t) * (size_t)ix->capacity) != cudaSuccess) {
            cudaGetLastError();
            vrq_set_error("cudaMalloc of the id map failed");
            r = VRQ_ERR_NOMEM;
        }
        if (r == 0) r = stream_in(ctx, f, (uint8_t*)ix->ids, (size_t)nids * 8, pb, path);
    }
    fclose(f);
    if (r != 0) {
        vrq_index_free(ix);
        return r;
    }
    ix->ntotal = h2.nt;
    *out = ix;
    return 0;
}

// Payload sidecar ("VRQP"): what the reference keeps in RocksDB pickles next to index.bin (CohereEnhancedVectorDB.py:221 {"int8"},
// VectorDBInt8.py:179 {"emb_int8","min_max"}), as one flat file: header, payload rows, aux rows - streamed like the codes.
struct VrqpHeader {
    char magic[4];
    int32_t version, kind, d;
    int64_t ntotal;
    uint64_t payload_row, aux_row;
    double limit;
};

extern "C" int vrq_index_write_payload(vrq_index* ix, const char* path) {
    VRQ_CHECK_ARG(ix != nullptr && path != nullptr, "null argument");
    VRQ_TRY(flush_dead(ix));
    if (ix->payload_kind == VRQ_PAYLOAD_NONE || ix->synth_payload) {
        vrq_set_error("index has no stored payload");
        return VRQ_ERR_STATE;
    }
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    AtomicFile af;
    VRQ_TRY(af.open(path));
    VrqpHeader h{};
    memcpy(h.magic, "VRQP", 4);
    h.version = 1;
    h.kind = ix->payload_kind;
    h.d = ix->d;
    h.ntotal = ix->ntotal;
    h.payload_row = ix->payload_row;
    h.aux_row = ix->aux_row;
    h.limit = ix->limit;
    if (!af.put(&h, sizeof(h))) return af.fail("short write");
    PinnedBounce pb;
    VRQ_TRY(stream_out(ix, af, ix->payload, ix->payload_row * (size_t)ix->ntotal, pb));
    VRQ_TRY(stream_out(ix, af, ix->aux, ix->aux_row * (size_t)ix->ntotal, pb));
    return af.commit();
}

extern "C" int vrq_index_read_payload(vrq_index* ix, const char* path) {
    VRQ_CHECK_ARG(ix != nullptr && path != nullptr, "null argument");
    VRQ_TRY(flush_dead(ix));
    FILE* f = fopen(path, "rb");
    if (!f) {
        vrq_set_error("cannot open %s", path);
        return VRQ_ERR_IO;
    }
    VrqpHeader h{};
    auto fail = [&](const char* why) {
        fclose(f);
        vrq_set_error("%s: %s", path, why);
        return VRQ_ERR_IO;
    };
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "VRQP", 4) != 0 || h.version != 1) return fail("not a payload sidecar (VRQP v1)");
    if (h.d != ix->d) return fail("embedding_dim differs from the index");
    if (h.ntotal != ix->ntotal) return fail("row count differs from index.bin (the two files are not from the same save)");
    if (h.kind <= VRQ_PAYLOAD_NONE || h.kind > VRQ_PAYLOAD_CODES_PM1 || h.payload_row != payload_row_bytes(h.kind, h.d) ||
        h.aux_row != aux_row_bytes(h.kind))
        return fail("corrupt header");
    const int64_t fsz = file_size(f);
    if (fsz >= 0 && (uint64_t)fsz != sizeof(h) + (h.payload_row + h.aux_row) * (uint64_t)h.ntotal) return fail("file size does not match its header");
    int r = vrq_index_attach_payload(ix, h.kind, h.limit);
    PinnedBounce pb;
    if (r == 0) r = stream_in(ix->ctx, f, ix->payload, ix->payload_row * (size_t)ix->ntotal, pb, path);
    if (r == 0) r = stream_in(ix->ctx, f, ix->aux, ix->aux_row * (size_t)ix->ntotal, pb, path);
    fclose(f);
    return r;
}

// ---- float32 inner-product index: faiss.IndexIDMap(faiss.IndexFlatIP(d)) of CohereVectorDBFloat.py ----------------------------
extern "C" int vrq_index_search_ip(vrq_index* ix, int64_t nq, const float* q_float, int k, float* scores, int64_t* labels) {
    VRQ_CHECK_ARG(ix != nullptr && nq >= 0, "bad argument");
    VRQ_TRY(flush_dead(ix));
    if (nq == 0) return 0;
    VRQ_CHECK_ARG(q_float && scores && labels, "null pointer");
    if (ix->payload_kind != VRQ_PAYLOAD_F32) {
        vrq_set_error("search_ip needs an index with the F32 payload (float32 rows)");
        return VRQ_ERR_STATE;
    }
    const void* all[3] = {q_float, scores, labels};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 3, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    DevIO io{ctx, !is_dev, {}};
    const void* dq;
    void *ds, *dl;
    VRQ_TRY(io.in(q_float, sizeof(float) * (size_t)nq * ix->d, VRQ_WS_QUERY_A, &dq));
    VRQ_TRY(io.out(scores, sizeof(float) * (size_t)nq * k, VRQ_WS_OUT_A, &ds));
    VRQ_TRY(io.out(labels, sizeof(int64_t) * (size_t)nq * k, VRQ_WS_OUT_B, &dl));
    VRQ_TRY(vrq_launch_ip_topk(ctx, (const float*)ix->payload, ix->ntotal, ix->d, (const float*)dq, nq, k, ix->implicit_ids ? nullptr : ix->ids,
                               ix->id0, (float*)ds, (int64_t*)dl, ctx->stream));
    return io.finish();
}

// faiss.write_index / read_index of IndexIDMap(IndexFlatIP) (CohereVectorDBFloat.py:184,58): "IxMp" header, "IxFI" header,
// u64 count + float32 rows, u64 count + int64 ids; headers = {d i32, ntotal i64, 1 << 20, 1 << 20, is_trained u8, metric i32 = 0}.
extern "C" int vrq_index_write_float(vrq_index* ix, const char* path) {
    VRQ_CHECK_ARG(ix != nullptr && path != nullptr, "null argument");
    VRQ_TRY(flush_dead(ix));
    if (ix->payload_kind != VRQ_PAYLOAD_F32) {
        vrq_set_error("write_float needs an index with the F32 payload");
        return VRQ_ERR_STATE;
    }
    VRQ_CUDA(cudaSetDevice(ix->ctx->device));
    VRQ_TRY(materialise_ids(ix));
    AtomicFile af;
    VRQ_TRY(af.open(path));
    auto hdr = [&](const char* fourcc) -> bool {
        int32_t d = ix->d, metric = 0;
        int64_t nt = ix->ntotal, dummy = 1 << 20;
        uint8_t trained = 1;
        return af.put(fourcc, 4) && af.put(&d, 4) && af.put(&nt, 8) && af.put(&dummy, 8) && af.put(&dummy, 8) && af.put(&trained, 1) &&
               af.put(&metric, 4);
    };
    const uint64_t nfl = (uint64_t)ix->ntotal * ix->d, nids = (uint64_t)ix->ntotal;
    if (!hdr("IxMp") || !hdr("IxFI") || !af.put(&nfl, 8)) return af.fail("short write");
    PinnedBounce pb;
    VRQ_TRY(stream_out(ix, af, ix->payload, (size_t)nfl * 4, pb));
    if (!af.put(&nids, 8)) return af.fail("short write");
    if (nids) VRQ_TRY(stream_out(ix, af, (const uint8_t*)ix->ids, (size_t)nids * 8, pb));
    return af.commit();
}

extern "C" int vrq_index_read_float(vrq_ctx* ctx, const char* path, vrq_index** out) {
    VRQ_CHECK_ARG(ctx != nullptr && path != nullptr && out != nullptr, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) {
        vrq_set_error("cannot open %s", path);
        return VRQ_ERR_IO;
    }
    vrq_index* ix = nullptr;
    auto fail = [&](const char* why) {
        fclose(f);
        if (ix) vrq_index_free(ix);
        vrq_set_error("%s: %s", path, why);
        return VRQ_ERR_IO;
    };
    struct H {
        char cc[4];
        int32_t d;
        int64_t nt, a, b;
        uint8_t trained;
        int32_t metric;
    };
    auto rd = [&](H* h) -> bool {
        return fread(h->cc, 1, 4, f) == 4 && fread(&h->d, 4, 1, f) == 1 && fread(&h->nt, 8, 1, f) == 1 && fread(&h->a, 8, 1, f) == 1 &&
               fread(&h->b, 8, 1, f) == 1 && fread(&h->trained, 1, 1, f) == 1 && fread(&h->metric, 4, 1, f) == 1;
    };
    H h1, h2;
    if (!rd(&h1) || memcmp(h1.cc, "IxMp", 4) != 0) return fail("not an IndexIDMap file (fourcc IxMp expected)");
    if (!rd(&h2) || memcmp(h2.cc, "IxFI", 4) != 0) return fail("inner index is not IndexFlatIP (fourcc IxFI expected)");
    if (h2.d <= 0 || h2.d % 8 != 0 || h2.nt < 0 || h2.metric != 0) return fail("unsupported header (d % 8 == 0 and the inner-product metric are required)");
    uint64_t nfl = 0;
    if (fread(&nfl, 8, 1, f) != 1 || nfl != (uint64_t)h2.nt * h2.d) return fail("vector array size mismatch");
    const int64_t fsz = file_size(f);
    if (fsz >= 0 && (uint64_t)fsz != 74 + 8 + 4 * nfl + 8 + 8 * (uint64_t)h2.nt) return fail("file size does not match its header");
    int r = vrq_index_create(ctx, h2.d, &ix);
    if (r == 0) r = vrq_index_set_payload(ix, VRQ_PAYLOAD_F32, 0.0);
    if (r != 0) {
        fclose(f);
        if (ix) vrq_index_free(ix);
        return r;
    }
    r = cudaSetDevice(ctx->device) == cudaSuccess ? 0 : VRQ_ERR_STATE;
    if (r == 0 && h2.nt > 0) r = ensure_capacity(ix, h2.nt);
    if (r == 0 && h2.nt > 0 && cudaMemsetAsync(ix->codes, 0, (size_t)h2.nt * ix->code_bytes, ctx->stream) != cudaSuccess) r = VRQ_ERR_STATE;
    PinnedBounce pb;
    if (r == 0) r = stream_in(ctx, f, ix->payload, (size_t)nfl * 4, pb, path);
    uint64_t nids = 0;
    if (r == 0 && (fread(&nids, 8, 1, f) != 1 || nids != (uint64_t)h2.nt)) return fail("id map size mismatch");
    if (r == 0 && h2.nt > 0) {
        ix->implicit_ids = false;
        if (cudaMalloc((void**)&ix->ids, sizeof(int64_t) * (size_t)ix->capacity) != cudaSuccess) {
            cudaGetLastError();
            vrq_set_error("cudaMalloc of the id map failed");
            r = VRQ_ERR_NOMEM;
        }
        if (r == 0) r = stream_in(ctx, f, (uint8_t*)ix->ids, (size_t)nids * 8, pb, path);
    }
    fclose(f);
    if (r != 0) {
        vrq_index_free(ix);
        return r;
    }
    ix->ntotal = h2.nt;
    *out = ix;
    return 0;
}

// ---- fused searches ---------------------------------------------------------------------------------------------------
extern "C" int vrq_index_search3_local(vrq_index* ix, int64_t nq, const float* q_float, const uint8_t* q_ubin, int binary_k,
                                       int64_t pos_base, uint64_t* keys, int64_t* labels, double* score_binary,
                                       double* score_cosine) {
    VRQ_CHECK_ARG(ix && q_float && q_ubin && keys && labels && score_binary && score_cosine, "null argument");
    if (ix) VRQ_TRY(flush_dead(ix));
    VRQ_CHECK_ARG(nq >= 0 && binary_k > 0, "bad sizes");
    if (ix->payload_kind != VRQ_PAYLOAD_INT8_RAW) {
        vrq_set_error("search3 needs an index with the INT8_RAW payload (CohereEnhancedVectorDB layout)");
        return VRQ_ERR_STATE;
    }
    if (nq == 0) return 0;
    vrq_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    VRQ_TRY(vrq_hamming_topk_dev(ctx, ix->codes, ix->ntotal, ix->code_bytes, pos_base, q_ubin, nq, binary_k, keys, st));
    VRQ_TRY(vrq_launch_keys_to_dist_labels(ctx, keys, nq * (int64_t)binary_k, pos_base, ix->implicit_ids ? nullptr : ix->ids,
                                           ix->id0, nullptr, labels, st));
    VRQ_TRY(vrq_launch_rescore_binary(ctx, ix->codes, ix->d, keys, nullptr, pos_base, nq, binary_k, q_float, score_binary, st));
    if (ix->synth_payload) {
        if (ix->d != 1024) {
            vrq_set_error("the synthetic payload needs d == 1024");
            return VRQ_ERR_UNSUPPORTED;
        }
        VRQ_TRY(vrq_launch_rescore_int8cos_synth(ctx, ix->synth_seed, ix->synth_row0, keys, nullptr, pos_base, nq, binary_k, q_float,
                                                 score_cosine, st));
        return 0;
    }
    VRQ_TRY(vrq_launch_rescore_int8cos(ctx, (const int8_t*)ix->payload, ix->d, keys, nullptr, pos_base, nq, binary_k, q_float,
                                       score_cosine, st));
    return 0;
}

extern "C" int vrq_index_search3(vrq_index* ix, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k,
                                 int binary_oversample, int int8_oversample, int64_t* labels, int32_t* hamming,
                                 double* score_binary, double* score_cosine, int32_t* out_count) {
    VRQ_CHECK_ARG(ix && q_float && q_ubin && labels && hamming && score_binary && score_cosine && out_count, "null argument");
    VRQ_TRY(flush_dead(ix));
    VRQ_CHECK_ARG(nq >= 0 && k > 0 && binary_oversample > 0 && int8_oversample > 0, "bad sizes");
    if (nq == 0) return 0;
    const void* all[7] = {q_float, q_ubin, labels, hamming, score_binary, score_cosine, out_count};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 7, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    // binary_k = min(k * binary_oversample, ntotal)   (CohereEnhancedVectorDB.py:267)
    int64_t bk64 = std::min<int64_t>((int64_t)k * binary_oversample, ix->ntotal);
    if (bk64 > VRQ_MAX_K) {
        vrq_set_error("k * binary_oversample = %lld exceeds the supported %d", (long long)bk64, VRQ_MAX_K);
        return VRQ_ERR_UNSUPPORTED;
    }
    DevIO io{ctx, !is_dev, {}};
    const void *dqf, *dqb;
    void *ol, *oh, *ob, *oc, *on;
    VRQ_TRY(io.in(q_float, sizeof(float) * (size_t)nq * ix->d, VRQ_WS_QUERY_A, &dqf));
    VRQ_TRY(io.in(q_ubin, (size_t)nq * ix->code_bytes, VRQ_WS_QUERY_B, &dqb));
    VRQ_TRY(io.out(labels, 8 * (size_t)nq * k, VRQ_WS_OUT_A, &ol));
    VRQ_TRY(io.out(hamming, 4 * (size_t)nq * k, VRQ_WS_OUT_B, &oh));
    VRQ_TRY(io.out(score_binary, 8 * (size_t)nq * k, VRQ_WS_OUT_C, &ob));
    VRQ_TRY(io.out(score_cosine, 8 * (size_t)nq * k, VRQ_WS_OUT_D, &oc));
    VRQ_TRY(io.out(out_count, 4 * (size_t)nq, VRQ_WS_OUT_E, &on));
    if (bk64 == 0) {
        // empty index: the reference logs an error and returns [] (CohereEnhancedVectorDB.py:247-249)
        VRQ_CUDA(cudaMemsetAsync(on, 0, 4 * (size_t)nq, ctx->stream));
        VRQ_CUDA(cudaMemsetAsync(ol, 0xFF, 8 * (size_t)nq * k, ctx->stream));
        return io.finish();
    }
    const int bk = (int)bk64;
    void *keys, *lab, *sb, *sc;
    const size_t cnt = (size_t)nq * bk;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_A, 8 * cnt, &keys));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_C, 8 * cnt, &lab));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_D, 8 * cnt, &sb));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_E, 8 * cnt, &sc));
    VRQ_TRY(vrq_index_search3_local(ix, nq, (const float*)dqf, (const uint8_t*)dqb, bk, 0, (uint64_t*)keys, (int64_t*)lab,
                                    (double*)sb, (double*)sc));
    VRQ_TRY(vrq_launch_merge3(ctx, 1, nq, bk, 0, (const uint64_t*)keys, (const int64_t*)lab, (const double*)sb, (const double*)sc, k,
                              k * int8_oversample, (int64_t*)ol, (int32_t*)oh, (double*)ob, (double*)oc, (int32_t*)on, ctx->stream));
    return io.finish();
}

extern "C" int vrq_index_search2(vrq_index* ix, int64_t nq, const float* q_float, const uint8_t* q_ubin, int k,
                                 int binary_oversample, int64_t* labels, float* score, int32_t* out_count) {
    VRQ_CHECK_ARG(ix && q_float && labels && score && out_count, "null argument");
    VRQ_TRY(flush_dead(ix));
    VRQ_CHECK_ARG(nq >= 0 && k > 0 && binary_oversample > 0, "bad sizes");
    if (ix->payload_kind == VRQ_PAYLOAD_NONE || ix->payload_kind == VRQ_PAYLOAD_INT8_RAW) {
        vrq_set_error("search2 needs a quantised (or float32) payload");
        return VRQ_ERR_STATE;
    }
    if (nq == 0) return 0;
    const void* all[5] = {q_float, q_ubin, labels, score, out_count};  // vrq_space_of skips null pointers
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 5, &is_dev));
    vrq_ctx* ctx = ix->ctx;
    VRQ_CUDA(cudaSetDevice(ctx->device));
    int64_t bk64 = std::min<int64_t>((int64_t)k * binary_oversample, ix->ntotal);
    if (bk64 > VRQ_MAX_K) {
        vrq_set_error("k * binary_oversample = %lld exceeds the supported %d", (long long)bk64, VRQ_MAX_K);
        return VRQ_ERR_UNSUPPORTED;
    }
    DevIO io{ctx, !is_dev, {}};
    const void *dqf, *dqb;
    void *ol, *os, *on;
    VRQ_TRY(io.in(q_float, sizeof(float) * (size_t)nq * ix->d, VRQ_WS_QUERY_A, &dqf));
    if (q_ubin) {
        VRQ_TRY(io.in(q_ubin, (size_t)nq * ix->code_bytes, VRQ_WS_QUERY_B, &dqb));
    } else {
        // query_bin = self._to_binary(query float) (VectorDBInt8.py:213, :140-146) on the device: packbits(q > mean_f32(q)) - one
        // host round trip less for the one-query-per-call pattern of the classes
        void* qb_v;
        VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_QUERY_B, (size_t)nq * ix->code_bytes, &qb_v));
        vrq_encode_args ea{};
        ea.x = (const float*)dqf;
        ea.n = nq;
        ea.d = ix->d;
        ea.codec = VRQ_CODEC_NONE;
        ea.limit_f32 = 1.0f;
        ea.scale_f32 = 1.0f;
        ea.ubin = (uint8_t*)qb_v;
        ea.ge = 0;
        VRQ_TRY(vrq_launch_encode(ctx, ea, ctx->stream));
        dqb = qb_v;
    }
    VRQ_TRY(io.out(labels, 8 * (size_t)nq * k, VRQ_WS_OUT_A, &ol));
    VRQ_TRY(io.out(score, 4 * (size_t)nq * k, VRQ_WS_OUT_B, &os));
    VRQ_TRY(io.out(out_count, 4 * (size_t)nq, VRQ_WS_OUT_E, &on));
    if (bk64 == 0) {
        VRQ_CUDA(cudaMemsetAsync(on, 0, 4 * (size_t)nq, ctx->stream));
        VRQ_CUDA(cudaMemsetAsync(ol, 0xFF, 8 * (size_t)nq * k, ctx->stream));
        return io.finish();
    }
    const int bk = (int)bk64;
    const size_t cnt = (size_t)nq * bk;
    void *keys, *lab, *sc;
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_A, 8 * cnt, &keys));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_C, 8 * cnt, &lab));
    VRQ_TRY(vrq_ws_get(ctx, VRQ_WS_SEARCH_D, 4 * cnt, &sc));
    cudaStream_t st = ctx->stream;
    VRQ_TRY(vrq_hamming_topk_dev(ctx, ix->codes, ix->ntotal, ix->code_bytes, 0, (const uint8_t*)dqb, nq, bk, (uint64_t*)keys, st));
    VRQ_TRY(vrq_launch_keys_to_dist_labels(ctx, (const uint64_t*)keys, (int64_t)cnt, 0, ix->implicit_ids ? nullptr : ix->ids, ix->id0,
                                           nullptr, (int64_t*)lab, st));
    vrq_rescore2_args a{};
    a.kind = ix->payload_kind;
    a.payload = ix->payload_kind == VRQ_PAYLOAD_CODES_PM1 ? (const void*)ix->codes : (const void*)ix->payload;
    a.aux = ix->aux;
    a.limit = ix->limit;
    a.d = ix->d;
    a.keys = (const uint64_t*)keys;
    a.pos_base = 0;
    a.nq = nq;
    a.m = bk;
    a.qf = (const float*)dqf;
    a.score = (float*)sc;
    VRQ_TRY(vrq_launch_rescore_payload_dot(ctx, a, st));
    VRQ_TRY(vrq_launch_select2(ctx, nq, bk, (const uint64_t*)keys, (const int64_t*)lab, (const float*)sc, k, (int64_t*)ol, (float*)os,
                               (int32_t*)on, st));
    return io.finish();
}
