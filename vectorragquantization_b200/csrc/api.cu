// C-ABI entry points for the encoders / decoders / stand-alone kernels: argument checking, host-vs-device
// pointer dispatch and the chunked host pipeline (H2D of chunk i+1 and D2H of chunk i-1 overlap the kernel of
// chunk i on two streams).
#include <string.h>

#include <algorithm>

#include "vrq_internal.cuh"

namespace {

struct IoBuf {
    const void* in;   // host input  (one of in/out is set)
    void* out;        // host output
    size_t row_bytes;
    int ws_slot[2];
};

// Run `launch(dev_ptrs, row0, rows, stream)` over n rows.  bufs lists the row-wise inputs / outputs.
// Device space: a single launch on ctx->stream with the caller's pointers.  Host space: chunked pipeline.
template <class Launch>
int run_rows(vrq_ctx* ctx, int64_t n, std::vector<IoBuf>& bufs, bool is_dev, Launch launch) {
    VRQ_CUDA(cudaSetDevice(ctx->device));
    const int nb = (int)bufs.size();
    std::vector<void*> ptrs(nb);
    if (is_dev) {
        for (int i = 0; i < nb; i++) ptrs[i] = bufs[i].in ? const_cast<void*>(bufs[i].in) : bufs[i].out;
        return launch(ptrs.data(), (int64_t)0, n, ctx->stream);
    }
    size_t per_row = 0;
    for (auto& b : bufs) per_row += b.row_bytes;
    int64_t chunk = (int64_t)((size_t)(96u << 20) / std::max<size_t>(per_row, 1));
    chunk = std::max<int64_t>(256, std::min<int64_t>(chunk, n));
    // stage buffers: slot 0/1, all bufs packed one after another inside VRQ_WS_STAGE_IN{0,1}
    void* stage[2];
    for (int s = 0; s < 2; s++) VRQ_TRY(vrq_ws_get(ctx, s == 0 ? VRQ_WS_STAGE_IN0 : VRQ_WS_STAGE_IN1, per_row * (size_t)chunk + 256 * nb, &stage[s]));
    VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
    int it = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk, it++) {
        const int s = it & 1;
        const int64_t rows = std::min<int64_t>(chunk, n - r0);
        cudaStream_t st = ctx->pipe[s];
        size_t off = 0;
        for (int i = 0; i < nb; i++) {
            ptrs[i] = (uint8_t*)stage[s] + off;
            off += (bufs[i].row_bytes * (size_t)chunk + 255) & ~(size_t)255;
        }
        for (int i = 0; i < nb; i++)
            if (bufs[i].in)
                VRQ_CUDA(cudaMemcpyAsync(ptrs[i], (const uint8_t*)bufs[i].in + (size_t)r0 * bufs[i].row_bytes,
                                         bufs[i].row_bytes * (size_t)rows, cudaMemcpyHostToDevice, st));
        VRQ_TRY(launch(ptrs.data(), r0, rows, st));
        for (int i = 0; i < nb; i++)
            if (bufs[i].out)
                VRQ_CUDA(cudaMemcpyAsync((uint8_t*)bufs[i].out + (size_t)r0 * bufs[i].row_bytes, ptrs[i],
                                         bufs[i].row_bytes * (size_t)rows, cudaMemcpyDeviceToHost, st));
    }
    VRQ_CUDA(cudaStreamSynchronize(ctx->pipe[0]));
    VRQ_CUDA(cudaStreamSynchronize(ctx->pipe[1]));
    return 0;
}

int check_common(vrq_ctx* ctx, const void* x, int64_t n, int d) {
    VRQ_CHECK_ARG(ctx != nullptr, "ctx is null");
    VRQ_CHECK_ARG(n >= 0, "n < 0");
    VRQ_CHECK_ARG(n == 0 || x != nullptr, "input pointer is null");
    VRQ_CHECK_ARG(d > 0 && d % 8 == 0, "embedding_dim must be a positive multiple of 8");
    return 0;
}

int encode_common(vrq_ctx* ctx, int codec, const float* x, int64_t n, int d, double limit, double qmax, void* q,
                  size_t q_row_bytes, void* mn, void* mx, size_t stat_bytes, uint8_t* ubin, int ge) {
    VRQ_TRY(check_common(ctx, x, n, d));
    if (n == 0) return 0;
    if (codec == VRQ_CODEC_INT8_GLOBAL || codec == VRQ_CODEC_INT16_GLOBAL)
        VRQ_CHECK_ARG(limit > 0.0, "global_limit must be > 0");
    const void* all[5] = {x, q, mn, mx, ubin};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 5, &is_dev));
    std::vector<IoBuf> bufs;
    bufs.push_back({x, nullptr, sizeof(float) * (size_t)d, {0, 0}});
    int iq = -1, imn = -1, imx = -1, iub = -1;
    if (q) { iq = (int)bufs.size(); bufs.push_back({nullptr, q, q_row_bytes, {0, 0}}); }
    if (mn) { imn = (int)bufs.size(); bufs.push_back({nullptr, mn, stat_bytes, {0, 0}}); }
    if (mx) { imx = (int)bufs.size(); bufs.push_back({nullptr, mx, stat_bytes, {0, 0}}); }
    if (ubin) { iub = (int)bufs.size(); bufs.push_back({nullptr, ubin, (size_t)d / 8, {0, 0}}); }
    const float lim32 = (float)limit;              // np.clip bound: np.float32(limit)
    const float scale32 = (float)(qmax / limit);   // np.float32(qmax / limit): float64 divide, one rounding
    return run_rows(ctx, n, bufs, is_dev, [&](void** p, int64_t, int64_t rows, cudaStream_t st) {
        vrq_encode_args a{};
        a.x = (const float*)p[0];
        a.n = rows;
        a.d = d;
        a.codec = codec;
        a.limit_f32 = lim32;
        a.scale_f32 = scale32;
        a.q = iq >= 0 ? p[iq] : nullptr;
        a.mn = imn >= 0 ? p[imn] : nullptr;
        a.mx = imx >= 0 ? p[imx] : nullptr;
        a.ubin = iub >= 0 ? (uint8_t*)p[iub] : nullptr;
        a.ge = ge;
        return vrq_launch_encode(ctx, a, st);
    });
}

}  // namespace

extern "C" int vrq_quantize_int8_perdoc(vrq_ctx* ctx, const float* x, int64_t n, int d, int8_t* q, float* mn, float* mx,
                                        uint8_t* ubin) {
    VRQ_CHECK_ARG(n == 0 || q != nullptr, "q is null");
    return encode_common(ctx, VRQ_CODEC_INT8_PERDOC, x, n, d, 1.0, 1.0, q, (size_t)d, mn, mx, sizeof(float), ubin, 0);
}
extern "C" int vrq_quantize_int8_global(vrq_ctx* ctx, const float* x, int64_t n, int d, double limit, int8_t* q,
                                        uint8_t* ubin) {
    VRQ_CHECK_ARG(n == 0 || q != nullptr, "q is null");
    return encode_common(ctx, VRQ_CODEC_INT8_GLOBAL, x, n, d, limit, 127.0, q, (size_t)d, nullptr, nullptr, 0, ubin, 0);
}
extern "C" int vrq_quantize_int16_global(vrq_ctx* ctx, const float* x, int64_t n, int d, double limit, int16_t* q,
                                         uint8_t* ubin) {
    VRQ_CHECK_ARG(n == 0 || q != nullptr, "q is null");
    return encode_common(ctx, VRQ_CODEC_INT16_GLOBAL, x, n, d, limit, 32767.0, q, (size_t)d * 2, nullptr, nullptr, 0, ubin, 0);
}
extern "C" int vrq_quantize_int4(vrq_ctx* ctx, const float* x, int64_t n, int d, int8_t* packed, double* mn, double* mx,
                                 uint8_t* ubin) {
    VRQ_CHECK_ARG(n == 0 || packed != nullptr, "packed is null");
    return encode_common(ctx, VRQ_CODEC_INT4, x, n, d, 1.0, 1.0, packed, (size_t)d / 2, mn, mx, sizeof(double), ubin, 0);
}
extern "C" int vrq_to_binary_f32(vrq_ctx* ctx, const float* x, int64_t n, int d, int ge, uint8_t* ubin) {
    VRQ_CHECK_ARG(n == 0 || ubin != nullptr, "ubin is null");
    return encode_common(ctx, VRQ_CODEC_NONE, x, n, d, 1.0, 1.0, nullptr, 0, nullptr, nullptr, 0, ubin, ge);
}

static int to_binary_int(vrq_ctx* ctx, const void* x, int elem, int64_t n, int d, int ge, uint8_t* ubin) {
    VRQ_TRY(check_common(ctx, x, n, d));
    VRQ_CHECK_ARG(n == 0 || ubin != nullptr, "ubin is null");
    if (n == 0) return 0;
    const void* all[2] = {x, ubin};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 2, &is_dev));
    std::vector<IoBuf> bufs;
    bufs.push_back({x, nullptr, (size_t)elem * d, {0, 0}});
    bufs.push_back({nullptr, ubin, (size_t)d / 8, {0, 0}});
    return run_rows(ctx, n, bufs, is_dev, [&](void** p, int64_t, int64_t rows, cudaStream_t st) {
        return vrq_launch_to_binary_int(ctx, p[0], elem, rows, d, ge, (uint8_t*)p[1], st);
    });
}
extern "C" int vrq_to_binary_i8(vrq_ctx* ctx, const int8_t* x, int64_t n, int d, int ge, uint8_t* ubin) {
    return to_binary_int(ctx, x, 1, n, d, ge, ubin);
}
extern "C" int vrq_to_binary_i16(vrq_ctx* ctx, const int16_t* x, int64_t n, int d, int ge, uint8_t* ubin) {
    return to_binary_int(ctx, x, 2, n, d, ge, ubin);
}

static int dequant_common(vrq_ctx* ctx, int kind, const void* q, size_t q_row_bytes, int64_t n, int d, const void* mn,
                          const void* mx, size_t stat_bytes, double limit, float* out) {
    VRQ_TRY(check_common(ctx, q, n, d));
    VRQ_CHECK_ARG(n == 0 || out != nullptr, "out is null");
    if (n == 0) return 0;
    if (kind == VRQ_PAYLOAD_INT8_PERDOC || kind == VRQ_PAYLOAD_INT4_PERDOC)
        VRQ_CHECK_ARG(mn != nullptr && mx != nullptr, "min / max arrays are required");
    const void* all[4] = {q, mn, mx, out};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 4, &is_dev));
    std::vector<IoBuf> bufs;
    bufs.push_back({q, nullptr, q_row_bytes, {0, 0}});
    int imn = -1, imx = -1;
    if (mn) { imn = (int)bufs.size(); bufs.push_back({mn, nullptr, stat_bytes, {0, 0}}); }
    if (mx) { imx = (int)bufs.size(); bufs.push_back({mx, nullptr, stat_bytes, {0, 0}}); }
    const int iout = (int)bufs.size();
    bufs.push_back({nullptr, out, sizeof(float) * (size_t)d, {0, 0}});
    return run_rows(ctx, n, bufs, is_dev, [&](void** p, int64_t, int64_t rows, cudaStream_t st) {
        vrq_dequant_args a{};
        a.kind = kind;
        a.q = p[0];
        a.n = rows;
        a.d = d;
        a.mn = imn >= 0 ? p[imn] : nullptr;
        a.mx = imx >= 0 ? p[imx] : nullptr;
        a.limit = limit;
        a.out = (float*)p[iout];
        return vrq_launch_dequant(ctx, a, st);
    });
}
extern "C" int vrq_dequantize_int8_perdoc(vrq_ctx* ctx, const int8_t* q, int64_t n, int d, const float* mn, const float* mx,
                                          float* out) {
    return dequant_common(ctx, VRQ_PAYLOAD_INT8_PERDOC, q, (size_t)d, n, d, mn, mx, sizeof(float), 0.0, out);
}
extern "C" int vrq_dequantize_int8_global(vrq_ctx* ctx, const int8_t* q, int64_t n, int d, double limit, float* out) {
    return dequant_common(ctx, VRQ_PAYLOAD_INT8_GLOBAL, q, (size_t)d, n, d, nullptr, nullptr, 0, limit, out);
}
extern "C" int vrq_dequantize_int16_global(vrq_ctx* ctx, const int16_t* q, int64_t n, int d, double limit, float* out) {
    return dequant_common(ctx, VRQ_PAYLOAD_INT16_GLOBAL, q, (size_t)d * 2, n, d, nullptr, nullptr, 0, limit, out);
}
extern "C" int vrq_dequantize_int4_perdoc(vrq_ctx* ctx, const int8_t* packed, int64_t n, int d, const double* mn,
                                          const double* mx, float* out) {
    return dequant_common(ctx, VRQ_PAYLOAD_INT4_PERDOC, packed, (size_t)d / 2, n, d, mn, mx, sizeof(double), 0.0, out);
}
extern "C" int vrq_dequantize_int4_global(vrq_ctx* ctx, const int8_t* packed, int64_t n, int d, double limit, float* out) {
    return dequant_common(ctx, VRQ_PAYLOAD_INT4_GLOBAL, packed, (size_t)d / 2, n, d, nullptr, nullptr, 0, limit, out);
}

// ---- stand-alone rescoring + synthetic data: device or host pointers ---------------------------------------------
namespace {
// Copies host arrays to scratch when needed; returns device pointers.
struct Staged {
    vrq_ctx* ctx;
    bool host;
    std::vector<std::pair<void*, std::pair<void*, size_t>>> outs;  // (host dst, (dev src, bytes))
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
        for (auto& o : outs) VRQ_CUDA(cudaMemcpyAsync(o.first, o.second.first, o.second.second, cudaMemcpyDeviceToHost, ctx->stream));
        VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
};
}  // namespace

extern "C" int vrq_rescore_binary(vrq_ctx* ctx, const uint8_t* codes, int64_t n, int d, const int64_t* pos, int64_t nq, int m,
                                  const float* q_float, double* score) {
    VRQ_CHECK_ARG(ctx && codes && pos && q_float && score, "null argument");
    VRQ_CHECK_ARG(n > 0 && nq >= 0 && m >= 0 && d > 0, "bad sizes");
    const void* all[4] = {codes, pos, q_float, score};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 4, &is_dev));
    VRQ_CUDA(cudaSetDevice(ctx->device));
    Staged s{ctx, !is_dev, {}};
    const void *dc, *dp, *dq;
    void* ds;
    VRQ_TRY(s.in(codes, (size_t)n * (d / 8), VRQ_WS_SEARCH_A, &dc));
    VRQ_TRY(s.in(pos, sizeof(int64_t) * (size_t)nq * m, VRQ_WS_SEARCH_B, &dp));
    VRQ_TRY(s.in(q_float, sizeof(float) * (size_t)nq * d, VRQ_WS_QUERY_A, &dq));
    VRQ_TRY(s.out(score, sizeof(double) * (size_t)nq * m, VRQ_WS_OUT_A, &ds));
    VRQ_TRY(vrq_launch_rescore_binary(ctx, (const uint8_t*)dc, d, nullptr, (const int64_t*)dp, 0, nq, m, (const float*)dq,
                                      (double*)ds, ctx->stream));
    return s.finish();
}

extern "C" int vrq_rescore_int8cos(vrq_ctx* ctx, const int8_t* rows, int64_t n, int d, const int64_t* pos, int64_t nq, int m,
                                   const float* q_float, double* score) {
    VRQ_CHECK_ARG(ctx && rows && pos && q_float && score, "null argument");
    VRQ_CHECK_ARG(n > 0 && nq >= 0 && m >= 0 && d > 0, "bad sizes");
    const void* all[4] = {rows, pos, q_float, score};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 4, &is_dev));
    VRQ_CUDA(cudaSetDevice(ctx->device));
    Staged s{ctx, !is_dev, {}};
    const void *dr, *dp, *dq;
    void* ds;
    VRQ_TRY(s.in(rows, (size_t)n * d, VRQ_WS_SEARCH_A, &dr));
    VRQ_TRY(s.in(pos, sizeof(int64_t) * (size_t)nq * m, VRQ_WS_SEARCH_B, &dp));
    VRQ_TRY(s.in(q_float, sizeof(float) * (size_t)nq * d, VRQ_WS_QUERY_A, &dq));
    VRQ_TRY(s.out(score, sizeof(double) * (size_t)nq * m, VRQ_WS_OUT_A, &ds));
    VRQ_TRY(vrq_launch_rescore_int8cos(ctx, (const int8_t*)dr, d, nullptr, (const int64_t*)dp, 0, nq, m, (const float*)dq,
                                       (double*)ds, ctx->stream));
    return s.finish();
}

extern "C" int vrq_synth_f32(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale, float* out) {
    VRQ_CHECK_ARG(ctx && out && nrows >= 0 && d > 0, "bad argument");
    bool is_dev;
    VRQ_TRY(vrq_is_device_ptr(out, &is_dev));
    VRQ_CUDA(cudaSetDevice(ctx->device));
    Staged s{ctx, !is_dev, {}};
    void* dv;
    VRQ_TRY(s.out(out, sizeof(float) * (size_t)nrows * d, VRQ_WS_OUT_A, &dv));
    VRQ_TRY(vrq_launch_synth_f32(ctx, seed, row0, nrows, d, row_scale, (float*)dv, ctx->stream));
    return s.finish();
}

extern "C" int vrq_synth_codes_int8(vrq_ctx* ctx, uint64_t seed, int64_t row0, int64_t nrows, int d, uint8_t* codes,
                                    int8_t* int8_rows) {
    VRQ_CHECK_ARG(ctx && (codes || int8_rows) && nrows >= 0 && d > 0 && d % 8 == 0, "bad argument");
    const void* all[2] = {codes, int8_rows};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 2, &is_dev));
    VRQ_CUDA(cudaSetDevice(ctx->device));
    Staged s{ctx, !is_dev, {}};
    void *dc, *di;
    VRQ_TRY(s.out(codes, (size_t)nrows * (d / 8), VRQ_WS_OUT_A, &dc));
    VRQ_TRY(s.out(int8_rows, (size_t)nrows * d, VRQ_WS_OUT_B, &di));
    VRQ_TRY(vrq_launch_synth_codes_int8(ctx, seed, row0, nrows, d, (uint8_t*)dc, (int8_t*)di, ctx->stream));
    return s.finish();
}

extern "C" int vrq_merge3(vrq_ctx* ctx, int world, int64_t nq, int binary_k, int64_t rank_stride, const uint64_t* keys,
                          const int64_t* labels,
                          const double* score_binary, const double* score_cosine, int k, int k2, int64_t* out_labels,
                          int32_t* out_hamming, double* out_score_binary, double* out_score_cosine, int32_t* out_count) {
    VRQ_CHECK_ARG(ctx && keys && labels && score_binary && score_cosine, "null input");
    VRQ_CHECK_ARG(out_labels && out_hamming && out_score_binary && out_score_cosine && out_count, "null output");
    const void* all[9] = {keys, labels, score_binary, score_cosine, out_labels, out_hamming, out_score_binary, out_score_cosine, out_count};
    bool is_dev;
    VRQ_TRY(vrq_space_of(all, 9, &is_dev));
    VRQ_CUDA(cudaSetDevice(ctx->device));
    Staged s{ctx, !is_dev, {}};
    if (rank_stride <= 0) rank_stride = nq * (int64_t)binary_k;
    VRQ_CHECK_ARG(rank_stride >= nq * (int64_t)binary_k, "rank_stride smaller than one rank's block");
    const size_t cnt = (size_t)(world - 1) * (size_t)rank_stride + (size_t)nq * binary_k;
    const void *dk, *dl, *db, *dc;
    void *ol, *oh, *ob, *oc, *on;
    VRQ_TRY(s.in(keys, 8 * cnt, VRQ_WS_SEARCH_A, &dk));
    VRQ_TRY(s.in(labels, 8 * cnt, VRQ_WS_SEARCH_B, &dl));
    VRQ_TRY(s.in(score_binary, 8 * cnt, VRQ_WS_SEARCH_C, &db));
    VRQ_TRY(s.in(score_cosine, 8 * cnt, VRQ_WS_SEARCH_D, &dc));
    VRQ_TRY(s.out(out_labels, 8 * (size_t)nq * k, VRQ_WS_OUT_A, &ol));
    VRQ_TRY(s.out(out_hamming, 4 * (size_t)nq * k, VRQ_WS_OUT_B, &oh));
    VRQ_TRY(s.out(out_score_binary, 8 * (size_t)nq * k, VRQ_WS_OUT_C, &ob));
    VRQ_TRY(s.out(out_score_cosine, 8 * (size_t)nq * k, VRQ_WS_OUT_D, &oc));
    VRQ_TRY(s.out(out_count, 4 * (size_t)nq, VRQ_WS_OUT_E, &on));
    VRQ_TRY(vrq_launch_merge3(ctx, world, nq, binary_k, rank_stride, (const uint64_t*)dk, (const int64_t*)dl, (const double*)db,
                              (const double*)dc, k, k2, (int64_t*)ol, (int32_t*)oh, (double*)ob, (double*)oc, (int32_t*)on,
                              ctx->stream));
    return s.finish();
}
