// Context, error reporting, scratch memory and CUDA-event timing for libvrq.so.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "vrq_internal.cuh"

static thread_local char g_err[512] = "";

void vrq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* vrq_last_error(void) { return g_err; }
extern "C" int vrq_version(void) { return VRQ_VERSION; }

extern "C" int vrq_ctx_create(int device, vrq_ctx** out) {
    VRQ_CHECK_ARG(out != nullptr, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        vrq_set_error("no CUDA device available (%s): libvrq has no CPU fallback", cudaGetErrorString(e));
        return e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice;
    }
    VRQ_CHECK_ARG(device >= 0 && device < count, "device index out of range");
    VRQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VRQ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        vrq_set_error("device %d is sm_%d%d; libvrq is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return VRQ_ERR_UNSUPPORTED;
    }
    vrq_ctx* c = new vrq_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    VRQ_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int i = 0; i < 2; i++) {
        VRQ_CUDA(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
        VRQ_CUDA(cudaEventCreateWithFlags(&c->pipe_ev[i], cudaEventDisableTiming));
    }
    *out = c;
    return 0;
}

extern "C" int vrq_ctx_destroy(vrq_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& b : c->ws)
        if (b.p) cudaFree(b.p);
    for (auto& t : c->timed) {
        c->ev_pool.push_back(t.e0);
        c->ev_pool.push_back(t.e1);
    }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
        if (c->pipe_ev[i]) cudaEventDestroy(c->pipe_ev[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

extern "C" int vrq_ctx_set_stream(vrq_ctx* c, void* s) {
    VRQ_CHECK_ARG(c != nullptr, "ctx is null");
    c->stream = (cudaStream_t)s;  // NULL is CUDA's legacy default stream - which is what torch's default stream is
    return 0;
}

extern "C" int vrq_ctx_reset_stream(vrq_ctx* c) {
    VRQ_CHECK_ARG(c != nullptr, "ctx is null");
    c->stream = c->own_stream;
    return 0;
}

extern "C" int vrq_ctx_sync(vrq_ctx* c) {
    VRQ_CHECK_ARG(c != nullptr, "ctx is null");
    VRQ_CUDA(cudaSetDevice(c->device));
    VRQ_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int64_t vrq_ctx_launch_count(const vrq_ctx* c) { return c ? c->launches : 0; }
extern "C" int vrq_ctx_device(const vrq_ctx* c) { return c ? c->device : -1; }

int vrq_ws_get(vrq_ctx* ctx, int slot, size_t bytes, void** out) {
    vrq_buf& b = ctx->ws[slot];
    if (b.bytes < bytes) {
        if (b.p) {
            // the old block may still be in use by enqueued work
            VRQ_CUDA(cudaStreamSynchronize(ctx->stream));
            VRQ_CUDA(cudaFree(b.p));
            b.p = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&b.p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            vrq_set_error("cudaMalloc of %zu scratch bytes failed: %s", bytes, cudaGetErrorString(e));
            return VRQ_ERR_NOMEM;
        }
        b.bytes = want;
    }
    *out = b.p;
    return 0;
}

int vrq_is_device_ptr(const void* p, bool* is_dev) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *is_dev = false;
        return 0;
    }
    *is_dev = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
    return 0;
}

int vrq_space_of(const void* const* ptrs, int n, bool* is_dev) {
    int seen = -1;
    for (int i = 0; i < n; i++) {
        if (!ptrs[i]) continue;
        bool d;
        vrq_is_device_ptr(ptrs[i], &d);
        if (seen >= 0 && (int)d != seen) {
            vrq_set_error("data pointers of one call must all be host or all be device memory");
            return VRQ_ERR_ARG;
        }
        seen = (int)d;
    }
    *is_dev = seen == 1;
    return 0;
}

// ---- timing -----------------------------------------------------------------------------------------
static cudaEvent_t take_event(vrq_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

vrq_timer_scope::vrq_timer_scope(vrq_ctx* c, int cat, cudaStream_t s) : ctx(c), st(s) {
    if (!c->timing) return;
    vrq_timed t;
    t.cat = cat;
    t.e0 = take_event(c);
    t.e1 = take_event(c);
    cudaEventRecord(t.e0, s);
    c->timed.push_back(t);
    idx = (int)c->timed.size() - 1;
}
vrq_timer_scope::~vrq_timer_scope() {
    if (idx >= 0) cudaEventRecord(ctx->timed[idx].e1, st);
}

extern "C" int vrq_ctx_enable_timing(vrq_ctx* c, int on) {
    VRQ_CHECK_ARG(c != nullptr, "ctx is null");
    c->timing = on != 0;
    return 0;
}

static int cat_of(const char* which) {
    if (!which) return -1;
    if (!strcmp(which, "scan")) return VRQ_CAT_SCAN;
    if (!strcmp(which, "encode")) return VRQ_CAT_ENCODE;
    if (!strcmp(which, "rescore")) return VRQ_CAT_RESCORE;
    if (!strcmp(which, "merge")) return VRQ_CAT_MERGE;
    if (!strcmp(which, "scan_dense")) return VRQ_CAT_SCAN_DENSE;
    return -1;
}

// Sum of device milliseconds and number of timed regions of category `which` recorded since the last query.
extern "C" double vrq_ctx_timing_ms(vrq_ctx* c, const char* which, int64_t* count) {
    if (count) *count = 0;
    if (!c) return -1.0;
    int cat = cat_of(which);
    if (cat < 0) return -1.0;
    double total = 0.0;
    int64_t n = 0;
    std::vector<vrq_timed> keep;
    for (auto& t : c->timed) {
        if (t.cat != cat) {
            keep.push_back(t);
            continue;
        }
        float ms = 0.f;
        if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            total += ms;
            n++;
        } else {
            cudaGetLastError();
        }
        c->ev_pool.push_back(t.e0);
        c->ev_pool.push_back(t.e1);
    }
    c->timed.swap(keep);
    if (count) *count = n;
    return total;
}
