// context.cu — context lifetime, error reporting and buffer management of libaces_b200.so.
#include "aces_common.h"

static std::string g_init_err;

int aces_fail(aces_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_init_err = buf;
    return code;
}

int aces_reserve_dev(aces_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return ACES_OK;
    if (b.ptr) {
        // The buffer may still be read by work queued on the stream.
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ACES_CUDA(ctx, cudaFree(b.ptr));
        b.ptr = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.ptr, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.ptr = nullptr;
        return aces_fail(ctx, ACES_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes,
                         cudaGetErrorString(e));
    }
    b.cap = want;
    return ACES_OK;
}

int aces_reserve_host(aces_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return ACES_OK;
    if (b.ptr) {
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ACES_CUDA(ctx, cudaFreeHost(b.ptr));
        b.ptr = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&b.ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.ptr = nullptr;
        return aces_fail(ctx, ACES_E_NOMEM, "cudaMallocHost of %zu bytes failed: %s", bytes,
                         cudaGetErrorString(e));
    }
    b.cap = want;
    return ACES_OK;
}

bool aces_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" {

int32_t aces_version(void) { return ACES_VERSION; }

int32_t aces_init(int32_t device, aces_ctx** out) {
    if (!out) return aces_fail(nullptr, ACES_E_INVALID, "aces_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return aces_fail(nullptr, ACES_E_CUDA,
                         "aces_init: no CUDA device available (%s); libaces_b200 has no CPU path",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count)
        return aces_fail(nullptr, ACES_E_INVALID, "aces_init: device %d out of range [0,%d)",
                         device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return aces_fail(nullptr, ACES_E_CUDA, "cudaGetDeviceProperties: %s",
                         cudaGetErrorString(e));
    if (prop.major != 10)
        return aces_fail(nullptr, ACES_E_UNSUPPORTED,
                         "aces_init: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return aces_fail(nullptr, ACES_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    aces_ctx* ctx = new aces_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithPriority(&ctx->aux2_stream, cudaStreamNonBlocking, (prio_lo + prio_hi) / 2);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_crit, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_aux2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->t0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->t1);
    if (e != cudaSuccess) {
        int rc = aces_fail(nullptr, ACES_E_CUDA, "stream/event creation: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return ACES_OK;
}

int32_t aces_destroy(aces_ctx* ctx) {
    if (!ctx) return ACES_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->aux2_stream) cudaStreamSynchronize(ctx->aux2_stream);
    DevBuf* dev[] = {&ctx->d_arena, &ctx->d_raw, &ctx->d_tdesc, &ctx->d_dag_tasks, &ctx->d_dag_state, &ctx->d_dag_trace, &ctx->d_streams, &ctx->d_seg, &ctx->d_odd, &ctx->d_flags, &ctx->d_eig, &ctx->d_eigA};
    for (DevBuf* b : dev)
        if (b->ptr) cudaFree(b->ptr);
    for (DevBuf& b : ctx->d_work)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->h_stage.ptr) cudaFreeHost(ctx->h_stage.ptr);
    if (ctx->h_stage2.ptr) cudaFreeHost(ctx->h_stage2.ptr);
    if (ctx->h_pack.ptr) cudaFreeHost(ctx->h_pack.ptr);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev) cudaEventDestroy(ctx->ev);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_crit) cudaEventDestroy(ctx->ev_crit);
    if (ctx->ev_aux2) cudaEventDestroy(ctx->ev_aux2);
    if (ctx->aux2_stream) cudaStreamDestroy(ctx->aux2_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return ACES_OK;
}

const char* aces_last_error(const aces_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_init_err.c_str();
}

int32_t aces_set_stream(aces_ctx* ctx, void* cuda_stream) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return ACES_OK;
}

int32_t aces_synchronize(aces_ctx* ctx) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}

int32_t aces_host_alloc(aces_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1));
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return aces_fail(ctx, ACES_E_NOMEM, "cudaMallocHost(%lld): %s", (long long)bytes,
                         cudaGetErrorString(e));
    }
    return ACES_OK;
}

int32_t aces_host_free(aces_ctx* ctx, void* ptr) {
    if (!ctx) return ACES_E_INVALID;
    if (ptr) ACES_CUDA(ctx, cudaFreeHost(ptr));
    return ACES_OK;
}

int64_t aces_kernel_launches(const aces_ctx* ctx) { return ctx ? ctx->launches : 0; }
int32_t aces_last_rank_deficiency(const aces_ctx* ctx) { return ctx ? ctx->rank_deficiency : -1; }

int32_t aces_set_timing(aces_ctx* ctx, int32_t on) {
    if (!ctx) return ACES_E_INVALID;
    ctx->timing = on ? 1 : 0;
    ctx->timed = 0;
    return ACES_OK;
}

int32_t aces_last_kernel_ms(aces_ctx* ctx, float* ms) {
    if (!ctx || !ms) return ACES_E_INVALID;
    ACES_CHECK(ctx, ctx->timed, "last_kernel_ms: no timed kernel (call aces_set_timing(ctx, 1) first)");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    ACES_CUDA(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return ACES_OK;
}

}  // extern "C"
