// exchange.cu -- the count exchange of the sharded K1 path written as kernels over peer memory (NVLink),
// instead of a collective library call between launches.
//
// The odd-parity counts a rank produces per step are small (0.4 - 1.4 MB of int64) and every rank needs all of
// them (repetitions: all-gather; experiments / shots of one repetition: sum).  An NCCL collective of that
// size is latency bound (60 - 100 us on eight GPUs), which is as long as the parity kernel of a sharded
// repetition itself.  Here every rank owns a buffer that all ranks of the node map (CUDA IPC):
//
//     slots [4][world][n]  int64   slot [e % 4][r] receives rank r's counts of exchange number e
//     flags [4][world]     uint32  flag [e % 4][r] = e once those counts are complete
//
//   push (compute stream, right behind the prefix kernel): a grid-stride kernel stores the rank's counts
//     into slot [e % 4][rank] of EVERY rank with 16-byte stores over NVLink; the last CTA to finish
//     (ticket) fences at system scope and releases flag [e % 4][rank] = e on every rank.
//   pull (second stream): waits (acquire, system scope) until all `world` flags of this rank show e, then
//     sums the slots (or copies them side by side) into the caller's output.
//
// Four slot sets rotate.  The host lets a rank push exchange e only after its own pull of e - 2: that pull needed
// every peer's push of e - 2, which (same rule) followed that peer's pull of e - 4 -- so nobody overwrites a slot
// that is still being read.  Integer sums: the result is exact and identical on every rank.  A pull that
// does not see its flags within its time-out (ten seconds by default) gives up and reports it (a dead peer must not hang the GPU).
#include <algorithm>
#include <cstring>
#include <vector>

#include "aces_common.h"
#include "exchange.cuh"

namespace {

using xchg::flags_off;
using xchg::st_release_sys;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counts of this rank -> slot [epoch & 1][rank] of every rank; then the flags
__global__ void __launch_bounds__(256) exchange_push_kernel(uint8_t* const* base, int world, int rank, int64_t n, int64_t nc,
                                                            unsigned epoch, const int64_t* __restrict__ local, unsigned* ticket) {
    const size_t slot = ((size_t)(epoch % xchg::SLOTS) * world + rank) * (size_t)n;  // even: 16-byte aligned
    const int64_t n2 = nc / 2;
    const longlong2* src = reinterpret_cast<const longlong2*>(local);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const longlong2 v = src[i];
        for (int p = 0; p < world; ++p) {
            longlong2* dst = reinterpret_cast<longlong2*>(base[p]) + slot / 2 + i;
            *dst = v;
        }
    }
    if ((nc & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t v = local[nc - 1];
        for (int p = 0; p < world; ++p) reinterpret_cast<int64_t*>(base[p])[slot + nc - 1] = v;
    }
    xchg::PushArgs pa;
    pa.base = base; pa.world = world; pa.rank = rank; pa.n = n; pa.epoch = epoch; pa.ticket = ticket;
    xchg::publish_flags(pa, gridDim.x);
}

// waits for the flags of all ranks, then out = sum of the slots (reduce) or the slots side by side (gather)
__global__ void __launch_bounds__(256) exchange_pull_kernel(uint8_t* mine, int world, int64_t n, int64_t nc, unsigned epoch,
                                                            int reduce, int64_t* __restrict__ out, int* err, long long timeout_cycles) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const unsigned* f = reinterpret_cast<const unsigned*>(mine + flags_off(world, n)) + (size_t)(epoch % xchg::SLOTS) * world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != epoch) {
            __nanosleep(200);
            if (clock64() - t0 > timeout_cycles) {  // a peer is gone
                ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0 && blockIdx.x == 0) *err = 1;
        return;
    }
    const int64_t* slots = reinterpret_cast<const int64_t*>(mine) + (size_t)(epoch % xchg::SLOTS) * world * (size_t)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nc; i += (int64_t)gridDim.x * blockDim.x) {
        if (reduce) {
            int64_t s = 0;
            for (int r = 0; r < world; ++r) s += slots[(size_t)r * n + i];
            out[i] = s;
        } else {
            for (int r = 0; r < world; ++r) out[(size_t)r * nc + i] = slots[(size_t)r * n + i];
        }
    }
}

}  // namespace

extern "C" {

int64_t aces_exchange_bytes(int32_t world, int64_t n_counts) {
    if (world < 1 || n_counts < 1) return 0;
    const int64_t n = (n_counts + 1) & ~(int64_t)1;
    return (int64_t)(flags_off(world, n) + sizeof(unsigned) * xchg::SLOTS * (size_t)world + 256);
}

int32_t aces_peer_alloc(aces_ctx* ctx, int64_t bytes, void** dev_ptr, uint8_t* handle_out) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, bytes > 0 && dev_ptr && handle_out, "peer_alloc: bad argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    ACES_CUDA(ctx, cudaMalloc(&p, (size_t)bytes));
    ACES_CUDA(ctx, cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return aces_fail(ctx, ACES_E_CUDA, "peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    *dev_ptr = p;
    return ACES_OK;
}

int32_t aces_peer_open(aces_ctx* ctx, const uint8_t* handle, void** dev_ptr) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, handle && dev_ptr, "peer_open: bad argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ACES_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ACES_OK;
}

int32_t aces_peer_close(aces_ctx* ctx, void* dev_ptr) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return ACES_OK;
}

int32_t aces_peer_free(aces_ctx* ctx, void* dev_ptr) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaFree(dev_ptr));
    return ACES_OK;
}

int32_t aces_exchange_create(aces_ctx* ctx, int32_t world, int32_t rank, int64_t n_counts, void* const* peer_base,
                             aces_exchange** out) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, out && peer_base && world >= 1 && world <= 64 && rank >= 0 && rank < world && n_counts >= 1,
               "exchange_create: bad argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    aces_exchange* x = new aces_exchange();
    x->world = world;
    x->rank = rank;
    x->n = (n_counts + 1) & ~(int64_t)1;
    x->nc = n_counts;
    for (int r = 0; r < world; ++r) {
        if (!peer_base[r]) {
            delete x;
            return aces_fail(ctx, ACES_E_INVALID, "exchange_create: NULL buffer of rank %d", r);
        }
        x->base.push_back((uint8_t*)peer_base[r]);
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(uint8_t*) * (size_t)world + 64);
    if (e != cudaSuccess) {
        delete x;
        return aces_fail(ctx, ACES_E_NOMEM, "exchange_create: %s", cudaGetErrorString(e));
    }
    x->d_base = (uint8_t**)p;
    x->d_ticket = (unsigned*)((uint8_t*)p + sizeof(uint8_t*) * (size_t)world);
    x->d_err = (int*)(x->d_ticket + 1);
    cudaMemset(p, 0, sizeof(uint8_t*) * (size_t)world + 64);
    cudaMemcpy(x->d_base, x->base.data(), sizeof(uint8_t*) * (size_t)world, cudaMemcpyHostToDevice);
    *out = x;
    return ACES_OK;
}

int32_t aces_exchange_destroy(aces_ctx* ctx, aces_exchange* x) {
    if (!ctx) return ACES_E_INVALID;
    if (!x) return ACES_OK;
    if (ctx->xch == x) ctx->xch = nullptr;
    cudaSetDevice(ctx->device);
    if (x->d_base) cudaFree(x->d_base);
    delete x;
    return ACES_OK;
}

int32_t aces_exchange_push(aces_ctx* ctx, aces_exchange* x, const int64_t* local_counts, void* cuda_stream) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, x && local_counts, "exchange_push: bad argument");
    ACES_CHECK(ctx, ((uintptr_t)local_counts & 15) == 0, "exchange_push: the counts must be 16-byte aligned");
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    x->epoch += 1;
    const int grid = (int)std::min<int64_t>((x->n / 2 + 255) / 256, 2 * (int64_t)ctx->sm_count);
    exchange_push_kernel<<<std::max(grid, 1), 256, 0, st>>>(x->d_base, x->world, x->rank, x->n, x->nc, x->epoch, local_counts,
                                                            x->d_ticket);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

int32_t aces_exchange_pull(aces_ctx* ctx, aces_exchange* x, int32_t reduce, int64_t* out, void* cuda_stream) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, x && out && x->epoch > 0, "exchange_pull: bad argument (push first)");
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int grid = (int)std::min<int64_t>((x->nc + 255) / 256, (int64_t)ctx->sm_count);
    exchange_pull_kernel<<<std::max(grid, 1), 256, 0, st>>>(x->base[(size_t)x->rank], x->world, x->n, x->nc, x->epoch, reduce ? 1 : 0,
                                                            out, x->d_err, x->timeout_cycles);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

int32_t aces_exchange_set_timeout(aces_ctx* ctx, aces_exchange* x, int64_t milliseconds) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, x && milliseconds > 0, "exchange_set_timeout: bad argument");
    x->timeout_cycles = (long long)milliseconds * 2000000ll;  // SM clock of about 2 GHz
    return ACES_OK;
}

int32_t aces_exchange_attach(aces_ctx* ctx, aces_exchange* x) {
    if (!ctx) return ACES_E_INVALID;
    ctx->xch = x;  // NULL detaches
    return ACES_OK;
}

int32_t aces_exchange_status(aces_ctx* ctx, aces_exchange* x) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, x != nullptr, "exchange_status: NULL exchange");
    int err = 0;
    ACES_CUDA(ctx, cudaMemcpy(&err, x->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return aces_fail(ctx, ACES_E_CUDA, "exchange: a pull timed out waiting for a peer's counts");
    return ACES_OK;
}

}  // extern "C"
