// exchange.cuh -- state of the peer-memory count exchange (exchange.cu), shared with the K1 launcher, whose
// prefix kernel can do the push itself (parity_common.cuh: prefix_kernel with PushArgs).
#pragma once
#include <vector>

#include "aces_common.h"

struct aces_exchange {
    int world = 0, rank = 0;
    int64_t n = 0;                // slot stride: the count rounded up to even (16-byte stores)
    int64_t nc = 0;               // counts per rank
    unsigned epoch = 0;
    std::vector<uint8_t*> base;   // [world] mapped base address of every rank's buffer (own included)
    uint8_t** d_base = nullptr;   // the same on the device
    unsigned* d_ticket = nullptr; // CTA ticket of the push kernel
    int* d_err = nullptr;         // set by a pull that timed out
    long long timeout_cycles = 20000000000ll;  // ~10 s: how long a pull waits for a peer before giving up
};

namespace xchg {
// Slot sets in rotation.  A rank pushes exchange e only after its own pull of e - 2 (so that the pull of e - 1 may
// still be waiting for SMs behind the persistent parity kernel without holding up the next repetition); that pull
// proves every peer has pushed e - 2, hence (same rule there) finished its pull of e - 4: slot e - 4 is free.
constexpr unsigned SLOTS = 4;
__host__ __device__ inline size_t slots_bytes(int world, int64_t n) { return sizeof(int64_t) * SLOTS * (size_t)world * (size_t)n; }
__host__ __device__ inline size_t flags_off(int world, int64_t n) { return (slots_bytes(world, n) + 255) & ~(size_t)255; }

// What a kernel needs to publish into slot [epoch & 1][rank] of every rank and to release the flags.
struct PushArgs {
    uint8_t* const* base = nullptr;  // NULL: no exchange attached
    int world = 0, rank = 0;
    int64_t n = 0;
    unsigned epoch = 0;
    unsigned* ticket = nullptr;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Tail of a kernel that has stored its part of the counts into the peers' slots: the last CTA of the grid
// (ticket) releases flag [epoch & 1][rank] = epoch on every rank.  Every thread of every CTA calls it.
__device__ __forceinline__ void publish_flags(const PushArgs& a, unsigned n_ctas) {
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) last = atomicAdd(a.ticket, 1u) == n_ctas - 1 ? 1u : 0u;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *a.ticket = 0;  // ready for the next push (stream order)
        __threadfence_system();
        if ((int)threadIdx.x < a.world) {
            unsigned* f = reinterpret_cast<unsigned*>(a.base[threadIdx.x] + flags_off(a.world, a.n)) +
                          (size_t)(a.epoch % SLOTS) * a.world + a.rank;
            st_release_sys(f, a.epoch);
        }
    }
}
}  // namespace xchg
