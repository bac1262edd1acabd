// parity_common.cuh -- structures and PTX helpers shared by the K1 parity kernels.
#pragma once
#include "exchange.cuh"
#include "aces_common.h"

namespace k1 {


constexpr int NT = 512;       // threads per CTA
constexpr int MAX_STAGES = 4;
constexpr int HDR_BYTES = 384;  // mbarriers (64 B) + MAX_STAGES item descriptors (64 B each) + pad

struct PauliDesc {
    int32_t q[4];  // first four record bits of the support (device order)
    int32_t w;     // support size after mod-2 reduction
    int32_t ext;   // offset of the full support in `bits` (used when w > 4)
    int32_t orig;  // index of this Pauli in the caller's experiment
    int32_t pad;
};
static_assert(sizeof(PauliDesc) == 32, "PauliDesc must be 32 bytes");

struct StreamDesc {
    const uint8_t* base;  // row 0 of byte column 0 (device)
    int64_t ld;           // bytes between byte columns
    int64_t lo, hi;       // rows [lo, hi) of this budget segment
    int64_t copy_rows;    // rows rounded up to 16: limit for bulk-copy sizes
    int64_t tile_first;   // lo / TS
    int64_t out_off;      // seg[out_off + orig] accumulates this stream's counts
    int64_t pauli_off;    // first PauliDesc of this column group
    int32_t E;            // Paulis in this group
    int32_t ncols;        // byte columns the group reads (the kernel's local column 0 .. ncols-1)
    int32_t cols_off;     // their global column numbers: ParityParams::cols[cols_off .. cols_off+ncols)
    int32_t pad;
};

struct ItemDesc {  // lives in shared memory, one per pipeline stage
    // first 16 bytes: what every thread needs (one 128-bit shared load)
    int32_t stream;
    int32_t lo_rel, hi_rel;  // segment rows relative to the tile start, clamped to [0, TS]
    int32_t boundary;        // bit 0: the tile is not entirely inside the segment; bit 1: flush first
    // producer warp only
    const uint8_t* src;
    int64_t ld;
    uint32_t bytes;
    int32_t ncols, cols_off;
    uint32_t pad[5];
};
static_assert(sizeof(ItemDesc) == 64, "ItemDesc must fill its 64-byte slot (128-bit shared loads)");

struct ParityParams {
    const StreamDesc* streams;
    const int64_t* item_start;  // [n_streams + 1]
    const PauliDesc* pauli;
    const int32_t* bits;
    const int32_t* cols;        // column lists of the column groups
    unsigned long long* seg;
    int64_t total_items;
    int32_t n_streams;
    int32_t B;                  // most byte columns any stream reads (sizes the raw stages)
    int32_t NS;
    const int64_t* cta_start;   // [grid + 1] first work item of every CTA (cost-balanced), or NULL: equal counts
    int32_t grid;               // CTAs the partition was made for
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Blocks until the phase with the given parity completes.  try_wait suspends the thread in
// hardware until the phase flips or the time hint (in ns) expires, so a waiting warp does not
// burn issue slots re-polling; the loop only covers the (rare) expiry of the hint.
__device__ __forceinline__ void mbar_wait_s(uint32_t addr, uint32_t parity);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_s(smem_u32(bar), parity); }
__device__ __forceinline__ void mbar_wait_s(const uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ACES_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra ACES_DONE_%=;\n\t"
        "bra ACES_WAIT_%=;\n\t"
        "ACES_DONE_%=:\n\t}"
        ::"r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
}
// One bulk-async (TMA engine) copy global -> shared, completing `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}

// Keep bytes j of a word with lo <= j < hi.
__device__ __forceinline__ uint32_t word_byte_mask(int lo, int hi) {
    lo = max(lo, 0);
    hi = min(hi, 4);
    if (hi <= lo) return 0u;
    uint32_t mhi = hi >= 4 ? 0xFFFFFFFFu : ((1u << (8 * hi)) - 1u);
    uint32_t mlo = (1u << (8 * lo)) - 1u;
    return mhi & ~mlo;
}
// Zero the bytes of a 16-shot vector (tile rows row0 .. row0+15) that lie outside [lo, hi).
__device__ __forceinline__ uint4 mask_rows(uint4 v, int row0, int lo, int hi) {
    int l = min(max(lo - row0, 0), 16);
    int h = min(max(hi - row0, 0), 16);
    v.x &= word_byte_mask(l, h);
    v.y &= word_byte_mask(l - 4, h - 4);
    v.z &= word_byte_mask(l - 8, h - 8);
    v.w &= word_byte_mask(l - 12, h - 12);
    return v;
}

// (a & m) | (b & ~m) in one LOP3 (the compiler splits it into AND + OR-AND otherwise).
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}
// One butterfly exchange: after it `a` holds the bits of both words whose bit-index has
// (index & d) == 0 and `b` those with (index & d) != 0; m selects index bits with (index&d)==0.
// Two shifts + two LOP3 per pair, 12 pairs per 8x32-bit block.
#define ACES_BFLY(a, b, m, d)                     \
    {                                             \
        uint32_t _a = bitsel((a), (b) << (d), (m)); \
        uint32_t _b = bitsel((a) >> (d), (b), (m)); \
        (a) = _a;                                 \
        (b) = _b;                                 \
    }

__device__ __forceinline__ uint32_t popc4(const uint4 x) {
    return (__popc(x.x) + __popc(x.y)) + (__popc(x.z) + __popc(x.w));
}
__device__ __forceinline__ uint4 xor4(uint4 x, const uint4 y) {
    x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
    return x;
}
// Per-segment counts -> per-prefix (cumulative) counts; optionally added to the output.  Clears the
// segment counters it has consumed.  With an exchange attached (push.base != NULL) the same kernel IS the
// push of the multi-GPU count exchange: every count is also stored into this rank's slot of every rank's
// exchange buffer (peer memory, NVLink), and the last CTA releases the flags (exchange.cuh).
__global__ void prefix_kernel(unsigned long long* __restrict__ seg,
                              long long* __restrict__ out, const int64_t* __restrict__ item_off,
                              const int32_t* __restrict__ item_E, int n_items, int K,
                              int accumulate, const xchg::PushArgs push) {
    const size_t slot = push.base ? ((size_t)(push.epoch % xchg::SLOTS) * push.world + push.rank) * (size_t)push.n : 0;
    for (int item = blockIdx.y; item < n_items; item += gridDim.y) {
        const int E = item_E[item];
        const int64_t off = item_off[item];
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < E; p += gridDim.x * blockDim.x) {
            long long run = 0;
            for (int k = 0; k < K; ++k) {
                const int64_t idx = off + (int64_t)k * E + p;
                run += (long long)seg[idx];
                seg[idx] = 0ull;  // leave the counters clean for the next call
                out[idx] = accumulate ? out[idx] + run : run;
                if (push.base)
                    for (int r = 0; r < push.world; ++r) reinterpret_cast<long long*>(push.base[r])[slot + idx] = run;
            }
        }
    }
    if (push.base) xchg::publish_flags(push, gridDim.x * gridDim.y);
}


}  // namespace k1
