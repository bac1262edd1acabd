// parity.cu — K1: bit-packed shot records -> odd-parity counts per (experiment, Pauli, budget).
//
// Replaces pauli_estimate / estimate_experiment (reference src/estimate.jl:239-301) and the
// per-budget loop of simulate_stim_estimate (src/simulate.jl:195-234).
//
// Data layout (HBM): each sample matrix is column-major uint8 [rows x B] exactly as the Julia
// Matrix{UInt8} Stim returns (src/simulate.jl:60-63): byte column c holds record bits 8c..8c+7
// of every shot, shots contiguous.  The kernel never reshapes it in HBM; it is read once.
//
// Kernel (persistent, one launch per call covering every experiment of the batch):
//   work item  = one TS-shot tile of one (sample matrix, budget segment) "stream";
//   stage      = B bulk-async copies (cp.async.bulk, the TMA engine; SASS UBLKCP) of TS bytes,
//                one per byte column, into a multi-stage shared-memory ring, completion on an
//                mbarrier;
//   phase A    = every thread takes 8 words (32 shots) of one byte column and runs a 3-round
//                masked-shift butterfly that turns 32 shots x 8 bits into 8 bit-plane words
//                (one record bit across 32 shots each); planes go to shared memory;
//   phase B    = one thread per measured Pauli: XOR of the planes in its support (128-bit
//                shared loads, 128 shots per load), popc, 32-bit running count in a register;
//   flush      = when the CTA moves to another stream: one 64-bit atomic per Pauli.
// Budget prefixes (nested shot counts over the same matrix) become disjoint row segments;
// a tile that straddles a segment end is processed once per segment with the bytes outside
// the segment zeroed (a zero record has even parity for every Pauli, so it adds nothing to an
// odd count).  A second tiny kernel turns per-segment counts into per-prefix counts.
// Counts are exact integers: the result is bit-identical to the reference by construction.
#include <algorithm>
#include <cstring>

#include "aces_common.h"

namespace {

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
    int64_t pauli_off;    // first PauliDesc of this slice
    int32_t E;            // Paulis in this slice
    int32_t pad;
};

struct ItemDesc {  // lives in shared memory, one per pipeline stage
    int32_t stream;
    int32_t boundary;
    int64_t r0, lo, hi;
    const uint8_t* src;
    int64_t ld;
    uint32_t bytes;
    uint32_t pad;
};
static_assert(sizeof(ItemDesc) <= 64, "ItemDesc must fit its 64-byte slot");

struct ParityParams {
    const StreamDesc* streams;
    const int64_t* item_start;  // [n_streams + 1]
    const PauliDesc* pauli;
    const int32_t* bits;
    unsigned long long* seg;
    int64_t total_items;
    int32_t n_streams;
    int32_t B;
    int32_t NS;
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
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
// Zero the bytes of a 16-shot vector (rows row0 .. row0+15) that lie outside [lo, hi).
__device__ __forceinline__ uint4 mask_rows(uint4 v, int64_t row0, int64_t lo, int64_t hi) {
    int l = (int)min(max(lo - row0, (int64_t)0), (int64_t)16);
    int h = (int)min(max(hi - row0, (int64_t)0), (int64_t)16);
    v.x &= word_byte_mask(l, h);
    v.y &= word_byte_mask(l - 4, h - 4);
    v.z &= word_byte_mask(l - 8, h - 8);
    v.w &= word_byte_mask(l - 12, h - 12);
    return v;
}

// One butterfly exchange: after it `a` holds the bits of both words whose bit-index has
// (index & d) == 0 and `b` those with (index & d) != 0; m selects index bits with (index&d)==0.
#define ACES_BFLY(a, b, m, d)                              \
    {                                                      \
        uint32_t _a = ((a) & (m)) | (((b) << (d)) & ~(m)); \
        uint32_t _b = (((a) >> (d)) & (m)) | ((b) & ~(m)); \
        (a) = _a;                                          \
        (b) = _b;                                          \
    }

template <int TS, int PPT>
__global__ void __launch_bounds__(NT) parity_kernel(const ParityParams P) {
    constexpr int UPC = TS / 32;        // phase-A units (32 shots) per byte column
    constexpr int PSTRIDE = UPC + 4;    // words per bit-plane row (padded: conflict-free LDS.128)
    constexpr int NV = TS / 128;        // 128-shot vectors per plane per tile
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    ItemDesc* sdesc = reinterpret_cast<ItemDesc*>(smem + 64);
    uint32_t* planes = reinterpret_cast<uint32_t*>(smem + HDR_BYTES);
    const int B = P.B;
    const int NS = P.NS;
    const uint32_t plane_bytes = (uint32_t)(8 * B * PSTRIDE * 4);
    const uint32_t raw_off = (HDR_BYTES + plane_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = (uint32_t)(B * TS);
    uint8_t* raw = smem + raw_off;
    const uint32_t planes_s = smem_u32(planes);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    const int64_t i0 = (P.total_items * (int64_t)blockIdx.x) / gridDim.x;
    const int64_t i1 = (P.total_items * (int64_t)(blockIdx.x + 1)) / gridDim.x;
    if (i0 >= i1) return;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    // ---- producer (warp 0): locate item, describe it, launch its B column copies ------------
    int prod_stream = 0;  // cached stream index of the last issued item (lane 0 only)
    auto issue = [&](int64_t item, int stage) {
        ItemDesc* d = &sdesc[stage];
        if (lane == 0) {
            int s = prod_stream;
            if (!(P.item_start[s] <= item && item < P.item_start[s + 1])) {
                int a = 0, b = P.n_streams;
                while (b - a > 1) {
                    int m = (a + b) >> 1;
                    if (P.item_start[m] <= item) a = m; else b = m;
                }
                s = a;
                prod_stream = s;
            }
            const StreamDesc sd = P.streams[s];
            const int64_t r0 = (sd.tile_first + (item - P.item_start[s])) * TS;
            const int64_t avail = sd.copy_rows - r0;
            const uint32_t bytes = (uint32_t)(avail < TS ? avail : TS);
            d->stream = s;
            d->r0 = r0;
            d->lo = sd.lo > r0 ? sd.lo : r0;
            d->hi = sd.hi < r0 + TS ? sd.hi : r0 + TS;
            d->boundary = (sd.lo > r0) || (sd.hi < r0 + TS);
            d->src = sd.base + r0;
            d->ld = sd.ld;
            d->bytes = bytes;
            mbar_arrive_expect_tx(&bars[stage], bytes * (uint32_t)B);
        }
        __syncwarp();
        const uint8_t* src = d->src;
        const int64_t ld = d->ld;
        const uint32_t bytes = d->bytes;
        uint8_t* dst = raw + (size_t)stage * stage_bytes;
        for (int c = lane; c < B; c += 32)
            bulk_g2s(dst + c * TS, src + (int64_t)c * ld, bytes, &bars[stage], policy);
    };
    if (warp == 0) {
        for (int s = 0; s < NS; ++s)
            if (i0 + s < i1) issue(i0 + s, s);
    }

    // ---- per-thread Pauli state ---------------------------------------------------------------
    uint32_t addr[PPT][4];
    int32_t pw[PPT], pext[PPT], porig[PPT];
    uint32_t acc[PPT];
#pragma unroll
    for (int ch = 0; ch < PPT; ++ch) { pw[ch] = 0; acc[ch] = 0; pext[ch] = 0; porig[ch] = 0; }
    int cur_stream = -1;
    int64_t cur_out = 0;
    int g = 0, G = 1;  // this thread's shot-vector group and the number of groups
    uint32_t tiles_since_flush = 0;

    auto flush = [&]() {
#pragma unroll
        for (int ch = 0; ch < PPT; ++ch) {
            if (pw[ch] > 0 && acc[ch] != 0)
                atomicAdd(&P.seg[cur_out + porig[ch]], (unsigned long long)acc[ch]);
            acc[ch] = 0;
        }
        tiles_since_flush = 0;
    };

    for (int64_t it = i0; it < i1; ++it) {
        const int64_t rel = it - i0;
        const int stage = (int)(rel % NS);
        const uint32_t parity = (uint32_t)((rel / NS) & 1);
        mbar_wait(&bars[stage], parity);
        const ItemDesc d = sdesc[stage];

        if (d.stream != cur_stream || tiles_since_flush >= (0x7FFFFFFFu / TS)) {
            if (cur_stream >= 0) flush();
            if (d.stream != cur_stream) {
                cur_stream = d.stream;
                const StreamDesc* sd = &P.streams[cur_stream];
                cur_out = sd->out_off;
                const int E = sd->E;
                const int64_t poff = sd->pauli_off;
                if (PPT == 1) {
                    const int EP = (E + 31) & ~31;
                    G = EP > 0 ? NT / EP : 1;
                    if (G > NV) G = NV;
                    g = EP > 0 ? tid / EP : 0;
                }
#pragma unroll
                for (int ch = 0; ch < PPT; ++ch) {
                    int p;
                    bool on;
                    if (PPT == 1) {
                        const int EP = (E + 31) & ~31;
                        p = EP > 0 ? tid % EP : 0;
                        on = (g < G) && (p < E);
                    } else {
                        p = ch * NT + tid;
                        on = p < E;
                    }
                    pw[ch] = 0;
                    if (on) {
                        const uint4* pd4 = reinterpret_cast<const uint4*>(&P.pauli[poff + p]);
                        const uint4 qa = __ldg(pd4);
                        const uint4 qb = __ldg(pd4 + 1);
                        addr[ch][0] = planes_s + qa.x * (PSTRIDE * 4);
                        addr[ch][1] = planes_s + qa.y * (PSTRIDE * 4);
                        addr[ch][2] = planes_s + qa.z * (PSTRIDE * 4);
                        addr[ch][3] = planes_s + qa.w * (PSTRIDE * 4);
                        pw[ch] = (int32_t)qb.x;
                        pext[ch] = (int32_t)qb.y;
                        porig[ch] = (int32_t)qb.z;
                    }
                }
            }
        }
        ++tiles_since_flush;

        // ---- phase A: raw byte columns -> bit planes ------------------------------------------
        {
            const uint8_t* st = raw + (size_t)stage * stage_bytes;
            const int nunits = B * UPC;
            for (int unit = tid; unit < nunits; unit += NT) {
                const int c = unit / UPC;
                const int u = unit - c * UPC;
                const uint8_t* col = st + c * TS;
                uint4 a = *reinterpret_cast<const uint4*>(col + 16 * u);
                uint4 b = *reinterpret_cast<const uint4*>(col + TS / 2 + 16 * u);
                if (d.boundary) {
                    a = mask_rows(a, d.r0 + 16 * u, d.lo, d.hi);
                    b = mask_rows(b, d.r0 + TS / 2 + 16 * u, d.lo, d.hi);
                }
                uint32_t w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w;
                uint32_t w4 = b.x, w5 = b.y, w6 = b.z, w7 = b.w;
                ACES_BFLY(w0, w1, 0x55555555u, 1) ACES_BFLY(w2, w3, 0x55555555u, 1)
                ACES_BFLY(w4, w5, 0x55555555u, 1) ACES_BFLY(w6, w7, 0x55555555u, 1)
                ACES_BFLY(w0, w2, 0x33333333u, 2) ACES_BFLY(w1, w3, 0x33333333u, 2)
                ACES_BFLY(w4, w6, 0x33333333u, 2) ACES_BFLY(w5, w7, 0x33333333u, 2)
                ACES_BFLY(w0, w4, 0x0F0F0F0Fu, 4) ACES_BFLY(w1, w5, 0x0F0F0F0Fu, 4)
                ACES_BFLY(w2, w6, 0x0F0F0F0Fu, 4) ACES_BFLY(w3, w7, 0x0F0F0F0Fu, 4)
                // w_k now holds record bit 8c+k of the unit's 32 shots.
                uint32_t* pl = planes + (8 * c) * PSTRIDE + u;
                pl[0 * PSTRIDE] = w0; pl[1 * PSTRIDE] = w1;
                pl[2 * PSTRIDE] = w2; pl[3 * PSTRIDE] = w3;
                pl[4 * PSTRIDE] = w4; pl[5 * PSTRIDE] = w5;
                pl[6 * PSTRIDE] = w6; pl[7 * PSTRIDE] = w7;
            }
        }
        __syncthreads();  // planes complete; this raw stage is free again
        if (warp == 0 && it + NS < i1) issue(it + NS, stage);

        // ---- phase B: per-Pauli XOR of planes + popcount --------------------------------------
#pragma unroll
        for (int ch = 0; ch < PPT; ++ch) {
            const int w = pw[ch];
            if (w > 0) {
                uint32_t cnt = 0;
                for (int v = g; v < NV; v += G) {
                    const uint32_t vo = (uint32_t)v * 16u;
                    uint4 x = lds128(addr[ch][0] + vo);
                    if (w > 1) {
                        const uint4 y = lds128(addr[ch][1] + vo);
                        x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
                    }
                    if (w > 2) {
                        const uint4 y = lds128(addr[ch][2] + vo);
                        x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
                    }
                    if (w > 3) {
                        const uint4 y = lds128(addr[ch][3] + vo);
                        x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
                    }
                    for (int i = 4; i < w; ++i) {
                        const uint32_t q = (uint32_t)__ldg(&P.bits[pext[ch] + i]);
                        const uint4 y = lds128(planes_s + q * (PSTRIDE * 4) + vo);
                        x.x ^= y.x; x.y ^= y.y; x.z ^= y.z; x.w ^= y.w;
                    }
                    cnt += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
                }
                acc[ch] += cnt;
            }
        }
        __syncthreads();  // planes may be overwritten by the next item
    }
    if (cur_stream >= 0) flush();
}

// Per-segment counts -> per-prefix (cumulative) counts; optionally added to the output.
__global__ void prefix_kernel(const unsigned long long* __restrict__ seg,
                              long long* __restrict__ out, const int64_t* __restrict__ item_off,
                              const int32_t* __restrict__ item_E, int n_items, int K,
                              int accumulate) {
    for (int item = blockIdx.y; item < n_items; item += gridDim.y) {
        const int E = item_E[item];
        const int64_t off = item_off[item];
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < E; p += gridDim.x * blockDim.x) {
            long long run = 0;
            for (int k = 0; k < K; ++k) {
                const int64_t idx = off + (int64_t)k * E + p;
                run += (long long)seg[idx];
                out[idx] = accumulate ? out[idx] + run : run;
            }
        }
    }
}

struct TileCfg {
    int TS;
    int NS;
    size_t smem;
};

// Picks the tile length for B byte columns so that at least two stages fit in shared memory
// (and two CTAs per SM where possible).
bool pick_tile(int B, int smem_optin, TileCfg* cfg) {
    const int candidates[3] = {4096, 1024, 256};
    for (int i = 0; i < 3; ++i) {
        const int TS = candidates[i];
        if (TS == 4096 && B > 6) continue;
        if (TS == 1024 && B > 40) continue;
        const size_t planes = (size_t)8 * B * (TS / 32 + 4) * 4;
        const size_t raw_off = (HDR_BYTES + planes + 127) & ~(size_t)127;
        const size_t stage = (size_t)B * TS;
        // two CTAs per SM: each CTA also reserves 1 KB of the SM's 228 KB
        const size_t half = ((size_t)smem_optin + 1024) / 2 - 1024;
        int ns = 0;
        for (int s = MAX_STAGES; s >= 2; --s)
            if (raw_off + s * stage <= half) { ns = s; break; }
        if (ns == 0)
            for (int s = MAX_STAGES; s >= 2; --s)
                if (raw_off + s * stage <= (size_t)smem_optin) { ns = s; break; }
        if (ns == 0) continue;
        cfg->TS = TS;
        cfg->NS = ns;
        cfg->smem = raw_off + ns * stage;
        return true;
    }
    return false;
}

template <int TS, int PPT>
int launch_parity(aces_ctx* ctx, const ParityParams& P, const TileCfg& cfg) {
    auto kern = parity_kernel<TS, PPT>;
    ACES_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)cfg.smem));
    int per_sm = 0;
    ACES_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, cfg.smem));
    if (per_sm < 1)
        return aces_fail(ctx, ACES_E_UNSUPPORTED, "parity kernel does not fit on an SM (smem %zu)",
                         cfg.smem);
    int64_t grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > P.total_items) grid = P.total_items;
    if (ctx->timing) ACES_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    kern<<<(unsigned)grid, NT, cfg.smem, ctx->stream>>>(P);
    ACES_CUDA(ctx, cudaGetLastError());
    if (ctx->timing) {
        ACES_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
        ctx->timed = 1;
    }
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace

struct aces_tables {
    int32_t n_qubits = 0;
    int32_t B = 0;
    int32_t n_experiments = 0;
    int32_t max_E = 0;
    std::vector<int64_t> exp_ptr;  // host copy: Paulis of experiment e
    PauliDesc* d_pauli = nullptr;
    int32_t* d_bits = nullptr;
};

static int parity_run(aces_ctx* ctx, const aces_tables* t, int32_t n_items,
                      const int32_t* experiment_ids, const uint8_t* const* counts,
                      const int64_t* rows, const int64_t* ld, int32_t K,
                      const int64_t* prefix_shots, int64_t* odd_counts, int32_t accumulate,
                      bool async_only) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, t != nullptr, "parity_counts: tables is NULL");
    ACES_CHECK(ctx, n_items >= 0 && K >= 1, "parity_counts: n_items=%d K=%d", n_items, K);
    if (n_items == 0) return ACES_OK;
    ACES_CHECK(ctx, experiment_ids && counts && rows && ld && prefix_shots && odd_counts,
               "parity_counts: NULL argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = t->B;
    TileCfg cfg;
    if (!pick_tile(B, ctx->smem_optin, &cfg))
        return aces_fail(ctx, ACES_E_UNSUPPORTED,
                         "parity_counts: %d byte columns per shot exceed the shared-memory tiling", B);
    const int TS = cfg.TS;

    // ---- validate, lay out outputs, decide which matrices need a pitched device copy ----------
    std::vector<int64_t> item_off(n_items + 1, 0);
    std::vector<int32_t> item_E(n_items);
    std::vector<char> in_place(n_items);
    std::vector<size_t> arena_off(n_items, 0);
    size_t arena_bytes = 0;
    int max_E = 0;
    for (int i = 0; i < n_items; ++i) {
        const int e = experiment_ids[i];
        ACES_CHECK(ctx, e >= 0 && e < t->n_experiments, "parity_counts: item %d: experiment %d out of range", i, e);
        ACES_CHECK(ctx, rows[i] >= 0, "parity_counts: item %d: rows=%lld", i, (long long)rows[i]);
        ACES_CHECK(ctx, ld[i] >= rows[i], "parity_counts: item %d: ld=%lld < rows=%lld", i,
                   (long long)ld[i], (long long)rows[i]);
        ACES_CHECK(ctx, rows[i] == 0 || counts[i] != nullptr, "parity_counts: item %d: NULL matrix", i);
        int64_t prev = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t s = prefix_shots[(int64_t)i * K + k];
            ACES_CHECK(ctx, s >= prev && s <= rows[i],
                       "parity_counts: item %d: prefix_shots[%d]=%lld not in [%lld,%lld]", i, k,
                       (long long)s, (long long)prev, (long long)rows[i]);
            prev = s;
        }
        const int E = (int)(t->exp_ptr[e + 1] - t->exp_ptr[e]);
        item_E[i] = E;
        max_E = std::max(max_E, E);
        item_off[i + 1] = item_off[i] + (int64_t)E * K;
        const bool dev = rows[i] > 0 && aces_is_device_ptr(counts[i]);
        if (async_only)
            ACES_CHECK(ctx, rows[i] == 0 || dev, "parity_counts_async: item %d is not a device matrix", i);
        in_place[i] = dev && (((uintptr_t)counts[i]) % 16 == 0) && (ld[i] % 16 == 0);
        if (!in_place[i] && rows[i] > 0) {
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            arena_off[i] = arena_bytes;
            arena_bytes += pitch * B;
        }
    }
    const int64_t total_out = item_off[n_items];
    if (total_out == 0) return ACES_OK;
    const bool out_dev = aces_is_device_ptr(odd_counts);
    if (async_only) ACES_CHECK(ctx, out_dev, "parity_counts_async: odd_counts must be a device pointer");

    // ---- stream table: one entry per (item, non-empty budget segment, Pauli slice) -------------
    const int slice = NT * 4;  // Paulis one CTA pass can hold in registers (PPT <= 4)
    std::vector<StreamDesc> streams;
    std::vector<int64_t> item_start(1, 0);
    if (arena_bytes > 0) ACES_TRY(aces_reserve_dev(ctx, ctx->d_arena, arena_bytes));
    for (int i = 0; i < n_items; ++i) {
        if (rows[i] == 0 || item_E[i] == 0) continue;
        const int e = experiment_ids[i];
        const uint8_t* base;
        int64_t ldd;
        if (in_place[i]) {
            base = counts[i];
            ldd = ld[i];
        } else {
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            uint8_t* dst = (uint8_t*)ctx->d_arena.ptr + arena_off[i];
            ACES_CUDA(ctx, cudaMemcpy2DAsync(dst, pitch, counts[i], (size_t)ld[i], (size_t)rows[i],
                                             (size_t)B, cudaMemcpyDefault, ctx->stream));
            base = dst;
            ldd = (int64_t)pitch;
        }
        int64_t prev = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t s = prefix_shots[(int64_t)i * K + k];
            if (s > prev) {
                for (int p0 = 0; p0 < item_E[i]; p0 += slice) {
                    StreamDesc sd;
                    sd.base = base;
                    sd.ld = ldd;
                    sd.lo = prev;
                    sd.hi = s;
                    sd.copy_rows = (rows[i] + 15) & ~(int64_t)15;
                    sd.tile_first = prev / TS;
                    sd.out_off = item_off[i] + (int64_t)k * item_E[i];
                    sd.pauli_off = t->exp_ptr[e] + p0;
                    sd.E = std::min(slice, item_E[i] - p0);
                    sd.pad = 0;
                    const int64_t ntiles = (s - 1) / TS - prev / TS + 1;
                    streams.push_back(sd);
                    item_start.push_back(item_start.back() + ntiles);
                }
            }
            prev = s;
        }
    }
    const int n_streams = (int)streams.size();
    const int64_t total_items = item_start.back();

    // ---- upload the small tables through pinned staging -----------------------------------------
    const size_t sz_streams = sizeof(StreamDesc) * (size_t)std::max(n_streams, 1);
    const size_t sz_istart = sizeof(int64_t) * (size_t)(n_streams + 1);
    const size_t sz_ioff = sizeof(int64_t) * (size_t)(n_items + 1);
    const size_t sz_iE = sizeof(int32_t) * (size_t)n_items;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_streams = 0, o_istart = al(sz_streams), o_ioff = o_istart + al(sz_istart),
                 o_iE = o_ioff + al(sz_ioff), tbl_bytes = o_iE + al(sz_iE);
    // The pinned staging buffer is reused across calls: wait for earlier async work first.
    ACES_CUDA(ctx, cudaEventSynchronize(ctx->ev));
    ACES_TRY(aces_reserve_host(ctx, ctx->h_stage, tbl_bytes));
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_streams, tbl_bytes));
    uint8_t* hs = (uint8_t*)ctx->h_stage.ptr;
    if (n_streams) memcpy(hs + o_streams, streams.data(), sizeof(StreamDesc) * n_streams);
    memcpy(hs + o_istart, item_start.data(), sz_istart);
    memcpy(hs + o_ioff, item_off.data(), sz_ioff);
    memcpy(hs + o_iE, item_E.data(), sz_iE);
    uint8_t* ds = (uint8_t*)ctx->d_streams.ptr;
    ACES_CUDA(ctx, cudaMemcpyAsync(ds, hs, tbl_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaEventRecord(ctx->ev, ctx->stream));

    ACES_TRY(aces_reserve_dev(ctx, ctx->d_seg, sizeof(unsigned long long) * (size_t)total_out));
    ACES_CUDA(ctx, cudaMemsetAsync(ctx->d_seg.ptr, 0, sizeof(unsigned long long) * (size_t)total_out,
                                   ctx->stream));

    if (total_items > 0) {
        ParityParams P;
        P.streams = (const StreamDesc*)(ds + o_streams);
        P.item_start = (const int64_t*)(ds + o_istart);
        P.pauli = t->d_pauli;
        P.bits = t->d_bits;
        P.seg = (unsigned long long*)ctx->d_seg.ptr;
        P.total_items = total_items;
        P.n_streams = n_streams;
        P.B = B;
        P.NS = cfg.NS;
        const int ppt = max_E <= NT ? 1 : (max_E <= 2 * NT ? 2 : 4);
        int rc;
#define ACES_LAUNCH(TSV)                                              \
    (ppt == 1 ? launch_parity<TSV, 1>(ctx, P, cfg)                    \
              : ppt == 2 ? launch_parity<TSV, 2>(ctx, P, cfg) : launch_parity<TSV, 4>(ctx, P, cfg))
        if (TS == 4096) rc = ACES_LAUNCH(4096);
        else if (TS == 1024) rc = ACES_LAUNCH(1024);
        else rc = ACES_LAUNCH(256);
#undef ACES_LAUNCH
        ACES_TRY(rc);
    }

    // ---- segment counts -> prefix counts ----------------------------------------------------------
    long long* d_out;
    int acc_dev = 0;
    if (out_dev) {
        d_out = (long long*)odd_counts;
        acc_dev = accumulate ? 1 : 0;
    } else {
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_odd, sizeof(long long) * (size_t)total_out));
        d_out = (long long*)ctx->d_odd.ptr;
    }
    {
        dim3 grid((unsigned)std::min<int64_t>(((int64_t)max_E + 255) / 256, 64), (unsigned)std::min(n_items, 16384));
        prefix_kernel<<<grid, 256, 0, ctx->stream>>>((const unsigned long long*)ctx->d_seg.ptr, d_out,
                                                    (const int64_t*)(ds + o_ioff),
                                                    (const int32_t*)(ds + o_iE), n_items, K, acc_dev);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    if (async_only) return ACES_OK;
    if (out_dev) {
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return ACES_OK;
    }
    if (accumulate) {
        std::vector<int64_t> tmp((size_t)total_out);
        ACES_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_out, sizeof(int64_t) * (size_t)total_out,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t j = 0; j < total_out; ++j) odd_counts[j] += tmp[(size_t)j];
    } else {
        ACES_CUDA(ctx, cudaMemcpyAsync(odd_counts, d_out, sizeof(int64_t) * (size_t)total_out,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ACES_OK;
}

extern "C" {

int32_t aces_tables_create(aces_ctx* ctx, int32_t n_qubits, int32_t n_experiments,
                           const int64_t* exp_ptr, const int64_t* pauli_ptr,
                           const int32_t* pauli_bits, aces_tables** out) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, out != nullptr, "tables_create: out is NULL");
    *out = nullptr;
    ACES_CHECK(ctx, n_qubits >= 1 && n_experiments >= 0, "tables_create: n_qubits=%d n_experiments=%d",
               n_qubits, n_experiments);
    ACES_CHECK(ctx, exp_ptr && pauli_ptr, "tables_create: NULL table");
    ACES_CHECK(ctx, exp_ptr[0] == 0 && pauli_ptr[0] == 0, "tables_create: offsets must start at 0");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t P = exp_ptr[n_experiments];
    std::vector<PauliDesc> descs((size_t)P);
    std::vector<int32_t> bits;
    bits.reserve((size_t)pauli_ptr[P] + 4);
    int max_E = 0;
    std::vector<int32_t> sup;
    std::vector<std::pair<int32_t, int32_t>> order;  // (weight, original index)
    std::vector<std::vector<int32_t>> sups;
    for (int e = 0; e < n_experiments; ++e) {
        ACES_CHECK(ctx, exp_ptr[e + 1] >= exp_ptr[e], "tables_create: exp_ptr not monotone at %d", e);
        const int64_t p0 = exp_ptr[e], p1 = exp_ptr[e + 1];
        ACES_CHECK(ctx, p1 - p0 < (int64_t)1 << 30, "tables_create: experiment %d too large", e);
        const int E = (int)(p1 - p0);
        max_E = std::max(max_E, E);
        order.clear();
        sups.assign((size_t)E, std::vector<int32_t>());
        for (int p = 0; p < E; ++p) {
            const int64_t b0 = pauli_ptr[p0 + p], b1 = pauli_ptr[p0 + p + 1];
            ACES_CHECK(ctx, b1 >= b0, "tables_create: pauli_ptr not monotone at %lld", (long long)(p0 + p));
            sup.assign(pauli_bits + b0, pauli_bits + b1);
            for (int32_t b : sup)
                ACES_CHECK(ctx, b >= 0 && b < n_qubits, "tables_create: record bit %d out of range [0,%d)", b, n_qubits);
            std::sort(sup.begin(), sup.end());
            // a bit listed twice cancels: the per-shot sum is reduced mod 2 (src/estimate.jl:251)
            std::vector<int32_t>& red = sups[(size_t)p];
            for (size_t i = 0; i < sup.size();) {
                size_t j = i;
                while (j < sup.size() && sup[j] == sup[i]) ++j;
                if ((j - i) & 1) red.push_back(sup[i]);
                i = j;
            }
            order.push_back({(int32_t)red.size(), p});
        }
        // device order: by weight, so that the lanes of a warp run the same number of plane loads
        std::stable_sort(order.begin(), order.end(),
                         [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) {
                             return a.first < b.first;
                         });
        for (int r = 0; r < E; ++r) {
            const int p = order[(size_t)r].second;
            const std::vector<int32_t>& red = sups[(size_t)p];
            PauliDesc& d = descs[(size_t)(p0 + r)];
            d.w = (int32_t)red.size();
            d.ext = (int32_t)bits.size();
            d.orig = p;
            d.pad = 0;
            for (int i = 0; i < 4; ++i) d.q[i] = i < d.w ? red[(size_t)i] : 0;
            ACES_CHECK(ctx, bits.size() + red.size() < ((size_t)1 << 31), "tables_create: table too large");
            bits.insert(bits.end(), red.begin(), red.end());
        }
    }
    aces_tables* t = new aces_tables();
    t->n_qubits = n_qubits;
    t->B = (n_qubits + 7) / 8;
    t->n_experiments = n_experiments;
    t->max_E = max_E;
    t->exp_ptr.assign(exp_ptr, exp_ptr + n_experiments + 1);
    cudaError_t e1 = cudaMalloc(&t->d_pauli, sizeof(PauliDesc) * std::max<size_t>(descs.size(), 1));
    cudaError_t e2 = cudaMalloc(&t->d_bits, sizeof(int32_t) * std::max<size_t>(bits.size(), 1));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        if (t->d_pauli) cudaFree(t->d_pauli);
        if (t->d_bits) cudaFree(t->d_bits);
        delete t;
        return aces_fail(ctx, ACES_E_NOMEM, "tables_create: cudaMalloc failed");
    }
    if (!descs.empty())
        ACES_CUDA(ctx, cudaMemcpy(t->d_pauli, descs.data(), sizeof(PauliDesc) * descs.size(),
                                  cudaMemcpyHostToDevice));
    if (!bits.empty())
        ACES_CUDA(ctx, cudaMemcpy(t->d_bits, bits.data(), sizeof(int32_t) * bits.size(),
                                  cudaMemcpyHostToDevice));
    *out = t;
    return ACES_OK;
}

int32_t aces_tables_destroy(aces_ctx* ctx, aces_tables* t) {
    if (!ctx) return ACES_E_INVALID;
    if (!t) return ACES_OK;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t->d_pauli) cudaFree(t->d_pauli);
    if (t->d_bits) cudaFree(t->d_bits);
    delete t;
    return ACES_OK;
}

int32_t aces_parity_counts(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                           const int32_t* experiment_ids, const uint8_t* const* counts,
                           const int64_t* rows, const int64_t* ld, int32_t K,
                           const int64_t* prefix_shots, int64_t* odd_counts, int32_t accumulate) {
    return parity_run(ctx, tables, n_items, experiment_ids, counts, rows, ld, K, prefix_shots,
                      odd_counts, accumulate, false);
}

int32_t aces_parity_counts_async(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                                 const int32_t* experiment_ids, const uint8_t* const* counts,
                                 const int64_t* rows, const int64_t* ld, int32_t K,
                                 const int64_t* prefix_shots, int64_t* odd_counts,
                                 int32_t accumulate) {
    return parity_run(ctx, tables, n_items, experiment_ids, counts, rows, ld, K, prefix_shots,
                      odd_counts, accumulate, true);
}

int32_t aces_estimate_experiment(aces_ctx* ctx, const uint8_t* counts, int64_t rows, int64_t ld,
                                 int32_t n_cols, int32_t E, const int64_t* meas_ptr,
                                 const int32_t* meas_indices, const uint8_t* signs,
                                 int64_t shots_value, double* est) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, E >= 0 && n_cols >= 1, "estimate_experiment: E=%d n_cols=%d", E, n_cols);
    // The reference asserts that signs and indices have the same length (src/estimate.jl:292);
    // with C arrays that is the caller's contract.
    ACES_CHECK(ctx, shots_value >= 1 && shots_value <= rows,
               "estimate_experiment: shots_value=%lld not in [1, rows=%lld]", (long long)shots_value,
               (long long)rows);
    if (E == 0) return ACES_OK;
    ACES_CHECK(ctx, meas_ptr && signs && est, "estimate_experiment: NULL argument");
    const int64_t nb = meas_ptr[E];
    std::vector<int32_t> bits((size_t)nb);
    for (int64_t i = 0; i < nb; ++i) {
        ACES_CHECK(ctx, meas_indices[i] >= 1 && meas_indices[i] <= 8 * n_cols,
                   "estimate_experiment: qubit index %d outside 1..%d", meas_indices[i], 8 * n_cols);
        bits[(size_t)i] = meas_indices[i] - 1;
    }
    const int64_t exp_ptr[2] = {0, E};
    aces_tables* t = nullptr;
    ACES_TRY(aces_tables_create(ctx, 8 * n_cols, 1, exp_ptr, meas_ptr, bits.data(), &t));
    std::vector<int64_t> odd((size_t)E);
    const int32_t eid = 0;
    int rc = aces_parity_counts(ctx, t, 1, &eid, &counts, &rows, &ld, 1, &shots_value, odd.data(), 0);
    aces_tables_destroy(ctx, t);
    ACES_TRY(rc);
    for (int p = 0; p < E; ++p) {
        // ((-1)^sign * pauli_shots) / shots_value with pauli_shots = S - 2*odd (src/estimate.jl:297-298)
        const int64_t pauli_shots = shots_value - 2 * odd[(size_t)p];
        const int64_t sgn = signs[p] ? -1 : 1;
        est[p] = (double)(sgn * pauli_shots) / (double)shots_value;
    }
    return ACES_OK;
}

}  // extern "C"
