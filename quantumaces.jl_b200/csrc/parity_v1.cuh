// parity_v1.cuh -- K1 kernel, CTA-synchronous variant (general fallback: any number of Paulis per
// experiment, any number of byte columns that fits shared memory).  All warps of a CTA run
// phase A (bit-plane transposition) then phase B (per-Pauli XOR + popcount) on one tile with two
// CTA barriers per tile.
#pragma once
#include "parity_common.cuh"

namespace k1 {

// Phase A for one tile: every thread turns 8 words (32 shots) of one byte column into 8
// bit-plane words.  BOUNDARY tiles zero the bytes outside rows [lo_rel, hi_rel) first.
template <int TS, bool BOUNDARY>
__device__ __forceinline__ void phase_a(const uint8_t* __restrict__ st, uint32_t* __restrict__ planes,
                                        int B, int tid, int lo_rel, int hi_rel) {
    constexpr int UPC = TS / 32;
    constexpr int PSTRIDE = UPC + 4;
    const int nunits = B * UPC;
    for (int unit = tid; unit < nunits; unit += NT) {
        const int c = unit / UPC;
        const int u = unit - c * UPC;
        const uint8_t* col = st + c * TS;
        uint4 a = *reinterpret_cast<const uint4*>(col + 16 * u);
        uint4 b = *reinterpret_cast<const uint4*>(col + TS / 2 + 16 * u);
        if (BOUNDARY) {
            a = mask_rows(a, 16 * u, lo_rel, hi_rel);
            b = mask_rows(b, TS / 2 + 16 * u, lo_rel, hi_rel);
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

// Phase B for one Pauli over VPT consecutive 128-shot vectors: XOR of the planes of its support,
// popcount.  addr[i] = shared address of the first vector of support plane i (i < min(w, 4)).
template <int VPT, int PSTRIDE>
__device__ __forceinline__ uint32_t phase_b(const uint32_t (&addr)[4], int w,
                                            const int32_t* __restrict__ ext_bits, uint32_t plane0) {
    uint32_t cnt = 0;
    if (w == 1) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) cnt += popc4(lds128(addr[0] + i * 16));
    } else if (w == 2) {
#pragma unroll
        for (int i = 0; i < VPT; ++i)
            cnt += popc4(xor4(lds128(addr[0] + i * 16), lds128(addr[1] + i * 16)));
    } else {
#pragma unroll 2
        for (int i = 0; i < VPT; ++i) {
            uint4 x = xor4(xor4(lds128(addr[0] + i * 16), lds128(addr[1] + i * 16)),
                           lds128(addr[2] + i * 16));
            if (w > 3) {
                x = xor4(x, lds128(addr[3] + i * 16));
                for (int k = 4; k < w; ++k) {
                    const uint32_t q = (uint32_t)__ldg(&ext_bits[k]);
                    x = xor4(x, lds128(plane0 + q * (PSTRIDE * 4) + i * 16));
                }
            }
            cnt += popc4(x);
        }
    }
    return cnt;
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
    const int n_my = (int)(i1 - i0);

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
            const int64_t lo = sd.lo - r0, hi = sd.hi - r0;
            d->stream = s;
            d->lo_rel = (int32_t)(lo > 0 ? lo : 0);
            d->hi_rel = (int32_t)(hi < TS ? hi : TS);
            d->boundary = (lo > 0) || (hi < TS);
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
            if (s < n_my) issue(i0 + s, s);
    }

    // ---- per-thread Pauli state ---------------------------------------------------------------
    uint32_t addr[PPT][4];
    int32_t pw[PPT], pext[PPT], porig[PPT];
    uint32_t acc[PPT];
#pragma unroll
    for (int ch = 0; ch < PPT; ++ch) { pw[ch] = 0; acc[ch] = 0; pext[ch] = 0; porig[ch] = 0; }
    int cur_stream = -1;
    int64_t cur_out = 0;
    uint32_t v_first = 0;  // byte offset of this thread's first 128-shot vector within a plane row
    int vpt = NV;          // vectors per thread
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

    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < n_my; ++it) {
        mbar_wait(&bars[stage], parity);
        const int4 d = *reinterpret_cast<const int4*>(&sdesc[stage]);  // stream, lo_rel, hi_rel, boundary

        if (d.x != cur_stream || tiles_since_flush >= (0x7FFFFFFFu / TS)) {
            if (cur_stream >= 0) flush();
            if (d.x != cur_stream) {
                cur_stream = d.x;
                const StreamDesc* sd = &P.streams[cur_stream];
                cur_out = sd->out_off;
                const int E = sd->E;
                const int64_t poff = sd->pauli_off;
                int g = 0, G = 1;
                if (PPT == 1) {
                    // threads beyond the Paulis split the tile's vectors: G groups (a power of two)
                    const int EP = (E + 31) & ~31;
                    while (2 * G * EP <= NT && 2 * G <= NV) G *= 2;
                    g = tid / EP;
                    vpt = NV / G;
                    v_first = (uint32_t)(g * vpt) * 16u;
                }
#pragma unroll
                for (int ch = 0; ch < PPT; ++ch) {
                    int p;
                    bool on;
                    if (PPT == 1) {
                        const int EP = (E + 31) & ~31;
                        p = tid - g * EP;
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
                        addr[ch][0] = planes_s + qa.x * (PSTRIDE * 4) + v_first;
                        addr[ch][1] = planes_s + qa.y * (PSTRIDE * 4) + v_first;
                        addr[ch][2] = planes_s + qa.z * (PSTRIDE * 4) + v_first;
                        addr[ch][3] = planes_s + qa.w * (PSTRIDE * 4) + v_first;
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
            if (d.w)
                phase_a<TS, true>(st, planes, B, tid, d.y, d.z);
            else
                phase_a<TS, false>(st, planes, B, tid, 0, TS);
        }
        __syncthreads();  // planes complete; this raw stage is free again
        if (warp == 0 && it + NS < n_my) issue(i0 + it + NS, stage);

        // ---- phase B: per-Pauli XOR of planes + popcount --------------------------------------
#pragma unroll
        for (int ch = 0; ch < PPT; ++ch) {
            const int w = pw[ch];
            if (w > 0) {
                const uint32_t ext = (w > 4) ? (uint32_t)pext[ch] : 0u;
                switch (vpt) {
                    case 1: acc[ch] += phase_b<1, PSTRIDE>(addr[ch], w, P.bits + ext, planes_s + v_first); break;
                    case 2: acc[ch] += phase_b<2, PSTRIDE>(addr[ch], w, P.bits + ext, planes_s + v_first); break;
                    case 4: acc[ch] += phase_b<4, PSTRIDE>(addr[ch], w, P.bits + ext, planes_s + v_first); break;
                    case 8: acc[ch] += phase_b<8, PSTRIDE>(addr[ch], w, P.bits + ext, planes_s + v_first); break;
                    default:
                        for (int i = 0; i < vpt; i += 16) {
                            uint32_t a2[4] = {addr[ch][0] + i * 16u, addr[ch][1] + i * 16u,
                                              addr[ch][2] + i * 16u, addr[ch][3] + i * 16u};
                            acc[ch] += phase_b<16, PSTRIDE>(a2, w, P.bits + ext, planes_s + v_first + i * 16u);
                        }
                }
            }
        }
        __syncthreads();  // planes may be overwritten by the next item
        if (++stage == NS) { stage = 0; parity ^= 1u; }
    }
    if (cur_stream >= 0) flush();
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


}  // namespace k1
