// parity_v2.cuh -- K1 kernel, warp-autonomous variant (the default path).
//
// One CTA = NW consumer warps + 1 producer warp around a shared-memory ring of NS stages.
//   producer warp : per work item (one TS = NW*SUB shot tile of a stream) waits for the stage to
//                   be empty, writes the item descriptor, arms the stage's "full" mbarrier with
//                   the byte count and issues B bulk-async copies (TMA engine, SASS UBLKCP),
//                   one per byte column.
//   consumer warp w: owns rows [w*SUB, (w+1)*SUB) of every tile.  It waits on "full", runs
//                   phase A (8x32-bit butterfly transposition of its rows into private bit
//                   planes), arrives on "empty" (the raw stage is no longer needed), then runs
//                   phase B: lane l owns Paulis l, l+32, ... of the experiment; each is the XOR
//                   of a few plane rows (128-bit shared loads) + popcount into a register.
// There is no CTA-wide barrier in steady state: warps drift up to NS tiles apart, which absorbs
// the imbalance of either phase.  Consumers meet on a named barrier only when the stream
// (experiment, budget segment) changes, to flush the per-Pauli counts: registers -> shared
// (atomics across the NW warps) -> one 64-bit global atomic per Pauli.
#pragma once
#include "parity_common.cuh"

namespace k1 {

constexpr int V2_MAX_STAGES = 8;
constexpr int V2_OFF_EMPTY = 64;
constexpr int V2_OFF_DESC = 128;                                // 16 bytes per stage: what the consumers read
constexpr int V2_OFF_TABLES = V2_OFF_DESC + V2_MAX_STAGES * 16;  // 256

// Shared tables behind the header: the third / fourth plane row of weight-3/4 Paulis (s4) and, for
// the variants whose round count does not leave registers for them (RMAX > 8), the first two plane
// rows of every slot (dx).  The per-Pauli shared counters of a stream change (sacc) are NOT a table
// of their own: they alias the start of the plane area, which is dead while the consumers sit in
// change_stream (see there).
__host__ __device__ constexpr uint32_t v2_s4_off(int rmax) { (void)rmax; return (uint32_t)V2_OFF_TABLES; }
__host__ __device__ constexpr uint32_t v2_dx_off(int rmax) { return (uint32_t)(V2_OFF_TABLES + 4 * 32 * rmax); }
__host__ __device__ constexpr uint32_t v2_planes_off(int rmax) {
    return (uint32_t)((V2_OFF_TABLES + (rmax > 8 ? 2 : 1) * 4 * 32 * rmax + 127) & ~127);
}

// Byte pitch of a plane row: the SUB / 8 data bytes plus, for rows longer than one 16-byte vector,
// one vector of padding, so that the pitch is an odd number of vectors and the 128-bit loads of a
// quarter warp (eight different rows) fall into eight different bank groups.
__host__ __device__ constexpr int v2_row_pitch(int sub) { return sub / 8 + (sub / 8 > 16 ? 16 : 0); }

// Wait with an explicit back-off: one try_wait (which already suspends the warp for a short,
// implementation-defined time); while the phase has not flipped, sleep SLEEP_NS before polling again.
// A bare try_wait loop re-polls every few tens of nanoseconds: the producer warp (which waits for a
// whole tile's worth of compute) alone burned a tenth of the CTA's issue slots that way.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_backoff(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    while (!done) {
        __nanosleep(SLEEP_NS);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}


// Phase A for one warp: rows [row0, row0 + SUB) of every byte column of the stage -> the warp's
// private planes.  Plane row of record bit q = 8c + k is row (k * BMAX + c) (k-major: lanes of a
// warp then store to consecutive words; the compile-time row stride keeps the eight stores of a
// unit on immediate offsets), SUB/8 bytes per row.
// Lane l handles units l, l + 32, ...; a step of 32 units moves 32 / UPS columns forward and keeps
// the position inside the column, so both the source and the plane address advance by constants:
// col_s = stage column address of the lane's first unit, pl_s = its plane word.
template <int SUB, int BMAX, bool MASKED>
__device__ __forceinline__ void v2_phase_a(uint32_t col_s, uint32_t pl_s, int nunits, int CS, int lane, int lo,
                                           int hi) {
    constexpr int UPS = SUB / 32;  // 32-shot units per byte column
    constexpr uint32_t kstride = (uint32_t)(BMAX * v2_row_pitch(SUB));
    constexpr uint32_t pl_step = (uint32_t)((32 / UPS) * v2_row_pitch(SUB));
    const uint32_t col_step = (uint32_t)((32 / UPS) * CS);
    const int row0 = 16 * (lane & (UPS - 1));  // first row of the unit inside the warp's sub-tile
    for (int unit = lane; unit < nunits; unit += 32, col_s += col_step, pl_s += pl_step) {
        uint4 a = lds128(col_s);
        uint4 b = lds128(col_s + SUB / 2);
        if (MASKED) {
            a = mask_rows(a, row0, lo, hi);
            b = mask_rows(b, SUB / 2 + row0, lo, hi);
        }
        uint32_t w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w;
        uint32_t w4 = b.x, w5 = b.y, w6 = b.z, w7 = b.w;
        ACES_BFLY(w0, w1, 0x55555555u, 1) ACES_BFLY(w2, w3, 0x55555555u, 1)
        ACES_BFLY(w4, w5, 0x55555555u, 1) ACES_BFLY(w6, w7, 0x55555555u, 1)
        ACES_BFLY(w0, w2, 0x33333333u, 2) ACES_BFLY(w1, w3, 0x33333333u, 2)
        ACES_BFLY(w4, w6, 0x33333333u, 2) ACES_BFLY(w5, w7, 0x33333333u, 2)
        ACES_BFLY(w0, w4, 0x0F0F0F0Fu, 4) ACES_BFLY(w1, w5, 0x0F0F0F0Fu, 4)
        ACES_BFLY(w2, w6, 0x0F0F0F0Fu, 4) ACES_BFLY(w3, w7, 0x0F0F0F0Fu, 4)
        sts32(pl_s + 0 * kstride, w0); sts32(pl_s + 1 * kstride, w1);
        sts32(pl_s + 2 * kstride, w2); sts32(pl_s + 3 * kstride, w3);
        sts32(pl_s + 4 * kstride, w4); sts32(pl_s + 5 * kstride, w5);
        sts32(pl_s + 6 * kstride, w6); sts32(pl_s + 7 * kstride, w7);
    }
}

// BMAX = byte columns the plane layout is dimensioned for (B <= BMAX).
#ifndef ACES_K1_PSLEEP
#define ACES_K1_PSLEEP 128
#endif
#ifndef ACES_K1_CSLEEP
#define ACES_K1_CSLEEP 32
#endif
constexpr int PRODUCER_SLEEP_NS = ACES_K1_PSLEEP;  // back-off of the producer's wait for a free stage
constexpr int CONSUMER_SLEEP_NS = ACES_K1_CSLEEP;  // back-off of a consumer's wait for a full stage

template <int NW, int SUB, int BMAX, int RMAX>
__global__ void __launch_bounds__((NW + 1) * 32, RMAX <= 16 ? 3 : 2) parity_kernel_v2(const ParityParams P) {
    constexpr int TS = NW * SUB;
    constexpr int CS = TS + 64;     // raw column stride: +64 B keeps quarter-warp LDS.128 conflict-free
    constexpr int V = SUB / 128;    // 128-shot vectors per plane row
    constexpr int ROWB = SUB / 8;   // data bytes per plane row
    constexpr int RSTR = v2_row_pitch(SUB);  // byte pitch of a plane row
    constexpr int NCT = NW * 32;    // consumer threads
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem + V2_OFF_EMPTY);
    int4* sdesc = reinterpret_cast<int4*>(smem + V2_OFF_DESC);
    const int B = P.B;
    const int NS = P.NS;
    constexpr uint32_t planes_bytes = (uint32_t)((8 * BMAX + 1) * RSTR);  // + one all-zero row
    const uint32_t planes_off = v2_planes_off(RMAX);
    uint32_t* sacc = reinterpret_cast<uint32_t*>(smem + planes_off);  // aliases the planes: see change_stream
    const uint32_t raw_off = (planes_off + NW * planes_bytes + 127u) & ~127u;
    const uint32_t stage_bytes = (uint32_t)(B * CS);
    const uint32_t smem_s = smem_u32(smem);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform

    // contiguous work items per CTA: equal estimated cost when the host supplied a partition (tiles of
    // experiments with more Paulis take longer), equal counts otherwise
    const int64_t i0 = P.cta_start ? P.cta_start[blockIdx.x] : (P.total_items * (int64_t)blockIdx.x) / gridDim.x;
    const int64_t i1 = P.cta_start ? P.cta_start[blockIdx.x + 1] : (P.total_items * (int64_t)(blockIdx.x + 1)) / gridDim.x;
    if (i0 >= i1) return;
    const int n_my = (int)(i1 - i0);

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        // ================================ producer warp ========================================
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        int prod_stream = 0;
        int last_stream = -1;
        uint32_t since_flush = 0;
        int stage = 0;
        uint32_t parity = 1;  // a fresh mbarrier passes a wait on parity 1
        for (int it = 0; it < n_my; ++it) {
            // Everything an item needs is worked out BEFORE the wait for its stage (the producer
            // blocks there whenever the ring is full): the table reads from global memory are then
            // off the refill path, and the copies go out as soon as the last consumer lets go.
            int4 head = make_int4(0, 0, 0, 0);
            const uint8_t* src = nullptr;
            int64_t ld = 0;
            uint32_t bytes = 0;
            int ncols = 0, cols_off = 0;
            if (lane == 0) {
                const int64_t item = i0 + it;
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
                bytes = (uint32_t)(avail < TS ? avail : TS);
                const int64_t lo = sd.lo - r0, hi = sd.hi - r0;
                // bit 0: the tile is not entirely inside the segment; bit 1: the consumers must
                // flush their counters first (new stream, or 2^19 tiles since the last flush so
                // that the 32-bit shared counters cannot overflow)
                const bool flush = (s != last_stream) || (since_flush >= (1u << 19));
                since_flush = flush ? 1u : since_flush + 1u;
                last_stream = s;
                head = make_int4(s, (int32_t)(lo > 0 ? lo : 0), (int32_t)(hi < TS ? hi : TS),
                                 (((lo > 0) || (hi < TS)) ? 1 : 0) | (flush ? 2 : 0));
                src = sd.base + r0;
                ld = sd.ld;
                ncols = sd.ncols;
                cols_off = sd.cols_off;
            }
            src = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)src, 0);
            ld = __shfl_sync(0xffffffffu, ld, 0);
            bytes = __shfl_sync(0xffffffffu, bytes, 0);
            ncols = __shfl_sync(0xffffffffu, ncols, 0);
            cols_off = __shfl_sync(0xffffffffu, cols_off, 0);
            // global column numbers of this lane's copies (local column c holds global column
            // cols[c] of the sample matrix); a group has at most 41 columns
            const int32_t* cols = P.cols + cols_off;
            const int64_t off0 = lane < ncols ? (int64_t)__ldg(cols + lane) * ld : 0;
            const int64_t off1 = lane + 32 < ncols ? (int64_t)__ldg(cols + lane + 32) * ld : 0;
            mbar_wait_backoff<PRODUCER_SLEEP_NS>(smem_u32(&empty[stage]), parity);
            if (lane == 0) {
                sdesc[stage] = head;
                mbar_arrive_expect_tx(&full[stage], bytes * (uint32_t)ncols);
            }
            __syncwarp();
            uint8_t* dst = smem + raw_off + (size_t)stage * stage_bytes;
            if (lane < ncols) bulk_g2s(dst + lane * CS, src + off0, bytes, &full[stage], policy);
            if (lane + 32 < ncols) bulk_g2s(dst + (lane + 32) * CS, src + off1, bytes, &full[stage], policy);
            for (int c = lane + 64; c < ncols; c += 32)
                bulk_g2s(dst + c * CS, src + (int64_t)__ldg(cols + c) * ld, bytes, &full[stage], policy);
            if (++stage == NS) { stage = 0; parity ^= 1u; }
        }
        return;
    }

    // ================================== consumer warps =========================================
    uint32_t planes_s = smem_s + planes_off + (uint32_t)warp * planes_bytes;
    uint32_t empty_s = smem_u32(&empty[0]), full_s = smem_u32(&full[0]);
    // per-lane addresses of phase A: the lane's first unit in a raw stage and its plane word
    constexpr int UPS_ = SUB / 32;
    uint32_t raw_s = smem_s + raw_off + (uint32_t)(warp * SUB) + (uint32_t)((lane / UPS_) * CS + 16 * (lane % UPS_));
    uint32_t pl_lane = planes_s + (uint32_t)((lane / UPS_) * RSTR + 4 * (lane % UPS_)), lane_r = (uint32_t)lane;
    int nunits = 0;  // 32-shot units per sub-tile of the current stream: its columns x UPS
    // keep the loop invariants in registers (the compiler would otherwise rematerialise them from
    // the special registers every tile)
    asm volatile("" : "+r"(planes_s), "+r"(empty_s), "+r"(full_s), "+r"(raw_s), "+r"(lane_r), "+r"(pl_lane));
    const int lane_c = (int)lane_r;
    uint32_t acc[RMAX];
    // per register slot: weight-1 slots hold the shared-memory address of their plane row;
    // weight-2 slots hold the two plane-row offsets packed as a0 | a1 << 16
    constexpr bool DXS = RMAX > 8;   // slot descriptors in shared memory instead of registers
    uint32_t dx[DXS ? 1 : RMAX];
    constexpr uint32_t zero_row = (uint32_t)(8 * BMAX * RSTR);
    for (int i = lane; i < ROWB / 4; i += 32) sts32(planes_s + zero_row + 4 * i, 0u);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        acc[r] = 0;
        if (!DXS) dx[r] = 0;
    }
    // The Paulis of a slice are sorted by weight; lane l owns Paulis l, l+32, ... ("rounds").
    // Rounds [0,n1) hold only weight-1 Paulis, [n1,n12) only weight <= 2 (a missing second bit,
    // an identity Pauli or a lane past the end of the table points at the all-zero row),
    // [n12,R) are table-driven.  The classes live in fixed register slots so that each class runs
    // as straight-line code entered through one computed jump (no per-round branches):
    //   weight-2 rounds  -> slots [0, n2),  n2 = n12 - n1          round = n1 + slot
    //   weight<=4 rounds -> slots [n2, n2 + n4)                     round = n1 + slot   (third and
    //                       fourth plane row in the shared table s4; predicated per slot)
    //   table rounds     -> slots [n2 + n4, n2 + n4 + ng)           round = n1 + slot
    //   weight-1 rounds  -> slots [RMAX - n1, RMAX)                 round = slot - (RMAX - n1)
    int n1 = 0, n2 = 0, n4 = 0, ng = 0;
    const uint32_t s4_s = smem_s + v2_s4_off(RMAX) + 4u * (uint32_t)lane;
    const uint32_t dx_s = smem_s + v2_dx_off(RMAX) + 4u * (uint32_t)lane;
    int R = 0;                        // rounds in use
    int E = 0;
    int cur_stream = -1;
    int64_t cur_out = 0, cur_poff = 0;

    auto row_off = [&](uint32_t q) -> uint32_t {  // plane-row byte offset of record bit q
        return ((q & 7u) * (uint32_t)BMAX + (q >> 3)) * (uint32_t)RSTR;
    };
    auto round_of_slot = [&](int slot) -> int {  // -1: slot not in use
        if (slot < n2 + n4 + ng) return n1 + slot;
        if (slot >= RMAX - n1) return slot - (RMAX - n1);
        return -1;
    };

    // registers -> shared -> global; then (optionally) load the Pauli slice of stream `ns`
    auto change_stream = [&](int ns) {
        // The shared counters alias the start of the plane area.  Every consumer warp is between two
        // tiles here, so once all of them have arrived the planes are dead: clear the counters,
        // add the registers, drain to global memory (which leaves every counter zero again, so a
        // zero row the counters may overlap is zero when the planes come back into use).
        if (cur_stream >= 0) {
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            for (int p = tid; p < 32 * R; p += NCT) sacc[p] = 0;
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
#pragma unroll
            for (int sl = 0; sl < RMAX; ++sl) {
                const int r = round_of_slot(sl);
                const int p = r * 32 + lane;
                if (r >= 0 && p < E && acc[sl] != 0) atomicAdd(&sacc[p], acc[sl]);
                acc[sl] = 0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            for (int p = tid; p < E; p += NCT) {
                const uint32_t v = sacc[p];
                if (v != 0) {
                    const int orig = P.pauli[cur_poff + p].orig;
                    atomicAdd(&P.seg[cur_out + orig], (unsigned long long)v);
                    sacc[p] = 0;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
        if (ns < 0 || ns == cur_stream) return;
        const StreamDesc* sd = &P.streams[ns];
        cur_out = sd->out_off;
        // the next budget segment of the same matrix and column group keeps the Pauli slots: only
        // the output block changes
        const bool same_paulis = cur_stream >= 0 && sd->pauli_off == cur_poff && sd->E == E &&
                                 sd->ncols * UPS_ == nunits;
        cur_stream = ns;
        if (same_paulis) return;
        cur_poff = sd->pauli_off;
        E = sd->E;
        nunits = sd->ncols * UPS_;
        R = (E + 31) >> 5;
        // pass 1: classify the rounds (the table is sorted by weight: classes are prefixes)
        int c1 = 0, c12 = 0, c14 = 0;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int p = r * 32 + lane;
            int w = 0;
            if (r < R && p < E) w = (int)__ldg(&P.pauli[cur_poff + p].w);
            const bool all1 = __all_sync(0xffffffffu, w == 1);
            const bool all2 = __all_sync(0xffffffffu, w <= 2);
            const bool all4 = __all_sync(0xffffffffu, w <= 4);
            if (all1 && c1 == r) c1 = r + 1;
            if (all2 && c12 == r && r < R) c12 = r + 1;
            if (all4 && c14 == r && r < R) c14 = r + 1;
        }
        n1 = c1;
        n2 = c12 - c1;
        n4 = c14 - c12;
        ng = R - c14;
        // pass 2: plane rows of every slot
#pragma unroll
        for (int sl = 0; sl < RMAX; ++sl) {
            const int r = round_of_slot(sl);
            const int p = r * 32 + lane;
            uint32_t a0 = zero_row, a1 = zero_row, a2 = zero_row, a3 = zero_row;
            if (r >= 0 && p < E) {
                const PauliDesc* pd = &P.pauli[cur_poff + p];
                const int w = (int)__ldg(&pd->w);
                if (w >= 1) a0 = row_off((uint32_t)__ldg(&pd->q[0]));
                if (w >= 2) a1 = row_off((uint32_t)__ldg(&pd->q[1]));
                if (w >= 3) a2 = row_off((uint32_t)__ldg(&pd->q[2]));
                if (w >= 4) a3 = row_off((uint32_t)__ldg(&pd->q[3]));
            }
            if (DXS)  // every consumer warp writes the same table; weight-1 slots hold the row offset
                sts32(dx_s + 128u * (uint32_t)sl, (sl >= RMAX - n1) ? a0 : (a0 | (a1 << 16)));
            else
                dx[sl] = (sl >= RMAX - n1) ? planes_s + a0 : (a0 | (a1 << 16));
            // every consumer warp writes the same table (same lane <-> Pauli assignment)
            if (sl >= n2 && sl < n2 + n4) sts32(s4_s + 128u * (uint32_t)sl, a2 | (a3 << 16));
        }
    };

    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < n_my; ++it) {
        mbar_wait_backoff<CONSUMER_SLEEP_NS>(full_s + 8u * (uint32_t)stage, parity);
        const int4 d = sdesc[stage];  // stream, lo_rel, hi_rel, flags
        if (d.w & 2) change_stream(d.x);
        const uint32_t stage_s = raw_s + (uint32_t)stage * stage_bytes;
        bool any = true;
        if (d.w & 1) {
            // rows of this warp relative to its sub-tile start
            const int lo = d.y - warp * SUB, hi = d.z - warp * SUB;
            any = (hi > 0) && (lo < SUB);
            if (any) {
                if (lo > 0 || hi < SUB)
                    v2_phase_a<SUB, BMAX, true>(stage_s, pl_lane, nunits, CS, lane_c, lo, hi);
                else
                    v2_phase_a<SUB, BMAX, false>(stage_s, pl_lane, nunits, CS, lane_c, 0, SUB);
            }
        } else {
            v2_phase_a<SUB, BMAX, false>(stage_s, pl_lane, nunits, CS, lane_c, 0, SUB);
        }
        __syncwarp();
        if (lane_c == 0)  // the raw stage may be refilled
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_s + 8u * (uint32_t)stage) : "memory");

        if (any) {
#define ACES_DX(SL) (DXS ? lds32(dx_s + 128u * (uint32_t)(SL)) : dx[DXS ? 0 : (SL)])
#define ACES_W2(SL)                                                                          \
    {                                                                                        \
        const uint32_t dxv = ACES_DX(SL);                                                    \
        const uint32_t b0 = planes_s + (dxv & 0xFFFFu), b1 = planes_s + (dxv >> 16);         \
        uint32_t cnt = 0;                                                                    \
        _Pragma("unroll") for (int v = 0; v < V; ++v)                                        \
            cnt += popc4(xor4(lds128(b0 + 16 * v), lds128(b1 + 16 * v)));                    \
        acc[SL] += cnt;                                                                      \
    }
#define ACES_W1(SL)                                                                          \
    {                                                                                        \
        const uint32_t b0 = DXS ? planes_s + ACES_DX(SL) : ACES_DX(SL);                      \
        uint32_t cnt = 0;                                                                    \
        _Pragma("unroll") for (int v = 0; v < V; ++v) cnt += popc4(lds128(b0 + 16 * v));     \
        acc[SL] += cnt;                                                                      \
    }
#define ACES_C2(K) case K: if constexpr (RMAX >= K) ACES_W2(K - 1) [[fallthrough]];
#define ACES_C1(K) case K: if constexpr (RMAX >= K) ACES_W1(RMAX - K) [[fallthrough]];
            switch (n2) {
                ACES_C2(32) ACES_C2(31) ACES_C2(30) ACES_C2(29) ACES_C2(28) ACES_C2(27) ACES_C2(26) ACES_C2(25)
                ACES_C2(24) ACES_C2(23) ACES_C2(22) ACES_C2(21) ACES_C2(20) ACES_C2(19) ACES_C2(18) ACES_C2(17)
                ACES_C2(16) ACES_C2(15) ACES_C2(14) ACES_C2(13) ACES_C2(12) ACES_C2(11) ACES_C2(10) ACES_C2(9)
                ACES_C2(8) ACES_C2(7) ACES_C2(6) ACES_C2(5) ACES_C2(4) ACES_C2(3) ACES_C2(2) ACES_C2(1)
                default: break;
            }
            switch (n1) {
                ACES_C1(32) ACES_C1(31) ACES_C1(30) ACES_C1(29) ACES_C1(28) ACES_C1(27) ACES_C1(26) ACES_C1(25)
                ACES_C1(24) ACES_C1(23) ACES_C1(22) ACES_C1(21) ACES_C1(20) ACES_C1(19) ACES_C1(18) ACES_C1(17)
                ACES_C1(16) ACES_C1(15) ACES_C1(14) ACES_C1(13) ACES_C1(12) ACES_C1(11) ACES_C1(10) ACES_C1(9)
                ACES_C1(8) ACES_C1(7) ACES_C1(6) ACES_C1(5) ACES_C1(4) ACES_C1(3) ACES_C1(2) ACES_C1(1)
                default: break;
            }
#undef ACES_C1
#undef ACES_C2
#undef ACES_W1
#undef ACES_W2
            if (n4 > 0) {
                // supports of three or four bits (products of two measured Paulis: the covariance
                // Paulis of a full-covariance design)
#pragma unroll
                for (int sl = 0; sl < RMAX; ++sl) {
                    if (sl >= n2 && sl < n2 + n4) {
                        const uint32_t hi2 = lds32(s4_s + 128u * (uint32_t)sl);
                        const uint32_t dxv = ACES_DX(sl);
                        const uint32_t b0 = planes_s + (dxv & 0xFFFFu), b1 = planes_s + (dxv >> 16);
                        const uint32_t b2 = planes_s + (hi2 & 0xFFFFu), b3 = planes_s + (hi2 >> 16);
                        uint32_t cnt = 0;
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            const uint4 x0 = lds128(b0 + 16 * v), x1 = lds128(b1 + 16 * v);
                            const uint4 x2 = lds128(b2 + 16 * v), x3 = lds128(b3 + 16 * v);
                            cnt += (__popc(xor3(x0.x, x1.x, x2.x) ^ x3.x) + __popc(xor3(x0.y, x1.y, x2.y) ^ x3.y)) +
                                   (__popc(xor3(x0.z, x1.z, x2.z) ^ x3.z) + __popc(xor3(x0.w, x1.w, x2.w) ^ x3.w));
                        }
                        acc[sl] += cnt;
                    }
                }
            }
            if (ng > 0) {
                // wider supports: table-driven
#pragma unroll
                for (int sl = 0; sl < RMAX; ++sl) {
                    const int r = n1 + sl;
                    const int p = r * 32 + lane;
                    if (sl >= n2 + n4 && sl < n2 + n4 + ng && p < E) {
                        const PauliDesc* pd = &P.pauli[cur_poff + p];
                        const int w = pd->w;
                        const int ext = pd->ext;
                        uint32_t cnt = 0;
                        for (int v = 0; v < V && w > 0; ++v) {
                            uint4 x = make_uint4(0, 0, 0, 0);
                            for (int k = 0; k < w; ++k) {
                                const uint32_t q = (uint32_t)__ldg(&P.bits[ext + k]);
                                x = xor4(x, lds128(planes_s + row_off(q) + 16 * v));
                            }
                            cnt += popc4(x);
                        }
                        acc[sl] += cnt;
                    }
                }
            }
        }
        __syncwarp();  // every lane is done with the planes before the next phase A overwrites them
        if (++stage == NS) { stage = 0; parity ^= 1u; }
    }
    change_stream(-1);
#undef ACES_DX
}

struct V2Cfg {
    int NW, SUB, BMAX, RMAX, NS;
    size_t smem;
};

// Chooses the v2 tiling for B byte columns and up to max_E Paulis per slice.  Returns false when
// the shape is outside what v2 covers (the caller falls back to the v1 kernel).
inline bool v2_pick(int B, int max_E, int smem_optin, V2Cfg* cfg) {
    int SUB, BMAX;
    // sub-tile rows per warp: as large as three resident CTAs per SM (planes + two or more raw
    // stages each) allow -- more rows amortise the per-tile overhead, fewer CTAs lose more
    if (B <= 3) { SUB = 512; BMAX = 7; }
    else if (B <= 7) { SUB = 256; BMAX = 7; }
    else if (B <= 13) { SUB = 128; BMAX = 13; }
    else if (B <= 21) { SUB = 128; BMAX = 21; }  // d = 9 (B = 21)
    else if (B <= 22) { SUB = 128; BMAX = 22; }  // the widest record (or column group) that keeps 3 CTAs/SM with two stages
    else if (B <= 25) { SUB = 128; BMAX = 25; }
    else if (B <= 41) { SUB = 128; BMAX = 41; }
    else return false;
    const int NW = 8;
    if ((8 * BMAX + 1) * v2_row_pitch(SUB) >= 65536) return false;  // plane-row offsets are packed in 16 bits
    const int RMAX = max_E <= 256 ? 8 : (max_E <= 512 ? 16 : 32);
    const size_t planes = (size_t)NW * (8 * BMAX + 1) * v2_row_pitch(SUB);
    const size_t raw_off = (v2_planes_off(RMAX) + planes + 127) & ~(size_t)127;
    const size_t stage = (size_t)B * (NW * SUB + 64);
    const char* env_ns = getenv("ACES_K1_NS");
    const char* env_cta = getenv("ACES_K1_CTAS");
    // Resident CTAs per SM: what the register budget of the variant allows (launch bounds of the
    // kernel), fewer when the shared memory of that many CTAs does not hold two stages each.
    int ns = 0;
    for (int ctas = env_cta ? atoi(env_cta) : (RMAX <= 16 ? 3 : 2); ctas >= 1 && ns == 0; --ctas) {
        const size_t budget = ((size_t)smem_optin + 1024) / (size_t)ctas - 1024;
        for (int s = (env_ns ? atoi(env_ns) : 4); s >= 2; --s)
            if (raw_off + s * stage <= budget) { ns = s; break; }
    }
    if (ns == 0) return false;
    cfg->NW = NW;
    cfg->SUB = SUB;
    cfg->BMAX = BMAX;
    cfg->RMAX = RMAX;
    cfg->NS = ns;
    cfg->smem = raw_off + ns * stage;
    return true;
}

// Resident CTAs of the variant on the whole GPU (the persistent grid), or an error code (< 0).
template <int NW, int SUB, int BMAX, int RMAX>
int v2_grid_t(aces_ctx* ctx, const V2Cfg& cfg) {
    auto kern = parity_kernel_v2<NW, SUB, BMAX, RMAX>;
    ACES_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    int per_sm = 0;
    ACES_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (NW + 1) * 32, cfg.smem));
    if (per_sm < 1)
        return aces_fail(ctx, ACES_E_UNSUPPORTED, "parity kernel v2 does not fit on an SM (smem %zu)", cfg.smem);
    return ctx->sm_count * per_sm;
}

template <int NW, int SUB, int BMAX, int RMAX>
int v2_launch_t(aces_ctx* ctx, const ParityParams& P, const V2Cfg& cfg) {
    auto kern = parity_kernel_v2<NW, SUB, BMAX, RMAX>;
    int64_t grid = P.grid;
    if (grid < 1) {
        const int g = v2_grid_t<NW, SUB, BMAX, RMAX>(ctx, cfg);
        if (g < 0) return g;
        grid = g;
        if (grid > P.total_items) grid = P.total_items;
    }
    if (ctx->timing) ACES_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    kern<<<(unsigned)grid, (NW + 1) * 32, cfg.smem, ctx->stream>>>(P);
    ACES_CUDA(ctx, cudaGetLastError());
    if (ctx->timing) {
        ACES_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
        ctx->timed = 1;
    }
    ctx->launches += 1;
    return ACES_OK;
}

inline int v2_grid(aces_ctx* ctx, const V2Cfg& cfg) {
#define ACES_V2_G(SUBV, BMV)                                                    \
    (cfg.RMAX == 8 ? v2_grid_t<8, SUBV, BMV, 8>(ctx, cfg)                       \
                   : cfg.RMAX == 16 ? v2_grid_t<8, SUBV, BMV, 16>(ctx, cfg)     \
                                    : v2_grid_t<8, SUBV, BMV, 32>(ctx, cfg))
    if (cfg.SUB == 128 && cfg.BMAX == 21) return ACES_V2_G(128, 21);
    if (cfg.SUB == 128 && cfg.BMAX == 22) return ACES_V2_G(128, 22);
    if (cfg.SUB == 128 && cfg.BMAX == 25) return ACES_V2_G(128, 25);
    if (cfg.SUB == 128 && cfg.BMAX == 41) return ACES_V2_G(128, 41);
    if (cfg.SUB == 128) return ACES_V2_G(128, 13);
    if (cfg.SUB == 256) return ACES_V2_G(256, 7);
    return ACES_V2_G(512, 7);
#undef ACES_V2_G
}

inline int v2_launch(aces_ctx* ctx, const ParityParams& P, const V2Cfg& cfg) {
#define ACES_V2_R(SUBV, BMV)                                                             \
    (cfg.RMAX == 8 ? v2_launch_t<8, SUBV, BMV, 8>(ctx, P, cfg)                           \
                   : cfg.RMAX == 16 ? v2_launch_t<8, SUBV, BMV, 16>(ctx, P, cfg)         \
                                    : v2_launch_t<8, SUBV, BMV, 32>(ctx, P, cfg))
    if (cfg.SUB == 128 && cfg.BMAX == 21) return ACES_V2_R(128, 21);
    if (cfg.SUB == 128 && cfg.BMAX == 22) return ACES_V2_R(128, 22);
    if (cfg.SUB == 128 && cfg.BMAX == 25) return ACES_V2_R(128, 25);
    if (cfg.SUB == 128 && cfg.BMAX == 41) return ACES_V2_R(128, 41);
    if (cfg.SUB == 128) return ACES_V2_R(128, 13);
    if (cfg.SUB == 256) return ACES_V2_R(256, 7);
    return ACES_V2_R(512, 7);
#undef ACES_V2_R
}

}  // namespace k1
