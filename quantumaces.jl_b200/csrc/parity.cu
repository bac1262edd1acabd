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
#include <array>
#include <cstdio>
#include <climits>
#include <cstring>

#include "aces_common.h"

#include "parity_common.cuh"
#include "parity_v1.cuh"
#include "parity_v2.cuh"
#include <cstdlib>
#include <map>

using namespace k1;


// A column group: the Paulis of one experiment whose supports lie in a set of at most
// GROUP_COLS byte columns, at most GROUP_PAULIS of them.  One group = one kernel stream that
// reads only its columns (the producer issues one bulk copy per column anyway) and evaluates its
// Paulis on LOCAL record bits 8 * (local column) + bit.  Narrow experiments are one group over
// all columns; wide records (d = 17: 73 byte columns) and large experiments split into several,
// so that every stream runs the best-tuned kernel variant (<= 25 columns, 8 rounds of 32 Paulis).
constexpr int GROUP_COLS = 22;
constexpr int GROUP_PAULIS = 256;

struct ColGroup {
    int64_t pauli_off;  // first PauliDesc (device order: by weight)
    int32_t E;
    int32_t cols_off, ncols;
    int32_t r1 = 0, r2 = 0, r4 = 0, rg = 0;  // rounds of 32 Paulis by class (weight 1, <= 2, <= 4, wider), as the kernel classifies them
};

struct aces_tables {
    int32_t n_qubits = 0;
    int32_t B = 0;
    int32_t n_experiments = 0;
    int32_t max_E = 0;
    std::vector<int64_t> exp_ptr;    // host copy: Paulis of experiment e (caller order; sizes the outputs)
    std::vector<int32_t> grp_ptr;    // groups of experiment e: groups[grp_ptr[e] .. grp_ptr[e+1])
    std::vector<ColGroup> groups;    // identity Paulis (odd count always 0) belong to no group
    int32_t max_group_E = 0, max_group_cols = 0;
    bool grouped = false;            // false: one group per experiment over all columns, global bits
    PauliDesc* d_pauli = nullptr;
    int32_t* d_bits = nullptr;
    int32_t* d_cols = nullptr;
};

// ---- ingest of row-major records ----------------------------------------
// Stim's sample(bit_packed = true) returns a ROW-major uint8 [shots x B] array; the reference turns it
// into a column-major Julia matrix with a strided element-by-element copy on the host
// (pyconvert(Matrix{UInt8}, ...), src/simulate.jl:60-63).  Here the row-major buffer is taken as it
// is (one contiguous host-to-device copy) and transposed on the device into the pitched column-major
// arena the parity kernel reads: an HBM-bound byte transposition, one CTA per tile of TR shots.
struct TransDesc {
    const uint8_t* src;  // row-major, `pitch` bytes between shots (device)
    int64_t pitch, rows;
    uint8_t* dst;        // column-major, `dst_ld` bytes between byte columns (a multiple of 128)
    int64_t dst_ld;
};
constexpr int TR = 512;  // shots per tile

__global__ void __launch_bounds__(256) rowmajor_to_colmajor_kernel(const TransDesc* __restrict__ descs, int B) {
    extern __shared__ __align__(16) uint8_t tile[];  // [TR][B] as in the source
    const TransDesc d = descs[blockIdx.y];
    const int tid = threadIdx.x;
    for (int64_t r0 = (int64_t)blockIdx.x * TR; r0 < d.rows; r0 += (int64_t)gridDim.x * TR) {
        const int nr = (int)min((int64_t)TR, d.rows - r0);
        const uint8_t* src = d.src + r0 * d.pitch;
        if (d.pitch == B && (((uintptr_t)src) & 15) == 0) {
            const int nbytes = nr * B, nvec = nbytes >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* t4 = reinterpret_cast<uint4*>(tile);
            for (int i = tid; i < nvec; i += 256) t4[i] = __ldcs(s4 + i);
            for (int i = (nvec << 4) + tid; i < nbytes; i += 256) tile[i] = src[i];
        } else {
            for (int i = tid; i < nr * B; i += 256) {
                const int sidx = i / B, c = i - sidx * B;
                tile[i] = src[(int64_t)sidx * d.pitch + c];
            }
        }
        __syncthreads();
        // thread (c, q): four consecutive shots of byte column c -> one 32-bit store (the destination
        // pitch is a multiple of 128 and r0 of 512: aligned).  Shots past `rows` are written as zero.
        const int quads = TR / 4;
        for (int i = tid; i < B * quads; i += 256) {
            const int c = i / quads, q = i - c * quads;
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int sidx = 4 * q + k;
                if (sidx < nr) w |= (uint32_t)tile[sidx * B + c] << (8 * k);
            }
            if (4 * q < ((nr + 15) & ~15))  // up to the 16-shot copy granularity of the parity kernel
                *reinterpret_cast<uint32_t*>(d.dst + (int64_t)c * d.dst_ld + r0 + 4 * q) = w;
        }
        __syncthreads();
    }
}

// layout: 0 = column-major sample matrices (ld = bytes between byte columns); 1 = row-major records
// (ld = bytes between shots, >= B)
static int parity_run(aces_ctx* ctx, const aces_tables* t, int32_t n_items,
                      const int32_t* experiment_ids, const uint8_t* const* counts,
                      const int64_t* rows, const int64_t* ld, int32_t K,
                      const int64_t* prefix_shots, int64_t* odd_counts, int32_t accumulate,
                      bool async_only, int layout = 0) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, t != nullptr, "parity_counts: tables is NULL");
    ACES_CHECK(ctx, n_items >= 0 && K >= 1, "parity_counts: n_items=%d K=%d", n_items, K);
    if (n_items == 0) return ACES_OK;
    ACES_CHECK(ctx, experiment_ids && counts && rows && ld && prefix_shots && odd_counts,
               "parity_counts: NULL argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = t->B;
    // largest experiment of the tables decides how many Paulis a lane / thread must hold
    TileCfg cfg;
    V2Cfg cfg2;
    const char* env_impl = getenv("ACES_K1_IMPL");
    // the v2 kernel sizes its stages and planes for the widest column group of the tables
    const bool use_v2 = !(env_impl && atoi(env_impl) == 1) && t->max_group_E <= 1024 &&
                        v2_pick(std::max(t->max_group_cols, 1), std::max(t->max_group_E, 1), ctx->smem_optin, &cfg2);
    if (!use_v2 && !pick_tile(B, ctx->smem_optin, &cfg))
        return aces_fail(ctx, ACES_E_UNSUPPORTED,
                         "parity_counts: %d byte columns per shot exceed the shared-memory tiling", B);
    const int TS = use_v2 ? cfg2.NW * cfg2.SUB : cfg.TS;

    // ---- validate, lay out outputs, decide which matrices need a pitched device copy ----------
    std::vector<int64_t> item_off(n_items + 1, 0);
    std::vector<int32_t> item_E(n_items);
    std::vector<char> in_place(n_items);
    std::vector<size_t> arena_off(n_items, 0), raw_off(n_items, 0);
    std::vector<char> src_dev(n_items, 0), packed(n_items, 0);
    size_t arena_bytes = 0, raw_bytes = 0, pack_bytes = 0;
    constexpr size_t PACK_LIMIT = 256 * 1024;
    int max_E = 0;
    for (int i = 0; i < n_items; ++i) {
        const int e = experiment_ids[i];
        ACES_CHECK(ctx, e >= 0 && e < t->n_experiments, "parity_counts: item %d: experiment %d out of range", i, e);
        ACES_CHECK(ctx, rows[i] >= 0, "parity_counts: item %d: rows=%lld", i, (long long)rows[i]);
        if (layout == 0)
            ACES_CHECK(ctx, ld[i] >= rows[i], "parity_counts: item %d: ld=%lld < rows=%lld", i,
                       (long long)ld[i], (long long)rows[i]);
        else
            ACES_CHECK(ctx, ld[i] >= B, "parity_counts: item %d: row pitch %lld < %d bytes per shot", i,
                       (long long)ld[i], B);
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
        // The asynchronous entry point takes device matrices by contract: the per-item pointer query
        // (about a microsecond each, times ~100 experiments per call) is skipped there.
        const bool dev = rows[i] > 0 && (async_only || aces_is_device_ptr(counts[i]));
        in_place[i] = layout == 0 && dev && (((uintptr_t)counts[i]) % 16 == 0) && (ld[i] % 16 == 0);
        if (!in_place[i] && rows[i] > 0) {
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            if (layout == 0 && !dev && pitch * B <= PACK_LIMIT) {
                // a small host matrix (estimate_stim_ensemble: thousands of 64 ... 10^4-shot circuits per job,
                // src/device.jl:178-229): packed into pinned staging, all of them uploaded by ONE copy
                packed[i] = 1;
                arena_off[i] = pack_bytes;
                pack_bytes += pitch * B;
                continue;
            }
            arena_off[i] = arena_bytes;
            arena_bytes += pitch * B;
            if (layout == 1) {
                src_dev[i] = dev;
                if (!dev) {  // host records: staged contiguously, 16-byte aligned per item
                    raw_off[i] = raw_bytes;
                    raw_bytes += ((size_t)rows[i] * (size_t)ld[i] + 255) & ~(size_t)255;
                }
            }
        }
    }
    const int64_t total_out = item_off[n_items];
    if (total_out == 0) return ACES_OK;
    const bool out_dev = async_only || aces_is_device_ptr(odd_counts);

    // ---- stream table: one entry per (item, non-empty budget segment, Pauli slice) -------------
    // v1 (records wider than the v2 tiling, ungrouped tables only) holds 4 Paulis per thread
    const int slice = use_v2 ? 1024 : NT * 4;
    ACES_CHECK(ctx, use_v2 || !t->grouped, "parity_counts: grouped tables need the v2 kernel");
    static const std::array<double, 5> cost_w = [] {
        std::array<double, 5> w{1300.0, 114.0, 125.0, 210.0, 400.0};  // fixed, per column, per weight-1 / weight-2 / wider round
        if (const char* e = getenv("ACES_K1_COST"))  // development knob
            sscanf(e, "%lf,%lf,%lf,%lf,%lf", &w[0], &w[1], &w[2], &w[3], &w[4]);
        return w;
    }();
    std::vector<StreamDesc> streams;
    std::vector<double> stream_cost;  // estimated warp instructions per tile (cost-balanced partition below)
    std::vector<int64_t> item_start(1, 0);
    if (pack_bytes > 0) {
        // the packed matrices sit behind the others in the arena
        for (int i = 0; i < n_items; ++i)
            if (packed[i]) arena_off[i] += arena_bytes;
        ACES_TRY(aces_reserve_host(ctx, ctx->h_pack, pack_bytes));
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_arena, arena_bytes + pack_bytes));
        uint8_t* hp = (uint8_t*)ctx->h_pack.ptr;
        for (int i = 0; i < n_items; ++i) {
            if (!packed[i]) continue;
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            uint8_t* dst = hp + (arena_off[i] - arena_bytes);
            for (int c = 0; c < B; ++c) memcpy(dst + (size_t)c * pitch, counts[i] + (size_t)c * (size_t)ld[i], (size_t)rows[i]);
        }
        // blocking entry point: the copy has completed before the next call can touch h_pack
        ACES_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->d_arena.ptr + arena_bytes, hp, pack_bytes, cudaMemcpyHostToDevice,
                                       ctx->stream));
    }
    if (arena_bytes > 0) ACES_TRY(aces_reserve_dev(ctx, ctx->d_arena, arena_bytes + pack_bytes));
    if (layout == 1 && arena_bytes > 0) {
        // row-major records: contiguous copies into the raw staging area, then ONE batched transposition
        if (raw_bytes > 0) ACES_TRY(aces_reserve_dev(ctx, ctx->d_raw, raw_bytes));
        std::vector<TransDesc> td;
        int64_t max_rows = 0;
        for (int i = 0; i < n_items; ++i) {
            if (rows[i] == 0) continue;
            const uint8_t* src = counts[i];
            if (!src_dev[i]) {
                uint8_t* raw = (uint8_t*)ctx->d_raw.ptr + raw_off[i];
                ACES_CUDA(ctx, cudaMemcpyAsync(raw, counts[i], (size_t)rows[i] * (size_t)ld[i], cudaMemcpyHostToDevice,
                                               ctx->stream));
                src = raw;
            }
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            td.push_back(TransDesc{src, ld[i], rows[i], (uint8_t*)ctx->d_arena.ptr + arena_off[i], (int64_t)pitch});
            max_rows = std::max(max_rows, rows[i]);
        }
        if (!td.empty()) {
            ACES_TRY(aces_reserve_dev(ctx, ctx->d_tdesc, sizeof(TransDesc) * td.size()));
            // pageable source: the runtime stages it before returning, `td` may die afterwards
            ACES_CUDA(ctx, cudaMemcpyAsync(ctx->d_tdesc.ptr, td.data(), sizeof(TransDesc) * td.size(), cudaMemcpyHostToDevice,
                                           ctx->stream));
            const size_t smem = (size_t)TR * (size_t)B;
            ACES_CHECK(ctx, smem <= (size_t)ctx->smem_optin, "parity_counts: %d bytes per shot exceed the transposition tile", B);
            ACES_CUDA(ctx, cudaFuncSetAttribute(rowmajor_to_colmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t tiles = (max_rows + TR - 1) / TR;
            const int64_t want = std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / (int64_t)td.size());
            dim3 grid((unsigned)std::min<int64_t>(tiles, want), (unsigned)td.size());
            rowmajor_to_colmajor_kernel<<<grid, 256, smem, ctx->stream>>>((const TransDesc*)ctx->d_tdesc.ptr, B);
            ACES_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1;
        }
    }
    for (int i = 0; i < n_items; ++i) {
        const int e = experiment_ids[i];
        const int g0 = t->grp_ptr[(size_t)e], g1 = t->grp_ptr[(size_t)e + 1];
        if (rows[i] == 0 || g0 == g1) continue;
        const uint8_t* base;
        int64_t ldd;
        if (in_place[i]) {
            base = counts[i];
            ldd = ld[i];
        } else {
            const size_t pitch = ((size_t)rows[i] + 127) & ~(size_t)127;
            uint8_t* dst = (uint8_t*)ctx->d_arena.ptr + arena_off[i];
            if (layout == 0 && !packed[i])
                ACES_CUDA(ctx, cudaMemcpy2DAsync(dst, pitch, counts[i], (size_t)ld[i], (size_t)rows[i],
                                                 (size_t)B, cudaMemcpyDefault, ctx->stream));
            base = dst;
            ldd = (int64_t)pitch;
        }
        int64_t prev = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t s = prefix_shots[(int64_t)i * K + k];
            if (s > prev) {
                for (int g = g0; g < g1; ++g) {
                    const ColGroup& cg = t->groups[(size_t)g];
                    for (int p0 = 0; p0 < cg.E; p0 += slice) {
                        StreamDesc sd;
                        sd.base = base;
                        sd.ld = ldd;
                        sd.lo = prev;
                        sd.hi = s;
                        sd.copy_rows = (rows[i] + 15) & ~(int64_t)15;
                        sd.tile_first = prev / TS;
                        sd.out_off = item_off[i] + (int64_t)k * item_E[i];
                        sd.pauli_off = cg.pauli_off + p0;
                        sd.E = std::min(slice, cg.E - p0);
                        sd.ncols = cg.ncols;
                        sd.cols_off = cg.cols_off;
                        sd.pad = 0;
                        const int64_t ntiles = (s - 1) / TS - prev / TS + 1;
                        streams.push_back(sd);
                        stream_cost.push_back(use_v2 ? cost_w[0] + cost_w[1] * cg.ncols + cost_w[2] * cg.r1 + cost_w[3] * cg.r2 +
                                                           cost_w[4] * (cg.r4 + cg.rg)
                                                     : 1.0);
                        item_start.push_back(item_start.back() + ntiles);
                    }
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
    // ---- cost-balanced partition of the work items over the persistent CTAs (v2) ----------------
    // A CTA takes a contiguous range of tiles (a stream change costs a counter flush), and a tile's
    // cost depends on its stream: the transposition is proportional to the byte columns read, the
    // parity evaluation to the rounds of 32 Paulis.  Weights in warp instructions per 1024-shot
    // tile, from the ncu source view of the d = 9 kernel (profiles/k1_r01_ncu_full_summary.txt).
    std::vector<int64_t> cta_start;
    int grid = 0;
    static const bool balance = [] { const char* e = getenv("ACES_K1_BALANCE"); return !e || atoi(e) != 0; }();
    if (use_v2 && total_items > 0 && balance) {
        grid = v2_grid(ctx, cfg2);
        if (grid < 0) return grid;
        if ((int64_t)grid > total_items) grid = (int)total_items;
        std::vector<double> cum((size_t)n_streams + 1, 0.0), per((size_t)n_streams, 0.0);
        for (int s = 0; s < n_streams; ++s) {
            per[(size_t)s] = stream_cost[(size_t)s];
            cum[(size_t)s + 1] = cum[(size_t)s] + per[(size_t)s] * (double)(item_start[(size_t)s + 1] - item_start[(size_t)s]);
        }
        cta_start.assign((size_t)grid + 1, 0);
        int s = 0;
        for (int b = 1; b < grid; ++b) {
            const double target = cum[(size_t)n_streams] * (double)b / (double)grid;
            while (s + 1 < n_streams && cum[(size_t)s + 1] <= target) ++s;
            int64_t it = item_start[(size_t)s] + (int64_t)((target - cum[(size_t)s]) / per[(size_t)s]);
            it = std::min(it, item_start[(size_t)s + 1]);
            cta_start[(size_t)b] = std::max(it, cta_start[(size_t)b - 1]);
        }
        cta_start[(size_t)grid] = total_items;
    }
    const size_t sz_cta = sizeof(int64_t) * cta_start.size();
    const size_t o_streams = 0, o_istart = al(sz_streams), o_ioff = o_istart + al(sz_istart),
                 o_iE = o_ioff + al(sz_ioff), o_cta = o_iE + al(sz_iE), tbl_bytes = o_cta + al(sz_cta);
    // The tables are assembled in a host vector and uploaded only when they differ from what the
    // device already holds (a repetition loop re-fills the same sample buffers: identical tables).
    // Two pinned staging slots are used alternately: before a slot is refilled the copy that last
    // read it must have executed; the device-side table is ordered by the stream.
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_streams, tbl_bytes));
    uint8_t* ds = (uint8_t*)ctx->d_streams.ptr;
    std::vector<uint8_t>& tb = ctx->tbl_scratch;
    tb.assign(tbl_bytes, 0);
    if (n_streams) memcpy(tb.data() + o_streams, streams.data(), sizeof(StreamDesc) * n_streams);
    memcpy(tb.data() + o_istart, item_start.data(), sz_istart);
    memcpy(tb.data() + o_ioff, item_off.data(), sz_ioff);
    memcpy(tb.data() + o_iE, item_E.data(), sz_iE);
    if (sz_cta) memcpy(tb.data() + o_cta, cta_start.data(), sz_cta);
    const bool tables_cached = ctx->tbl_dev == ds && ctx->tbl_uploaded == tb;
    if (!tables_cached) {
        ctx->stage_slot ^= 1;
        DevBuf& hbuf = ctx->stage_slot ? ctx->h_stage2 : ctx->h_stage;
        cudaEvent_t hev = ctx->stage_slot ? ctx->ev2 : ctx->ev;
        ACES_CUDA(ctx, cudaEventSynchronize(hev));
        ACES_TRY(aces_reserve_host(ctx, hbuf, tbl_bytes));
        memcpy(hbuf.ptr, tb.data(), tbl_bytes);
        ACES_CUDA(ctx, cudaMemcpyAsync(ds, hbuf.ptr, tbl_bytes, cudaMemcpyHostToDevice, ctx->stream));
        ACES_CUDA(ctx, cudaEventRecord(hev, ctx->stream));
        ctx->tbl_uploaded.swap(tb);
        ctx->tbl_dev = ds;
    }

    // Per-segment counters: zeroed when (re)allocated; the prefix kernel clears what it has read,
    // so that a call leaves them clean for the next one (no memset per call).
    {
        const size_t need = sizeof(unsigned long long) * (size_t)total_out;
        const void* before = ctx->d_seg.ptr;
        const size_t cap_before = ctx->d_seg.cap;
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_seg, need));
        if (ctx->d_seg.ptr != before || ctx->d_seg.cap != cap_before || !ctx->seg_clean)
            ACES_CUDA(ctx, cudaMemsetAsync(ctx->d_seg.ptr, 0, ctx->d_seg.cap, ctx->stream));
        ctx->seg_clean = false;
    }

    if (total_items > 0) {
        ParityParams P;
        P.streams = (const StreamDesc*)(ds + o_streams);
        P.item_start = (const int64_t*)(ds + o_istart);
        P.pauli = t->d_pauli;
        P.bits = t->d_bits;
        P.cols = t->d_cols;
        P.seg = (unsigned long long*)ctx->d_seg.ptr;
        P.total_items = total_items;
        P.n_streams = n_streams;
        P.B = use_v2 ? std::max(t->max_group_cols, 1) : B;
        P.NS = use_v2 ? cfg2.NS : cfg.NS;
        P.cta_start = sz_cta ? (const int64_t*)(ds + o_cta) : nullptr;
        P.grid = grid;
        const int ppt = max_E <= NT ? 1 : (max_E <= 2 * NT ? 2 : 4);
        int rc;
        if (use_v2) {
            rc = v2_launch(ctx, P, cfg2);
        } else {
#define ACES_LAUNCH(TSV)                                              \
    (ppt == 1 ? launch_parity<TSV, 1>(ctx, P, cfg)                    \
              : ppt == 2 ? launch_parity<TSV, 2>(ctx, P, cfg) : launch_parity<TSV, 4>(ctx, P, cfg))
        if (TS == 4096) rc = ACES_LAUNCH(4096);
        else if (TS == 1024) rc = ACES_LAUNCH(1024);
        else rc = ACES_LAUNCH(256);
#undef ACES_LAUNCH
        }
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
        xchg::PushArgs push;
        if (ctx->xch && async_only) {
            // the exchange attached to this context: the prefix kernel publishes the counts to every rank
            aces_exchange* x = ctx->xch;
            ACES_CHECK(ctx, !accumulate && x->nc == total_out,
                       "parity_counts: the attached exchange holds %lld counts per rank, this call produces %lld%s",
                       (long long)x->nc, (long long)total_out, accumulate ? " (and accumulate is not supported)" : "");
            x->epoch += 1;
            push.base = x->d_base; push.world = x->world; push.rank = x->rank; push.n = x->n; push.epoch = x->epoch;
            push.ticket = x->d_ticket;
        }
        prefix_kernel<<<grid, 256, 0, ctx->stream>>>((unsigned long long*)ctx->d_seg.ptr, d_out,
                                                    (const int64_t*)(ds + o_ioff),
                                                    (const int32_t*)(ds + o_iE), n_items, K, acc_dev, push);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        ctx->seg_clean = true;
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
    const int B = (n_qubits + 7) / 8;
    std::vector<PauliDesc> descs;
    descs.reserve((size_t)P);
    std::vector<int32_t> bits, cols_all;
    bits.reserve((size_t)pauli_ptr[P] + 4);
    std::vector<ColGroup> groups;
    std::vector<int32_t> grp_ptr(1, 0);
    int max_E = 0, max_group_E = 0, max_group_cols = 0;

    // ---- pass 1: supports reduced mod 2 (a bit listed twice cancels: the per-shot sum is reduced
    // mod 2, src/estimate.jl:251) and the decision whether the tables are grouped ----------------
    std::vector<std::vector<std::vector<int32_t>>> sups((size_t)n_experiments);
    std::vector<int32_t> sup;
    bool need_groups = B > 41;  // wider than the widest single-group kernel variant
    // Records that fit one column group run the 16-round kernel variant (three CTAs per SM up to 21
    // byte columns, two above), so an experiment of up to 512 Paulis (a full-covariance design:
    // eigenvalue and covariance Paulis together) stays one group and its columns are transposed once.
    const char* gp_env = getenv("ACES_K1_GROUP_PAULIS");  // development knob
    const int group_paulis = gp_env ? atoi(gp_env) : (B <= GROUP_COLS ? 2 * GROUP_PAULIS : GROUP_PAULIS);
    bool groupable = true;      // every Pauli fits one group of GROUP_COLS columns
    for (int e = 0; e < n_experiments; ++e) {
        ACES_CHECK(ctx, exp_ptr[e + 1] >= exp_ptr[e], "tables_create: exp_ptr not monotone at %d", e);
        const int64_t p0 = exp_ptr[e], p1 = exp_ptr[e + 1];
        ACES_CHECK(ctx, p1 - p0 < (int64_t)1 << 30, "tables_create: experiment %d too large", e);
        const int E = (int)(p1 - p0);
        max_E = std::max(max_E, E);
        sups[(size_t)e].assign((size_t)E, std::vector<int32_t>());
        int n_dev = 0;
        for (int p = 0; p < E; ++p) {
            const int64_t b0 = pauli_ptr[p0 + p], b1 = pauli_ptr[p0 + p + 1];
            ACES_CHECK(ctx, b1 >= b0, "tables_create: pauli_ptr not monotone at %lld", (long long)(p0 + p));
            sup.assign(pauli_bits + b0, pauli_bits + b1);
            for (int32_t b : sup)
                ACES_CHECK(ctx, b >= 0 && b < n_qubits, "tables_create: record bit %d out of range [0,%d)", b, n_qubits);
            std::sort(sup.begin(), sup.end());
            std::vector<int32_t>& red = sups[(size_t)e][(size_t)p];
            for (size_t i = 0; i < sup.size();) {
                size_t j = i;
                while (j < sup.size() && sup[j] == sup[i]) ++j;
                if ((j - i) & 1) red.push_back(sup[i]);
                i = j;
            }
            if (!red.empty()) {
                ++n_dev;
                int ncol = 1;
                for (size_t i = 1; i < red.size(); ++i) ncol += (red[i] >> 3) != (red[i - 1] >> 3);
                if (ncol > GROUP_COLS) groupable = false;
            }
        }
        if (n_dev > group_paulis) need_groups = true;
    }
    const bool grouped = need_groups && groupable;

    // ---- pass 2: groups (membership and columns only; the device order needs the kernel variant,
    // which depends on the widest group of the whole table) ---------------------------------------
    struct PendingGroup {
        int e;
        std::vector<int> mem;
        std::vector<int32_t> gcols;
    };
    std::vector<PendingGroup> pending;
    std::vector<int> members, local_col((size_t)B, -1);
    auto emit_group = [&](int e, const std::vector<int>& mem, const std::vector<int32_t>& gcols) -> int {
        pending.push_back(PendingGroup{e, mem, gcols});
        max_group_E = std::max(max_group_E, (int)mem.size());
        max_group_cols = std::max(max_group_cols, (int)gcols.size());
        groups.push_back(ColGroup{});  // filled in pass 3
        return ACES_OK;
    };
    // Pass 3 (after the loop below).  Device order: by weight, so that the lanes of a warp run the
    // same number of plane loads.  Inside the weight-1 and weight-2 classes the order is chosen
    // against shared-memory bank conflicts: a 128-bit load is served a quarter warp at a time, and
    // the plane row pitch is an odd number of 16-byte vectors, so eight lanes are conflict-free
    // exactly when their plane rows differ mod 8.  Real designs pair distant qubits (a data qubit
    // and its ancilla), so sorting by the first bit leaves the second operand's rows colliding;
    // instead every aligned block of eight lanes is filled greedily with Paulis whose first rows
    // are pairwise different mod 8 and whose second rows are too (the operands of a weight-2 Pauli
    // may be swapped: XOR commutes).  BMAXV = plane rows per record-bit position of the kernel
    // variant (parity_v2.cuh: row = (bit & 7) * BMAX + (bit >> 3)).
    auto build_group = [&](size_t gi, int BMAXV) -> int {
        const PendingGroup& pg = pending[gi];
        const int e = pg.e;
        ColGroup cg;
        cg.pauli_off = (int64_t)descs.size();
        cg.E = (int32_t)pg.mem.size();
        cg.cols_off = (int32_t)cols_all.size();
        cg.ncols = (int32_t)pg.gcols.size();
        for (size_t c = 0; c < pg.gcols.size(); ++c) local_col[(size_t)pg.gcols[c]] = (int)c;
        cols_all.insert(cols_all.end(), pg.gcols.begin(), pg.gcols.end());
        auto local_bit = [&](int32_t q) -> int32_t { return 8 * local_col[(size_t)(q >> 3)] + (q & 7); };
        auto row_of = [&](int32_t lb) -> int { return (lb & 7) * BMAXV + (lb >> 3); };
        std::vector<std::pair<int64_t, int32_t>> byw;
        for (int p : pg.mem) {
            const std::vector<int32_t>& red = sups[(size_t)e][(size_t)p];
            const int32_t lb = local_bit(red.front());
            const int64_t row = (int64_t)(lb & 7) * 64 + (lb >> 3);  // 64 >= any BMAX: same order
            byw.push_back({((int64_t)red.size() << 32) | row, p});
        }
        std::sort(byw.begin(), byw.end());
        struct Slot { int32_t p; bool swap; };
        std::vector<Slot> order;
        order.reserve(byw.size());
        size_t n_w1 = 0, n_w2 = 0;
        for (const auto& wp : byw) {
            const size_t w = sups[(size_t)e][(size_t)wp.second].size();
            n_w1 += w == 1;
            n_w2 += w == 2;
        }
        {
            // rounds by class, as parity_kernel_v2 classifies them (the table is sorted by weight)
            size_t n_w4 = 0;
            for (const auto& wp : byw) {
                const size_t w = sups[(size_t)e][(size_t)wp.second].size();
                n_w4 += (w == 3 || w == 4);
            }
            const int R = (int)((byw.size() + 31) / 32);
            cg.r1 = (int)(n_w1 / 32);
            if (n_w1 == byw.size()) cg.r1 = R;  // a last, partial round of weight-1 Paulis only
            const int c12 = (n_w1 + n_w2 == byw.size()) ? R : (int)((n_w1 + n_w2) / 32);
            const int c14 = (n_w1 + n_w2 + n_w4 == byw.size()) ? R : (int)((n_w1 + n_w2 + n_w4) / 32);
            cg.r2 = std::max(c12 - cg.r1, 0);
            cg.r4 = std::max(c14 - std::max(c12, cg.r1), 0);
            cg.rg = std::max(R - std::max(c14, std::max(c12, cg.r1)), 0);
        }
        if (BMAXV > 0 && n_w1 + n_w2 > 8) {
            // pools in sorted order; `taken` marks what has been placed
            std::vector<int> ra(n_w1 + n_w2), rb(n_w1 + n_w2, -1);
            for (size_t i = 0; i < n_w1 + n_w2; ++i) {
                const std::vector<int32_t>& red = sups[(size_t)e][(size_t)byw[i].second];
                ra[i] = row_of(local_bit(red[0])) & 7;
                if (red.size() == 2) rb[i] = row_of(local_bit(red[1])) & 7;
            }
            std::vector<char> taken(n_w1 + n_w2, 0);
            const int zero_res = (8 * BMAXV) & 7;  // the all-zero row that stands in for a missing operand
            int cnt1[8] = {0}, cnt2[8][8] = {{0}};
            for (size_t i = 0; i < n_w1; ++i) cnt1[ra[i]]++;
            for (size_t i = n_w1; i < n_w1 + n_w2; ++i) cnt2[ra[i]][rb[i]]++;
            unsigned used_a = 0, used_b = 0;
            for (size_t pos = 0; pos < n_w1 + n_w2; ++pos) {
                if ((pos & 7) == 0) {
                    used_a = used_b = 0;
                    // a block that mixes the two classes sits in a weight-2 round: its weight-1 lanes
                    // load the zero row as their second operand
                    if (pos < n_w1 && pos + 8 > n_w1 && n_w2 > 0) used_b = 1u << zero_res;
                }
                int best = -1, best_score = -1;
                bool best_swap = false;
                if (pos < n_w1) {
                    for (size_t i = 0; i < n_w1; ++i) {
                        if (taken[i]) continue;
                        const int free_a = !(used_a >> ra[i] & 1);
                        const int score = free_a * 1000 + cnt1[ra[i]];
                        if (score > best_score) { best_score = score; best = (int)i; }
                    }
                    if (best >= 0) cnt1[ra[best]]--;
                } else {
                    for (size_t i = n_w1; i < n_w1 + n_w2; ++i) {
                        if (taken[i]) continue;
                        for (int sw = 0; sw < 2; ++sw) {
                            const int a = sw ? rb[i] : ra[i], b = sw ? ra[i] : rb[i];
                            const int free_a = !(used_a >> a & 1), free_b = !(used_b >> b & 1);
                            const int score = (free_a + free_b) * 1000 + cnt2[ra[i]][rb[i]];
                            if (score > best_score) { best_score = score; best = (int)i; best_swap = sw != 0; }
                        }
                    }
                    if (best >= 0) cnt2[ra[best]][rb[best]]--;
                }
                if (best < 0) return ACES_E_INVALID;  // cannot happen: the pool of this class is not empty
                taken[(size_t)best] = 1;
                const int a = best_swap ? rb[best] : ra[best], b = best_swap ? ra[best] : rb[best];
                used_a |= 1u << a;
                if (b >= 0) used_b |= 1u << b;
                order.push_back(Slot{byw[(size_t)best].second, best_swap});
            }
            for (size_t i = n_w1 + n_w2; i < byw.size(); ++i) order.push_back(Slot{byw[i].second, false});
        } else {
            for (const auto& wp : byw) order.push_back(Slot{wp.second, false});
        }
        for (const Slot& sl : order) {
            const std::vector<int32_t>& red = sups[(size_t)e][(size_t)sl.p];
            PauliDesc d;
            d.w = (int32_t)red.size();
            d.ext = (int32_t)bits.size();
            d.orig = sl.p;
            d.pad = 0;
            for (int i = 0; i < 4; ++i) d.q[i] = 0;
            for (size_t i = 0; i < red.size(); ++i) {
                const int32_t lb = local_bit(red[i]);
                if (i < 4) d.q[i] = lb;
                bits.push_back(lb);
            }
            if (sl.swap) std::swap(d.q[0], d.q[1]);
            if (bits.size() >= ((size_t)1 << 31)) return ACES_E_INVALID;
            descs.push_back(d);
        }
        groups[gi] = cg;
        return ACES_OK;
    };
    for (int e = 0; e < n_experiments; ++e) {
        const int E = (int)sups[(size_t)e].size();
        members.clear();
        for (int p = 0; p < E; ++p)
            if (!sups[(size_t)e][(size_t)p].empty()) members.push_back(p);
        if (!members.empty()) {
            if (!grouped) {
                // one group over all byte columns: local bits are the record bits
                std::vector<int32_t> all((size_t)B);
                for (int c = 0; c < B; ++c) all[(size_t)c] = c;
                ACES_CHECK(ctx, emit_group(e, members, all) == ACES_OK, "tables_create: table too large");
            } else {
                // Every group re-reads the columns it shares with another group, so the Paulis are
                // partitioned to share as few as possible.  (1) The byte columns are put in an order
                // that keeps columns joined by multi-column Paulis close together: starting from the
                // lowest column, the next one is the unvisited column most strongly connected to the
                // recently visited ones (a real design pairs a data qubit with an ancilla about n/2
                // away: the order then alternates data and ancilla columns along the band).  (2) Each
                // Pauli is keyed by the last position any of its columns takes in that order, and the
                // sorted list is cut into groups of at most GROUP_COLS columns and an equal share of
                // the Paulis.  (3) Local search: a column that a group needs only for Paulis which fit
                // (columns and capacity) into other groups is dropped by moving those Paulis.
                const size_t nm = members.size();
                std::vector<std::vector<int32_t>> pc(nm);
                for (size_t i = 0; i < nm; ++i) {
                    for (int32_t q : sups[(size_t)e][(size_t)members[i]]) pc[i].push_back(q >> 3);
                    pc[i].erase(std::unique(pc[i].begin(), pc[i].end()), pc[i].end());  // supports are sorted
                }
                std::vector<std::vector<std::pair<int32_t, int32_t>>> nbr((size_t)B);
                {
                    std::map<std::pair<int32_t, int32_t>, int32_t> w;
                    for (size_t i = 0; i < nm; ++i)
                        for (size_t a = 0; a < pc[i].size(); ++a)
                            for (size_t b = a + 1; b < pc[i].size(); ++b) w[{pc[i][a], pc[i][b]}] += 1;
                    for (const auto& kv : w) {
                        nbr[(size_t)kv.first.first].push_back({kv.first.second, kv.second});
                        nbr[(size_t)kv.first.second].push_back({kv.first.first, kv.second});
                    }
                }
                std::vector<char> used_col((size_t)B, 0), visited((size_t)B, 0);
                for (size_t i = 0; i < nm; ++i)
                    for (int32_t c : pc[i]) used_col[(size_t)c] = 1;
                std::vector<int32_t> pos((size_t)B, -1);
                std::vector<double> score((size_t)B, 0.0);
                int n_used = 0;
                for (int c = 0; c < B; ++c) n_used += used_col[(size_t)c];
                for (int k = 0; k < n_used; ++k) {
                    int best = -1;
                    for (int c = 0; c < B; ++c)
                        if (used_col[(size_t)c] && !visited[(size_t)c] && score[(size_t)c] > 0.0 &&
                            (best < 0 || score[(size_t)c] > score[(size_t)best]))
                            best = c;
                    if (best < 0)
                        for (int c = 0; c < B && best < 0; ++c)
                            if (used_col[(size_t)c] && !visited[(size_t)c]) best = c;
                    visited[(size_t)best] = 1;
                    pos[(size_t)best] = k;
                    for (int c = 0; c < B; ++c) score[(size_t)c] *= 0.5;  // the band follows the frontier
                    for (const auto& nb : nbr[(size_t)best])
                        if (!visited[(size_t)nb.first]) score[(size_t)nb.first] += (double)nb.second;
                }
                std::vector<size_t> idx(nm);
                std::vector<int32_t> last(nm, 0);
                for (size_t i = 0; i < nm; ++i) {
                    idx[i] = i;
                    for (int32_t c : pc[i]) last[i] = std::max(last[i], pos[(size_t)c]);
                }
                std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) {
                    if (last[x] != last[y]) return last[x] < last[y];
                    return sups[(size_t)e][(size_t)members[x]].front() < sups[(size_t)e][(size_t)members[y]].front();
                });
                const int n_by_count = ((int)nm + group_paulis - 1) / group_paulis;
                const int share = ((int)nm + n_by_count - 1) / n_by_count;
                std::vector<std::vector<size_t>> gm;   // members (indices into `members`) of every group
                {
                    std::vector<size_t> cur;
                    std::vector<int32_t> gcols;
                    for (size_t i : idx) {
                        std::vector<int32_t> merged = gcols;
                        merged.insert(merged.end(), pc[i].begin(), pc[i].end());
                        std::sort(merged.begin(), merged.end());
                        merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
                        if (!cur.empty() && ((int)merged.size() > GROUP_COLS || (int)cur.size() >= share)) {
                            gm.push_back(cur);
                            cur.clear();
                            merged = pc[i];
                        }
                        cur.push_back(i);
                        gcols.swap(merged);
                    }
                    if (!cur.empty()) gm.push_back(cur);
                }
                // local search (3): per group, users of every column
                {
                    const size_t G = gm.size();
                    std::vector<std::vector<int32_t>> cnt(G, std::vector<int32_t>((size_t)B, 0));
                    for (size_t g = 0; g < G; ++g)
                        for (size_t i : gm[g])
                            for (int32_t c : pc[i]) cnt[g][(size_t)c] += 1;
                    bool improved = G > 1;
                    int guard = 0;
                    while (improved && ++guard < 4 * B * (int)G) {
                        improved = false;
                        for (size_t g = 0; g < G && !improved; ++g) {
                            std::vector<int32_t> cols_g;
                            for (int c = 0; c < B; ++c)
                                if (cnt[g][(size_t)c] > 0) cols_g.push_back(c);
                            std::sort(cols_g.begin(), cols_g.end(), [&](int32_t x, int32_t y) {
                                return cnt[g][(size_t)x] != cnt[g][(size_t)y] ? cnt[g][(size_t)x] < cnt[g][(size_t)y] : x < y;
                            });
                            for (int32_t c : cols_g) {
                                std::vector<int32_t> cap(G);
                                for (size_t h = 0; h < G; ++h) cap[h] = group_paulis - (int32_t)gm[h].size();
                                std::vector<std::pair<size_t, size_t>> plan;  // (member, destination group)
                                bool ok = true;
                                for (size_t i : gm[g]) {
                                    if (std::find(pc[i].begin(), pc[i].end(), c) == pc[i].end()) continue;
                                    bool placed = false;
                                    for (size_t h = 0; h < G && !placed; ++h) {
                                        if (h == g || cap[h] <= 0) continue;
                                        bool fits = true;
                                        for (int32_t x : pc[i]) fits = fits && cnt[h][(size_t)x] > 0;
                                        if (fits) {
                                            plan.push_back({i, h});
                                            cap[h] -= 1;
                                            placed = true;
                                        }
                                    }
                                    if (!placed) { ok = false; break; }
                                }
                                if (!ok || plan.empty()) continue;
                                for (const auto& mv : plan) {
                                    gm[g].erase(std::find(gm[g].begin(), gm[g].end(), mv.first));
                                    gm[mv.second].push_back(mv.first);
                                    for (int32_t x : pc[mv.first]) {
                                        cnt[g][(size_t)x] -= 1;
                                        cnt[mv.second][(size_t)x] += 1;
                                    }
                                }
                                improved = true;
                                break;
                            }
                        }
                    }
                }
                for (const auto& g : gm) {
                    if (g.empty()) continue;
                    std::vector<int> mem;
                    std::vector<int32_t> gcols;
                    for (size_t i : g) {
                        mem.push_back(members[i]);
                        gcols.insert(gcols.end(), pc[i].begin(), pc[i].end());
                    }
                    std::sort(mem.begin(), mem.end());
                    std::sort(gcols.begin(), gcols.end());
                    gcols.erase(std::unique(gcols.begin(), gcols.end()), gcols.end());
                    ACES_CHECK(ctx, emit_group(e, mem, gcols) == ACES_OK, "tables_create: table too large");
                }
            }
        }
        grp_ptr.push_back((int32_t)groups.size());
    }
    {
        V2Cfg vc;
        const bool v2 = max_group_E <= 1024 &&
                        v2_pick(std::max(max_group_cols, 1), std::max(max_group_E, 1), ctx->smem_optin, &vc);
        for (size_t gi = 0; gi < pending.size(); ++gi)
            ACES_CHECK(ctx, build_group(gi, v2 ? vc.BMAX : 0) == ACES_OK, "tables_create: table too large");
    }
    aces_tables* t = new aces_tables();
    t->n_qubits = n_qubits;
    t->B = B;
    t->n_experiments = n_experiments;
    t->max_E = max_E;
    t->exp_ptr.assign(exp_ptr, exp_ptr + n_experiments + 1);
    t->grp_ptr = grp_ptr;
    t->groups = groups;
    t->max_group_E = max_group_E;
    t->max_group_cols = max_group_cols;
    t->grouped = grouped;
    cudaError_t e1 = cudaMalloc(&t->d_pauli, sizeof(PauliDesc) * std::max<size_t>(descs.size(), 1));
    cudaError_t e2 = cudaMalloc(&t->d_bits, sizeof(int32_t) * std::max<size_t>(bits.size(), 1));
    cudaError_t e3 = cudaMalloc(&t->d_cols, sizeof(int32_t) * std::max<size_t>(cols_all.size(), 1));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaGetLastError();
        if (t->d_pauli) cudaFree(t->d_pauli);
        if (t->d_bits) cudaFree(t->d_bits);
        if (t->d_cols) cudaFree(t->d_cols);
        delete t;
        return aces_fail(ctx, ACES_E_NOMEM, "tables_create: cudaMalloc failed");
    }
    if (!descs.empty())
        ACES_CUDA(ctx, cudaMemcpy(t->d_pauli, descs.data(), sizeof(PauliDesc) * descs.size(),
                                  cudaMemcpyHostToDevice));
    if (!bits.empty())
        ACES_CUDA(ctx, cudaMemcpy(t->d_bits, bits.data(), sizeof(int32_t) * bits.size(),
                                  cudaMemcpyHostToDevice));
    if (!cols_all.empty())
        ACES_CUDA(ctx, cudaMemcpy(t->d_cols, cols_all.data(), sizeof(int32_t) * cols_all.size(),
                                  cudaMemcpyHostToDevice));
    *out = t;
    return ACES_OK;
}

int64_t aces_tables_group_columns(const aces_tables* t, int32_t experiment) {
    if (!t || experiment < 0 || experiment >= t->n_experiments) return -1;
    int64_t n = 0;
    for (int g = t->grp_ptr[(size_t)experiment]; g < t->grp_ptr[(size_t)experiment + 1]; ++g) n += t->groups[(size_t)g].ncols;
    return n;
}

int32_t aces_tables_destroy(aces_ctx* ctx, aces_tables* t) {
    if (!ctx) return ACES_E_INVALID;
    if (!t) return ACES_OK;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t->d_pauli) cudaFree(t->d_pauli);
    if (t->d_bits) cudaFree(t->d_bits);
    if (t->d_cols) cudaFree(t->d_cols);
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

int32_t aces_parity_counts_rowmajor(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                                    const int32_t* experiment_ids, const uint8_t* const* records,
                                    const int64_t* rows, const int64_t* row_pitch, int32_t K,
                                    const int64_t* prefix_shots, int64_t* odd_counts, int32_t accumulate) {
    return parity_run(ctx, tables, n_items, experiment_ids, records, rows, row_pitch, K, prefix_shots,
                      odd_counts, accumulate, false, 1);
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
