// dense.cu -- FP64 dense engine: DMMA (mma.sync m8n8k4 f64, SASS DMMA) GEMM and the blocked
// Cholesky / triangular inverse / triangular solves built on it.  See dense.cuh for the
// conventions (column-major, padded to multiples of 128, no edge handling).
//
// Replaces, for the fit path, what the reference gets from LAPACK / SuiteSparse:
// cholesky + inv of the Gram matrix (src/merit.jl:564,586), the per-block covariance
// factorisations (src/merit.jl:398-402) and the least-squares solve (src/estimate.jl:500-502).
#include "dense.cuh"

#include <algorithm>
#include <functional>
#include <mutex>
#include <queue>
#include <vector>

namespace dense {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, GT = 256;

__device__ __forceinline__ int ld_acquire_fwd(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_fwd(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr int LD_MC = BM + 4;  // row stride of a [BK][BM] tile (operand contiguous along m / n)
constexpr int LD_KC = BK + 4;  // row stride of a [BM][BK] tile (operand contiguous along k)
constexpr int OPER_DOUBLES = BM * LD_KC;  // 2560 >= BK * LD_MC = 2112
constexpr int STAGE_DOUBLES = 2 * OPER_DOUBLES;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
// The 64-row variant keeps a compact A operand so that two of its CTAs fit one SM (92 KB each): the
// prologue / epilogue of one tile then overlaps the main loop of another (short-K updates).
__host__ __device__ constexpr int a_doubles(int bmt) { return bmt == BM ? OPER_DOUBLES : (bmt * LD_KC > BK * (bmt + 4) ? bmt * LD_KC : BK * (bmt + 4)); }
__host__ __device__ constexpr int stage_doubles(int bmt) { return a_doubles(bmt) + OPER_DOUBLES; }
constexpr size_t gemm_smem(int bmt) { return (size_t)STAGES * stage_doubles(bmt) * sizeof(double); }

struct GemmP {
    const double* A;
    const double* B;
    double* C;
    int64_t lda, ldb, ldc;
    int k;
    double alpha, beta;
};

__device__ __forceinline__ void cp_async16(double* dst, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// D (8x8) += A (8x4, row) * B (4x8, col) on the FP64 tensor pipe.
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// One CTA computes the BMT x 128 tile (bm, bn) of C (BMT = 128 or 64).  8 warps as 2 (m) x 4 (n),
// warp tile (BMT/2) x 32 = (BMT/16) x 4 DMMA tiles; operands staged by cp.async in a 3-stage ring.
// The 64-row variant doubles the number of CTAs of a launch: the skinny, short-K updates on the
// critical path of the Cholesky (a 128-row tile with K = 128 occupies one SM for >= 15 us).
// COH: the C tile is read past L1 (__ldcg) -- for the persistent factorisation kernel, where another SM
// may have updated the tile since this SM last touched it (operands always come through cp.async.cg).
template <bool A_KC, bool B_KC, int BMT, bool COH = false>
__device__ __forceinline__ void gemm_tile(const GemmP& p, const int bm, const int bn, double* sm) {
    constexpr int MI = BMT / 16;  // DMMA tiles per warp along m
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * (BMT / 2), wn = (warp >> 1) * 32;
    // A_KC: op(A)(i,k) = A[k + i*lda]; else A[i + k*lda].  Same for B with (j,k).
    const double* Ag = A_KC ? p.A + (int64_t)bm * BMT * p.lda : p.A + (int64_t)bm * BMT;
    const double* Bg = B_KC ? p.B + (int64_t)bn * BN * p.ldb : p.B + (int64_t)bn * BN;
    const int nk = p.k / BK;
    constexpr int LDA_MC = BMT + 4;            // row stride of the A tile in the [BK][BMT] layout
    constexpr int SD = stage_doubles(BMT);     // doubles per pipeline stage
    constexpr int AD = a_doubles(BMT);

    auto load = [&](int stage, int kt) {
        double* sA = sm + stage * SD;
        double* sB = sA + AD;
        const int k0 = kt * BK;
#pragma unroll
        for (int c = tid; c < BMT * 8; c += GT) {
            if (!A_KC) {
                const int kk = c / (BMT / 2), ci = c % (BMT / 2);
                cp_async16(sA + kk * LDA_MC + 2 * ci, Ag + 2 * ci + (int64_t)(k0 + kk) * p.lda);
            } else {
                const int i = c >> 3, ck = c & 7;
                cp_async16(sA + i * LD_KC + 2 * ck, Ag + (k0 + 2 * ck) + (int64_t)i * p.lda);
            }
        }
#pragma unroll
        for (int c = tid; c < 1024; c += GT) {
            if (!B_KC) {
                const int kk = c >> 6, ci = c & 63;
                cp_async16(sB + kk * LD_MC + 2 * ci, Bg + 2 * ci + (int64_t)(k0 + kk) * p.ldb);
            } else {
                const int j = c >> 3, ck = c & 7;
                cp_async16(sB + j * LD_KC + 2 * ck, Bg + (k0 + 2 * ck) + (int64_t)j * p.ldb);
            }
        }
    };

    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load(s, s);
        cp_async_commit();
    }
    // COH: the accumulators START from the C tile, acc = (beta / alpha) C, so that the result alpha * acc
    // = alpha A B' + beta C needs no read-modify-write pass at the end (64 dependent L2 round trips per
    // thread, a fifth of a K = 512 task); the loads are in flight together with the first operand
    // stages.  Exact for the |alpha| = |beta| = 1 updates this mode is used for.
    const bool preload = COH && p.beta != 0.0;
    if (preload) {
        const double sc = p.beta / p.alpha;
        const double* Cg0 = p.C + (int64_t)bm * BMT + (int64_t)bn * BN * p.ldc;
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double* c0 = Cg0 + (wm + i * 8 + g) + (int64_t)(wn + j * 8 + 2 * t) * p.ldc;
                acc[i][j][0] = sc * __ldcg(c0);
                acc[i][j][1] = sc * __ldcg(c0 + p.ldc);
            }
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load(nxt % STAGES, nxt);
        cp_async_commit();
        const double* sA = sm + (kt % STAGES) * SD;
        const double* sB = sA + AD;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[MI], b[4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
                a[i] = A_KC ? sA[(wm + i * 8 + g) * LD_KC + k4 * 4 + t]
                            : sA[(k4 * 4 + t) * LDA_MC + wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = B_KC ? sB[(wn + j * 8 + g) * LD_KC + k4 * 4 + t]
                            : sB[(k4 * 4 + t) * LD_MC + wn + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[g][2t], C[g][2t+1] of every 8x8 tile
    double* Cg = p.C + (int64_t)bm * BMT + (int64_t)bn * BN * p.ldc;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wm + i * 8 + g;
            const int c = wn + j * 8 + 2 * t;
            double* c0 = Cg + r + (int64_t)c * p.ldc;
            double* c1 = c0 + p.ldc;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            if (p.beta != 0.0 && !preload) {
                v0 += p.beta * *c0;
                v1 += p.beta * *c1;
            }
            *c0 = v0;
            *c1 = v1;
        }
}

template <bool A_KC, bool B_KC, int BMT>
__global__ void __launch_bounds__(GT, BMT <= 64 ? 2 : 1) gemm_kernel(const GemmP p, const int lower_only) {
    // lower_only: skip the tiles that lie entirely above the block diagonal
    if (lower_only && (int)blockIdx.y * BN >= ((int)blockIdx.x + 1) * BMT) return;
    extern __shared__ __align__(16) double sm[];
    gemm_tile<A_KC, B_KC, BMT>(p, blockIdx.x, blockIdx.y, sm);
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GT, 1) gemm_batched_kernel(const GemmDesc* __restrict__ descs) {
    const GemmDesc& d = descs[blockIdx.z];
    const int bm = blockIdx.x, bn = blockIdx.y;
    if (d.k == 0 || bm >= d.tiles_m || bn >= d.tiles_n || (d.lower_only && bn > bm)) return;
    GemmP p{d.A, d.B, d.C, d.lda, d.ldb, d.ldc, d.k, d.alpha, d.beta};
    extern __shared__ __align__(16) double sm[];
    gemm_tile<A_KC, B_KC, BM>(p, bm, bn, sm);
}

// ---- 128 x 128 diagonal block: Cholesky factor and its inverse, one CTA of 512 threads ---------
//
// The block lives in shared memory (column-major, odd row pitch).  Factorisation: right-looking
// with 16-wide panels -- the 16 x 16 diagonal piece by one warp, the panel rows one thread per
// row, the trailing update as static 4 x 8 register tiles per thread.  Inverse: in place, the
// eight 16 x 16 diagonal pieces by eight warps at once, then block columns from right to left
// (X21 = -X22 * L21 * X11).
constexpr int PLD = DB + 1;
constexpr int PW = 16;
constexpr int DIAG_THREADS = 512;
constexpr int INV_TMP = (DB / 2) * (DB / 2);  // scratch of the inverse: one 64 x 64 product
constexpr size_t DIAG_SMEM = (size_t)(DB * PLD + INV_TMP + 2 * DB) * sizeof(double);

// 1 / d to full precision without the division subroutine: hardware seed (about 20 bits) and two
// Newton steps.  d is a pivot that already passed the positivity test.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// Cholesky of the 16 x 16 piece at (k0, k0) of the shared-memory block by ONE warp, rows in
// registers (lane i holds row i; lanes 16..31 mirror 0..15 so that every shuffle source is
// defined).  Kept out of line so that its registers are allocated separately from the caller's
// tile accumulators.  Returns the number of pivots that were not above their floor; each was replaced
// by `repl` (1: the caller reports a failed factorisation; DEAD_PIVOT: the column is eliminated --
// its entries of L become ~1e-150 times the data, so the variable drops out of every later update
// and solve: the basic solution of a rank-deficient least-squares problem).
constexpr double DEAD_PIVOT = 1e300;
__device__ __noinline__ int diag16(double* s, double* rdiag, const double* pfloor, int k0, int lane, double repl) {
    const int li = lane & (PW - 1);
    double a[PW];
    int bad = 0;
#pragma unroll
    for (int c = 0; c < PW; ++c) a[c] = s[(k0 + c) * PLD + k0 + li];
    // Square-root-free elimination: column k stays unscaled (a_ik) while the update uses
    // a_ik * a_jk / d_k, so that the only operation on the critical path of a pivot is one
    // reciprocal; the 1/sqrt(d_k) scaling of all columns is applied once at the end.
    double dmine = 1.0;  // pivot of row li
#pragma unroll
    for (int k = 0; k < PW; ++k) {
        double dkk = __shfl_sync(0xffffffffu, a[k], k);
        if (!(dkk > pfloor[k0 + k])) {
            bad += 1;
            dkk = repl;
        }
        if (li == k) dmine = dkk;
        const double tk = a[k] * fast_rcp(dkk);  // a_ik / d_k
#pragma unroll
        for (int j = k + 1; j < PW; ++j) {
            const double ajk = __shfl_sync(0xffffffffu, a[k], j);
            if (li >= j) a[j] -= tk * ajk;
        }
    }
    const double rmine = rsqrt(dmine);
#pragma unroll
    for (int c = 0; c < PW; ++c) {
        const double rc = __shfl_sync(0xffffffffu, rmine, c);
        a[c] = (li == c) ? dmine * rc : a[c] * rc;  // L(c, c) = sqrt(d_c); L(i, c) = a_ic / sqrt(d_c)
    }
    if (lane < PW) {
        rdiag[k0 + lane] = rmine;
#pragma unroll
        for (int c = 0; c < PW; ++c)
            if (c <= li) s[(k0 + c) * PLD + k0 + li] = a[c];
    }
    return bad;
}

// Trailing update of one row (one lane) for NB columns c0, c0 + 15, ... of the shared-memory block:
// s(row, j) -= sum_k s(row, k0 + k) * s(j, k0 + k) for the 16 columns k of the current panel.
template <int NB, int CS>
__device__ __forceinline__ void c2_rows(double* s, const int k0, const int c0, const int row) {
    double acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < PW; ++k) {
        const double* col = s + (k0 + k) * PLD;
        const double ra = col[row];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] += ra * col[c0 + CS * b];
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = c0 + CS * b;
        if (row >= j) s[j * PLD + row] -= acc[b];
    }
}

// ---- in-place inverse of the lower-triangular factor in shared memory ---------------------------------
// Recursive doubling: the eight 16 x 16 diagonal pieces are inverted first (one warp each, columns in
// registers); then pairs of inverted blocks of size h = 16, 32, 64 are merged,
//     inv([L11 0; L21 L22]) = [X11 0; -X22 L21 X11, X22],
// all pairs of a level at once, every product as 4 x 4 register tiles (two shared loads per four
// multiply-adds).  Six block-wide barriers in all; the previous version walked the sixteen-column
// panels from right to left with one thread per (row, four columns) and a serial inner loop over up
// to 112 terms, and took 47 k of the 122 k cycles of a diagonal block.
// One 4 x 4 tile of  T = L21 * X11  (MODE 0: T to tmp)  or  X21 = -X22 * T  (MODE 1: into s).
template <int MODE>
__device__ __forceinline__ void inv_tile(double* s, double* tmp, const int o, const int h, const int pair,
                                         const int r0, const int c0) {
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    double* T = tmp + pair * h * h;  // column-major h x h
    if (MODE == 0) {
        // T[r][c] = sum_{m >= c} L21[r][m] * X11[m][c];  L21[r][m] = s(o + h + r, o + m), X11[m][c] = s(o + m, o + c)
        for (int m = c0; m < h; ++m) {
            double a[4], x[4];
            const double* colm = s + (o + m) * PLD;
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = colm[o + h + r0 + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = (m >= c0 + q) ? s[(o + c0 + q) * PLD + o + m] : 0.0;
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a_][b] += a[a_] * x[b];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_) T[(c0 + b) * h + r0 + a_] = acc[a_][b];
    } else {
        // X21[r][c] = -sum_{m <= r} X22[r][m] * T[m][c];  X22[r][m] = s(o + h + r, o + h + m)
        for (int m = 0; m < r0 + 4; ++m) {
            double x[4], t[4];
            const double* colm = s + (o + h + m) * PLD + o + h + r0;
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = (m <= r0 + q) ? colm[q] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) t[q] = T[(c0 + q) * h + m];
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a_][b] += x[a_] * t[b];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_) s[(o + c0 + b) * PLD + o + h + r0 + a_] = -acc[a_][b];
    }
}

template <int NTH>
__device__ __forceinline__ void invert_lower_inplace(double* s, double* tmp, const double* rdiag, const int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    if (warp < DB / PW) {
        const int b0 = warp * PW;
        const int c = lane & (PW - 1);  // lanes 16..31 mirror 0..15 (no divergence), only 0..15 store
        double x[PW];
#pragma unroll
        for (int i = 0; i < PW; ++i) {
            double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) v -= s[(b0 + k) * PLD + b0 + i] * x[k];
            x[i] = (i < c) ? 0.0 : v * rdiag[b0 + i];
        }
        __syncwarp();
        if (lane < PW)
#pragma unroll
            for (int i = 0; i < PW; ++i)
                if (i >= c) s[(b0 + c) * PLD + b0 + i] = x[i];
    }
    __syncthreads();
    for (int h = PW; h < DB; h *= 2) {
        const int tiles_per_pair = (h / 4) * (h / 4), pairs = DB / (2 * h), total = tiles_per_pair * pairs;
        // tiles are dealt so that consecutive threads take consecutive rows of one tile column (shared
        // loads of a warp then touch consecutive addresses); heavy and light tiles alternate over the passes
        for (int it = tid; it < total; it += NTH) {
            const int pair = it / tiles_per_pair, w = it - pair * tiles_per_pair;
            inv_tile<0>(s, tmp, pair * 2 * h, h, pair, 4 * (w % (h / 4)), 4 * (w / (h / 4)));
        }
        __syncthreads();
        for (int it = tid; it < total; it += NTH) {
            const int pair = it / tiles_per_pair, w = it - pair * tiles_per_pair;
            inv_tile<1>(s, tmp, pair * 2 * h, h, pair, 4 * (w % (h / 4)), 4 * (w / (h / 4)));
        }
        __syncthreads();
    }
}

#ifdef ACES_DIAG_PROFILE
__device__ int g_diag_prof_count = 0;
__device__ unsigned long long g_c2_cycles[2] = {0, 0};
#define DPROF_DECL long long _t[12]; int _ti = 0; long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define DPROF(slot) { if (threadIdx.x == 0) { long long _n = clock64(); _acc[slot] += _n - _last; _last = _n; } }
#define DPROF_START long long _last = clock64();
#define DPROF_WARP(slot, t0) { if (threadIdx.x == 0) _acc[slot] += clock64() - (t0); }
#else
#define DPROF_DECL
#define DPROF(slot)
#define DPROF_START
#define DPROF_WARP(slot, t0)
#endif
// NTH threads (512: the stand-alone kernel, 256: inside the persistent factorisation kernel).  Global
// loads bypass L1 (__ldcg): in the persistent kernel another SM may have written the block since this
// SM last read it.
template <int NTH>
__device__ __forceinline__ void diag_block(const DiagDesc& d, const double rel_tol, double* sh) {
    DPROF_DECL
    DPROF_START
    constexpr int NWARP = NTH / 32, JB = NTH / DB, C2W = NWARP - 1;
    double* s = sh;                  // s[j*PLD + i] = A(i, j)
    double* tmp = sh + DB * PLD;     // scratch of the inverse
    double* rdiag = tmp + INV_TMP;   // reciprocals of the diagonal of L
    double* pfloor = rdiag + DB;     // smallest acceptable pivot per column
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // loads issued in batches of 8 before their shared-memory stores
        const int i = tid & (DB - 1), jb = tid >> 7;  // JB columns in flight per pass
#pragma unroll
        for (int j0 = 0; j0 < DB; j0 += 8 * JB) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __ldcg(d.A + i + (int64_t)(j0 + JB * q + jb) * d.lda);
#pragma unroll
            for (int q = 0; q < 8; ++q) s[(j0 + JB * q + jb) * PLD + i] = v[q];
        }
    }
    if (tid < DB) pfloor[tid] = d.diag0 ? rel_tol * fabs(__ldcg(d.diag0 + tid)) : 0.0;
    __syncthreads();
    DPROF(0)  // load
    // Right-looking factorisation with look-ahead inside the CTA: after the panel rows of panel p
    // are solved, all warps update the 16 columns of panel p+1 (c1); then warp 0 factorises the
    // next 16 x 16 diagonal piece while warps 1..15 update the remaining trailing columns (c2).
    int bad = 0;
    const double repl = d.dead_mode ? DEAD_PIVOT : 1.0;
    if (warp == 0) bad = diag16(s, rdiag, pfloor, 0, lane, repl);
    for (int k0 = 0; k0 < DB - PW; k0 += PW) {
        __syncthreads();
        DPROF(1)  // diag16 (warp 0) / c2 (other warps), whichever is longer
        const int r0 = k0 + PW;  // first trailing row / column
        // (b) panel rows: x * L_dd' = a, one thread per row
        if (tid < DB - r0) {
            const int i = r0 + tid;
            double a[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) a[c] = s[(k0 + c) * PLD + i];
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                double v = a[c];
#pragma unroll
                for (int k = 0; k < c; ++k) v -= a[k] * s[(k0 + k) * PLD + k0 + c];
                a[c] = v * rdiag[k0 + c];
            }
#pragma unroll
            for (int c = 0; c < PW; ++c) s[(k0 + c) * PLD + i] = a[c];
        }
        __syncthreads();
        DPROF(2)  // panel rows
        // (c1) the columns of the next panel: thread t owns column r0 + (t & 15), rows r0 + (t >> 4) + RS a
        {
            constexpr int RS = NTH / PW, NA = DB / RS;
            const int j = r0 + (tid & (PW - 1)), ib = r0 + (tid >> 4);
            double acc[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) acc[a] = 0.0;
#pragma unroll 4
            for (int k = 0; k < PW; ++k) {
                const double* col = s + (k0 + k) * PLD;
                const double cj = col[j];
#pragma unroll
                for (int a = 0; a < NA; ++a)
                    if (ib + RS * a < DB) acc[a] += col[ib + RS * a] * cj;
            }
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const int i = ib + RS * a;
                if (i < DB && i >= j) s[j * PLD + i] -= acc[a];
            }
        }
        __syncthreads();
        DPROF(3)  // c1
#ifdef ACES_DIAG_PROFILE
        const long long _tw = clock64();
#endif
        if (warp == 0) {
            bad += diag16(s, rdiag, pfloor, r0, lane, repl);
            DPROF_WARP(7, _tw)  // diag16 alone
        } else if (r0 + PW < DB) {
            // (c2) the columns behind the next panel: warp w owns columns j_b = c0 + C2W b (b = 0, 1, ...
            // while j_b < 128); a group of 32 rows [32 a, 32 a + 31] only needs the columns up to its
            // last row (lower triangle), a prefix of this warp's columns: the k loop runs once per row
            // group with exactly that many accumulators.
            const int c0 = r0 + PW + (warp - 1);
#ifdef ACES_DIAG_PROFILE
            const long long _tc = clock64();
#endif
            for (int a = (r0 + PW) >> 5; a < 4; ++a) {
                const int last = 32 * a + 31;
                if (last < c0) continue;
                const int nb = (last - c0) / C2W + 1;  // <= 8 (15 update warps) or 16 (7)
                const int row = lane + 32 * a;
                switch (nb) {
                    case 1: c2_rows<1, C2W>(s, k0, c0, row); break;
                    case 2: c2_rows<2, C2W>(s, k0, c0, row); break;
                    case 3: c2_rows<3, C2W>(s, k0, c0, row); break;
                    case 4: c2_rows<4, C2W>(s, k0, c0, row); break;
                    case 5: c2_rows<5, C2W>(s, k0, c0, row); break;
                    case 6: c2_rows<6, C2W>(s, k0, c0, row); break;
                    case 7: c2_rows<7, C2W>(s, k0, c0, row); break;
                    case 8: c2_rows<8, C2W>(s, k0, c0, row); break;
                    default:
                        if constexpr (C2W < 15) {  // fewer update warps: up to 16 columns each, in two halves
                            c2_rows<8, C2W>(s, k0, c0, row);
                            switch (nb - 8) {
                                case 1: c2_rows<1, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 2: c2_rows<2, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 3: c2_rows<3, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 4: c2_rows<4, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 5: c2_rows<5, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 6: c2_rows<6, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                case 7: c2_rows<7, C2W>(s, k0, c0 + 8 * C2W, row); break;
                                default: c2_rows<8, C2W>(s, k0, c0 + 8 * C2W, row); break;
                            }
                        } else {
                            c2_rows<8, C2W>(s, k0, c0, row);
                        }
                        break;
                }
            }
#ifdef ACES_DIAG_PROFILE
            if (blockIdx.x == 0 && (tid == 32 || tid == 15 * 32)) atomicAdd(&g_c2_cycles[tid == 32 ? 0 : 1], (unsigned long long)(clock64() - _tc));
#endif
        }
    }
    if (warp == 0 && bad && lane == 0 && d.flag) {
        if (d.dead_mode) atomicAdd(d.flag, bad);  // number of eliminated columns
        else atomicCAS(d.flag, 0, d.code);
    }
    __syncthreads();
    DPROF(1)
    // L -> global (strict upper triangle zeroed)
    for (int e = tid; e < DB * DB; e += NTH) {
        const int i = e & (DB - 1), j = e >> 7;
        d.A[i + (int64_t)j * d.lda] = (i >= j) ? s[j * PLD + i] : 0.0;
    }
    __syncthreads();
    DPROF(4)  // store L
    // ---- in-place inverse ------------------------------------------------------------------
    invert_lower_inplace<NTH>(s, tmp, rdiag, tid);
    DPROF(5)  // inverse
    for (int e = tid; e < DB * DB; e += NTH) {
        const int i = e & (DB - 1), j = e >> 7;
        d.dinv[i + j * DB] = (i >= j) ? s[j * PLD + i] : 0.0;
    }
    DPROF(6)  // store inverse
#ifdef ACES_DIAG_PROFILE
    if (threadIdx.x == 0 && blockIdx.x == 0 && atomicAdd(&g_diag_prof_count, 1) % 40 == 5)
        printf("[diag prof] cycles: load %lld | diag16-or-c2 wait %lld | panel rows %lld | c1 %lld | store L %lld | inverse %lld | "
               "store inv %lld | diag16 alone (7 calls) %lld | c2 of warp 1 / warp 15, sum over all launches so far %llu / %llu (launch %d)\n",
               _acc[0], _acc[1], _acc[2], _acc[3], _acc[4], _acc[5], _acc[6], _acc[7], g_c2_cycles[0], g_c2_cycles[1], g_diag_prof_count);
#endif
}

__global__ void __launch_bounds__(DIAG_THREADS, 1) diag_kernel(const DiagDesc* __restrict__ descs,
                                                               const DiagDesc one, const double rel_tol) {
    extern __shared__ __align__(16) double sh[];
    const DiagDesc d = descs ? descs[blockIdx.x] : one;
    if (!d.A) return;
    diag_block<DIAG_THREADS>(d, rel_tol, sh);
}

// ---- Cholesky factorisation as ONE persistent kernel over a static task list ---------------------------
//
// The blocked right-looking factorisation is a DAG of tile tasks (128 x 128 tiles, nb = n / 128):
//   F(k)            L_kk = chol(A_kk), inv(L_kk)                         needs every update of tile (k, k)
//   S(i, k)         A_ik <- A_ik inv(L_kk)'                              needs F(k) and every update of (i, k)
//   U(i, j, k0, k1) A_ij -= sum_{k0 <= k < k1} L_ik L_jk'  (K = 128 (k1 - k0) <= 512)
//                                                                      needs S(i, k), S(j, k) and the earlier chunks of (i, j)
// The host orders the tasks once per matrix size by simulating a list schedule on `grid` workers
// (critical path first: smallest target column, then row) and uploads that order.  Every CTA of the
// kernel repeatedly takes the next task of the list (one atomic), waits for the task's dependencies
// on counters in global memory, runs it and publishes its completion.  Taking tasks strictly in list
// order makes the scheme deadlock free however many CTAs are resident: a CTA only ever waits for
// tasks that precede its own in a topological order, and those are held by running CTAs.
// What this replaces: ~300 kernel launches on three streams per factorisation whose critical chain
// (diagonal block -> panel tile -> update tile -> next diagonal block) paid a launch gap per link and
// whose trailing updates could not start a CTA while the chain's small kernels waited for an SM.
// The two tiles on the critical path of every step, S(k+1, k) and the last update of (k+1, k+1), are
// split into eight 16-row slabs taken by eight CTAs.
struct DagTask {
    uint8_t type;   // 0 F, 1 S, 2 U, 3 S slab, 4 U slab
    uint8_t slab;   // 16-row slab of the tile (types 3, 4)
    uint16_t i, j;  // target tile (for F: i = j = k; for S: j = k)
    uint16_t k0, k1;
    uint16_t pad;
};
static_assert(sizeof(DagTask) == 12, "DagTask layout");

struct DagParams {
    double* A;
    int64_t lda;
    double* dinv;
    const double* diag0;
    int* flag;
    double rel_tol;
    int dead_mode;
    int nb;
    const DagTask* tasks;  // [far | near | F]
    int n_bulk, n_list;    // far list length; far + near: the F tasks start there
    int n_near;            // CTAs of the near pool
    int* next;     // [0] far cursor, [1] arrivals (role), [2] chain cursor, [3] near cursor
    int* cnt;      // [nb * nb] 8 x (k-blocks applied to tile (i, j))
    int* sdone;    // [nb * nb] slabs of L_ik that are final (8 = all)
    int* fdone;    // [nb]
    unsigned long long* trace;  // [n_tasks][3] globaltimer at grab / dependencies met / done (NULL: off)
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// the counters only grow: wait until *p >= want (sibling slabs of one chunk raise cnt past 8 * k0)
__device__ __forceinline__ void spin_until(const int* p, int want) {
    while (ld_acquire_fwd(p) < want) __nanosleep(64);
}

constexpr size_t DAG_SMEM = DIAG_SMEM > GEMM_SMEM ? DIAG_SMEM : GEMM_SMEM;

__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Are the dependencies of a list task met?  All counters are read before any is tested (the loads are
// independent: one L2 round trip); the caller fences after a positive answer.
__device__ __forceinline__ bool deps_ready(const DagParams& P, const DagTask& tk) {
    const int nb = P.nb, i = tk.i, j = tk.j;
    if (tk.type == 1 || tk.type == 3) {
        const int f = ld_relaxed(P.fdone + j), c = ld_relaxed(P.cnt + i * nb + j);
        return f >= 1 && c >= 8 * j;
    }
    const int k0 = tk.k0, nk = tk.k1 - tk.k0;
    int v[9];
    v[0] = ld_relaxed(P.cnt + i * nb + j) - 8 * k0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[1 + q] = q < nk ? ld_relaxed(P.sdone + i * nb + k0 + q) - 8 : 0;
        v[5 + q] = (q < nk && i != j) ? ld_relaxed(P.sdone + j * nb + k0 + q) - 8 : 0;
    }
    bool ok = true;
#pragma unroll
    for (int q = 0; q < 9; ++q) ok = ok && v[q] >= 0;
    return ok;
}

// Task lists (one device array): [far | near | F].  Roles by order of arrival:
//   chain CTA (first): F(0), F(1), ... in order;
//   the next P.n_near CTAs: the NEAR list -- everything the chain waits for within a few steps: the 16-row
//     slabs of S(k+1, k), S(k+2, k) and of the recent updates of tiles (j, j), (j+1, j), the older update
//     chunks of those tiles, and S(i, k) for the rows just below;
//   every other CTA: the FAR list (the bulk of the trailing updates).
// Both lists are taken in order with one atomicAdd per task; a CTA waits for the dependencies of the task
// it took.  Two pools instead of one list because the far updates are throughput bound and queue up:
// in a single queue the few tasks the chain is waiting for wait behind them.
// Every list is a subsequence of one topological order (the simulated start order), every list is served
// by resident CTAs that take its tasks in order and wait only for earlier tasks: no deadlock.
__global__ void __launch_bounds__(GT, 1) potrf_dag_kernel(const DagParams P) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_task, s_role;
    const int tid = threadIdx.x, nb = P.nb;
    if (tid == 0) s_role = atomicAdd(P.next + 1, 1);
    __syncthreads();
    const int role = s_role;  // 0 chain, 1 .. n_near: near pool, above: far pool
    const bool chain = role == 0, near = role >= 1 && role <= P.n_near;
    for (;;) {
        __syncthreads();  // the previous task is done with shared memory and s_task
        unsigned long long t_grab = 0;
        if (tid == 0) {
            if (P.trace) t_grab = global_ns();
            int got = -1;
            if (chain) {
                const int k = atomicAdd(P.next + 2, 1);
                if (k < nb) got = P.n_list + k;
            } else if (near) {
                const int b = atomicAdd(P.next + 3, 1);
                if (b < P.n_list - P.n_bulk) got = P.n_bulk + b;
            } else {
                const int b = atomicAdd(P.next, 1);
                if (b < P.n_bulk) got = b;
            }
            if (got >= 0) {
                const DagTask tk = P.tasks[got];
                if (tk.type == 0) spin_until(P.cnt + tk.j * nb + tk.j, 8 * tk.j);
                else
                    while (!deps_ready(P, tk)) __nanosleep(32);
                __threadfence();  // acquire: the counters were read with relaxed loads
            }
            s_task = got;
        }
        __syncthreads();
        const int t = s_task;
        if (t < 0) break;
        const DagTask tk = P.tasks[t];
        const int i = tk.i, j = tk.j, k0 = tk.k0, k1 = tk.k1;
        if (P.trace && tid == 0) {
            P.trace[3 * (size_t)t] = t_grab;
            P.trace[3 * (size_t)t + 1] = global_ns();
        }
        // ---- run ----
        double* Aij = P.A + (int64_t)i * DB + (int64_t)j * DB * P.lda;
        if (tk.type == 0) {
            DiagDesc d{Aij, P.lda, P.dinv + (int64_t)j * DB * DB, P.diag0 ? P.diag0 + (int64_t)j * DB : nullptr, P.flag,
                       j + 1, P.dead_mode};
            diag_block<GT>(d, P.rel_tol, sm);
        } else if (tk.type == 1 || tk.type == 3) {
            GemmP g{Aij, P.dinv + (int64_t)j * DB * DB, Aij, P.lda, DB, P.lda, DB, 1.0, 0.0};
            if (tk.type == 1) gemm_tile<false, false, 128, true>(g, 0, 0, sm);
            else gemm_tile<false, false, 16, true>(g, tk.slab, 0, sm);
        } else {
            const double* Li = P.A + (int64_t)i * DB + (int64_t)k0 * DB * P.lda;
            const double* Lj = P.A + (int64_t)j * DB + (int64_t)k0 * DB * P.lda;
            GemmP g{Li, Lj, Aij, P.lda, P.lda, P.lda, (k1 - k0) * DB, -1.0, 1.0};
            if (tk.type == 2) gemm_tile<false, false, 128, true>(g, 0, 0, sm);
            else gemm_tile<false, false, 16, true>(g, tk.slab, 0, sm);
        }
        // ---- publish ----
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            if (tk.type == 0) st_release_fwd(P.fdone + j, 1);
            else if (tk.type == 1) atomicAdd(P.sdone + i * nb + j, 8);
            else if (tk.type == 3) atomicAdd(P.sdone + i * nb + j, 1);
            else if (tk.type == 2) atomicAdd(P.cnt + i * nb + j, 8 * (k1 - k0));
            else atomicAdd(P.cnt + i * nb + j, k1 - k0);
            if (P.trace) P.trace[3 * (size_t)t + 2] = global_ns();
        }
    }
}

// ---- triangular solves with one right-hand side -------------------------------------------------
constexpr int TRSV_THREADS = 256;

// out[c] = sum_r T[r + c*ld] * v[r] for one 128 x 128 tile; warp w handles columns w, w+8, ...
// All 64 loads of a lane are issued before the reductions (the tile comes from L2 / HBM).
__device__ __forceinline__ void tile_t_times_vec(const double* __restrict__ T, int64_t ld,
                                                 const double* v, double* out, int lane, int warp) {
    double s[16];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
        const double* col = T + (int64_t)(warp + 8 * cc) * ld;
        s[cc] = col[lane] * v[lane] + col[lane + 32] * v[lane + 32] + col[lane + 64] * v[lane + 64] +
                col[lane + 96] * v[lane + 96];
    }
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
        double t = s[cc];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) out[warp + 8 * cc] = t;
    }
}


// ---- persistent triangular solves --------------------------------------------------------------
// One launch per solve.  CTA c owns the 128-row blocks c, c + G, ... and runs them in dependency
// order; a block publishes its solution block and raises a flag, the blocks that depend on it
// spin on that flag (all CTAs are co-resident: cooperative launch, grid <= SM count) and fold the
// corresponding 128 x 128 tile of L into their right-hand side.  The chain of 128-blocks is
// inherently sequential; what this removes is one launch (and its drain / fill) per block.
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// out[i] = sum_j T[i + j*ld] * v[j] for a 128 x 128 tile, 256 threads (two column halves)
__device__ __forceinline__ void tile_times_vec(const double* __restrict__ T, int64_t ld, const double* v,
                                               double* part, int tid) {
    const int i = tid & (DB - 1), h = tid >> 7;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double* Tp = T + i + (int64_t)(h * 64) * ld;
    const double* vp = v + h * 64;
#pragma unroll
    for (int j0 = 0; j0 < 64; j0 += 16) {
        double t[16];  // 16 independent loads in flight per thread (the tile streams from L2 / HBM)
#pragma unroll
        for (int q = 0; q < 16; ++q) t[q] = Tp[(int64_t)(j0 + q) * ld];
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            s0 += t[q] * vp[j0 + q];
            s1 += t[q + 1] * vp[j0 + q + 1];
            s2 += t[q + 2] * vp[j0 + q + 2];
            s3 += t[q + 3] * vp[j0 + q + 3];
        }
    }
    part[tid] = (s0 + s1) + (s2 + s3);
}

// The last dependency of a block is on the critical path of the whole solve (block r can only
// finish after y_{r-1} arrives), so its tile of L is preloaded into registers and the block
// inverse into shared memory BEFORE the wait: once the flag is up only one 1 KB vector load and
// on-chip arithmetic remain.
constexpr size_t TRSV_SMEM = (size_t)DB * DB * sizeof(double);

__global__ void __launch_bounds__(TRSV_THREADS, 1) trsv_fwd_persistent(const double* __restrict__ L, int64_t ldl,
                                                                       const double* __restrict__ dinv,
                                                                       const double* __restrict__ b, double* y,
                                                                       int nb, int* flags) {
    extern __shared__ __align__(16) double sX[];  // inv(L_rr), column-major 128 x 128
    __shared__ double acc[DB], vk[DB], part[TRSV_THREADS];
    const int tid = threadIdx.x, i = tid & (DB - 1), h = tid >> 7;
    for (int r = blockIdx.x; r < nb; r += gridDim.x) {
        __syncthreads();  // the previous block of this CTA is done with sX
        if (tid < DB) acc[tid] = b[(int64_t)r * DB + tid];
        {
            const double* X = dinv + (int64_t)r * DB * DB;
            for (int e = tid; e < DB * DB; e += TRSV_THREADS) sX[e] = X[e];
        }
        __syncthreads();
        for (int k = 0; k + 1 < r; ++k) {
            if (tid == 0)
                while (ld_acquire(flags + k) == 0) {}
            __syncthreads();
            if (tid < DB) vk[tid] = __ldcg(y + (int64_t)k * DB + tid);
            __syncthreads();
            tile_times_vec(L + (int64_t)r * DB + (int64_t)k * DB * ldl, ldl, vk, part, tid);
            __syncthreads();
            if (tid < DB) acc[tid] -= part[tid] + part[tid + DB];
            __syncthreads();
        }
        if (r > 0) {
            const int k = r - 1;
            double t[64];  // this thread's half row of L[r, r-1]
            const double* Tp = L + (int64_t)r * DB + i + ((int64_t)k * DB + h * 64) * ldl;
#pragma unroll
            for (int q = 0; q < 64; ++q) t[q] = Tp[(int64_t)q * ldl];
            if (tid == 0)
                while (ld_acquire(flags + k) == 0) {}
            __syncthreads();
            if (tid < DB) vk[tid] = __ldcg(y + (int64_t)k * DB + tid);
            __syncthreads();
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int q = 0; q < 64; q += 4) {
                s0 += t[q] * vk[h * 64 + q];
                s1 += t[q + 1] * vk[h * 64 + q + 1];
                s2 += t[q + 2] * vk[h * 64 + q + 2];
                s3 += t[q + 3] * vk[h * 64 + q + 3];
            }
            part[tid] = (s0 + s1) + (s2 + s3);
            __syncthreads();
            if (tid < DB) acc[tid] -= part[tid] + part[tid + DB];
            __syncthreads();
        }
        tile_times_vec(sX, DB, acc, part, tid);  // the block inverse is lower triangular (zeros above)
        __syncthreads();
        if (tid < DB) y[(int64_t)r * DB + tid] = part[tid] + part[tid + DB];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(flags + r, 1);
    }
}

__global__ void __launch_bounds__(TRSV_THREADS, 1) trsv_bwd_persistent(const double* __restrict__ L, int64_t ldl,
                                                                       const double* __restrict__ dinv,
                                                                       const double* __restrict__ yin, double* x,
                                                                       int nb, int* flags) {
    extern __shared__ __align__(16) double sX[];
    __shared__ double acc[DB], vk[DB], out[DB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = blockIdx.x; q < nb; q += gridDim.x) {
        const int k = nb - 1 - q;  // blocks in descending order
        __syncthreads();
        if (tid < DB) acc[tid] = yin[(int64_t)k * DB + tid];
        {
            const double* X = dinv + (int64_t)k * DB * DB;
            for (int e = tid; e < DB * DB; e += TRSV_THREADS) sX[e] = X[e];
        }
        __syncthreads();
        for (int r = nb - 1; r > k + 1; --r) {
            if (tid == 0)
                while (ld_acquire(flags + r) == 0) {}
            __syncthreads();
            if (tid < DB) vk[tid] = __ldcg(x + (int64_t)r * DB + tid);
            __syncthreads();
            tile_t_times_vec(L + (int64_t)r * DB + (int64_t)k * DB * ldl, ldl, vk, out, lane, warp);
            __syncthreads();
            if (tid < DB) acc[tid] -= out[tid];
            __syncthreads();
        }
        if (k + 1 < nb) {
            const int r = k + 1;
            double t[16][4];  // columns warp, warp + 8, ... of L[k+1, k], four row chunks per lane
            const double* T = L + (int64_t)r * DB + (int64_t)k * DB * ldl;
#pragma unroll
            for (int cc = 0; cc < 16; ++cc)
#pragma unroll
                for (int m = 0; m < 4; ++m) t[cc][m] = T[lane + 32 * m + (int64_t)(warp + 8 * cc) * ldl];
            if (tid == 0)
                while (ld_acquire(flags + r) == 0) {}
            __syncthreads();
            if (tid < DB) vk[tid] = __ldcg(x + (int64_t)r * DB + tid);
            __syncthreads();
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) {
                double sum = (t[cc][0] * vk[lane] + t[cc][1] * vk[lane + 32]) + (t[cc][2] * vk[lane + 64] + t[cc][3] * vk[lane + 96]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) out[warp + 8 * cc] = sum;
            }
            __syncthreads();
            if (tid < DB) acc[tid] -= out[tid];
            __syncthreads();
        }
        tile_t_times_vec(sX, DB, acc, out, lane, warp);
        __syncthreads();
        if (tid < DB) x[(int64_t)k * DB + tid] = out[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(flags + k, 1);
    }
}

__global__ void set_identity_kernel(int64_t n, double* A, int64_t lda) {
    const int64_t total = n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % n, j = e / n;
        A[i + j * lda] = (i == j) ? 1.0 : 0.0;
    }
}
__global__ void get_diag_kernel(int64_t n, const double* A, int64_t lda, double* d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = A[i + i * lda];
}
__global__ void fill_diag_kernel(double* A, int64_t lda, int64_t from, int64_t to, double v) {
    for (int64_t i = from + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < to;
         i += (int64_t)gridDim.x * blockDim.x)
        A[i + i * lda] = v;
}
__global__ void transpose_kernel(int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + threadIdx.x, j = j0 + r;
        tile[r][threadIdx.x] = (i < m && j < n) ? A[i + j * lda] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t j = j0 + threadIdx.x, i = i0 + r;
        if (i < m && j < n) B[j + i * ldb] = tile[threadIdx.x][r];
    }
}
// A(i, j) = A(j, i) for i < j, tile-wise through shared memory so that both sides are coalesced.
__global__ void symmetrize_kernel(int64_t n, double* A, int64_t lda) {
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.x, bj = blockIdx.y;
    if (bj > bi) return;  // (bi, bj) is a lower tile; it is mirrored into (bj, bi)
    const int64_t i0 = bi * 32, j0 = bj * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + threadIdx.x, j = j0 + r;
        tile[r][threadIdx.x] = (i < n && j < n) ? A[i + j * lda] : 0.0;  // tile[r][c] = A(i0+c, j0+r)
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        // write A(j0 + c, i0 + r) = A(i0 + r, j0 + c) = tile[c][r]
        const int64_t row = j0 + threadIdx.x, col = i0 + r;
        if (row < n && col < n && row < col) A[row + col * lda] = tile[threadIdx.x][r];
    }
}

template <typename K>
int set_smem(aces_ctx* ctx, K kern, size_t bytes) {
    ACES_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return ACES_OK;
}

// Function attributes are per device: one flag per device, set under a lock (several contexts, one
// host thread each, may make their first call at the same time).
int configure(aces_ctx* ctx) {
    static std::mutex mu;
    static bool done_dev[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    bool& done = done_dev[ctx->device & 63];
    if (done) return ACES_OK;
    ACES_TRY(set_smem(ctx, gemm_kernel<false, false, 128>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, true, 128>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, gemm_kernel<true, true, 128>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, false, 16>, gemm_smem(16)));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, true, 16>, gemm_smem(16)));
    ACES_TRY(set_smem(ctx, gemm_kernel<true, true, 16>, gemm_smem(16)));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, false, 32>, gemm_smem(32)));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, true, 32>, gemm_smem(32)));
    ACES_TRY(set_smem(ctx, gemm_kernel<true, true, 32>, gemm_smem(32)));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, false, 64>, gemm_smem(64)));
    ACES_TRY(set_smem(ctx, gemm_kernel<false, true, 64>, gemm_smem(64)));
    ACES_TRY(set_smem(ctx, gemm_kernel<true, true, 64>, gemm_smem(64)));
    ACES_TRY(set_smem(ctx, gemm_batched_kernel<false, false>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, gemm_batched_kernel<false, true>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, gemm_batched_kernel<true, true>, GEMM_SMEM));
    ACES_TRY(set_smem(ctx, diag_kernel, DIAG_SMEM));
    ACES_TRY(set_smem(ctx, potrf_dag_kernel, DAG_SMEM));
    ACES_TRY(set_smem(ctx, trsv_fwd_persistent, TRSV_SMEM));
    ACES_TRY(set_smem(ctx, trsv_bwd_persistent, TRSV_SMEM));
    done = true;
    return ACES_OK;
}

template <bool A_KC, bool B_KC>
int launch_gemm(aces_ctx* ctx, const GemmP& p, int64_t m, int64_t n, int lower_only) {
    // 64-row tiles when 128-row tiles would leave most SMs idle or give a ragged second wave
    const int64_t tiles128 = lower_only ? ((m / BM) * (n / BN) + std::min(m / BM, n / BN)) / 2 : (m / BM) * (n / BN);
    static const int force64 = [] { const char* e = getenv("ACES_GEMM_64"); return e ? atoi(e) : 1; }();
    // force64: 1 = every launch whose K is at most 512 (the Cholesky updates), 2 = every launch
    const bool want64 = tiles128 < 2 * (int64_t)ctx->sm_count || force64 == 2 || (force64 == 1 && p.k <= 512);
    // Small launches are latency, not throughput: a single 128 x 128 tile on the Cholesky's critical
    // path runs as eight CTAs of 16 rows (a seventh of the time one CTA needs), and launches that
    // would fill less than half the SMs with 64-row tiles use 32-row tiles.
    // ACES_GEMM_32 (development knob): 0 off, 1 = 32-row tiles only, 2 = both, 3 = 16-row for single tiles only
    static const int tiny32 = [] { const char* e = getenv("ACES_GEMM_32"); return e ? atoi(e) : 2; }();
    // 32-row tiles when 64-row tiles would give fewer than (8 / fill32) CTAs per SM ... fill32 = 16: less than half the SMs
    static const int fill32 = [] { const char* e = getenv("ACES_GEMM_32_FILL"); return e ? atoi(e) : 16; }();
    const bool single = tiles128 <= 2;
    if (tiny32 && (const double*)p.C != p.B && (single || (tiny32 != 3 && fill32 * tiles128 < 4 * (int64_t)ctx->sm_count))) {
        if (single && tiny32 >= 2) {
            dim3 grid((unsigned)(m / 16), (unsigned)(n / BN));
            gemm_kernel<A_KC, B_KC, 16><<<grid, GT, gemm_smem(16), ctx->stream>>>(p, lower_only);
        } else {
            dim3 grid((unsigned)(m / 32), (unsigned)(n / BN));
            gemm_kernel<A_KC, B_KC, 32><<<grid, GT, gemm_smem(32), ctx->stream>>>(p, lower_only);
        }
    } else if (want64 && (const double*)p.C != p.B) {
        dim3 grid((unsigned)(m / 64), (unsigned)(n / BN));
        gemm_kernel<A_KC, B_KC, 64><<<grid, GT, gemm_smem(64), ctx->stream>>>(p, lower_only);
    } else {
        dim3 grid((unsigned)(m / BM), (unsigned)(n / BN));
        gemm_kernel<A_KC, B_KC, 128><<<grid, GT, GEMM_SMEM, ctx->stream>>>(p, lower_only);
    }
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace

int gemm(aces_ctx* ctx, GemmForm form, int64_t m, int64_t n, int64_t k, double alpha,
         const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
         int64_t ldc, int lower_only) {
    if (m == 0 || n == 0) return ACES_OK;
    ACES_CHECK(ctx, m % BM == 0 && n % BN == 0 && k % BK == 0 && k > 0,
               "dense::gemm: dimensions %lld x %lld x %lld are not padded", (long long)m, (long long)n, (long long)k);
    ACES_CHECK(ctx, lda % 2 == 0 && ldb % 2 == 0, "dense::gemm: odd leading dimension");
    ACES_TRY(configure(ctx));
    GemmP p{A, B, C, lda, ldb, ldc, (int)k, alpha, beta};
    switch (form) {
        case GEMM_NT: return launch_gemm<false, false>(ctx, p, m, n, lower_only);
        case GEMM_NN: return launch_gemm<false, true>(ctx, p, m, n, lower_only);
        case GEMM_TN: return launch_gemm<true, true>(ctx, p, m, n, lower_only);
    }
    return aces_fail(ctx, ACES_E_INVALID, "dense::gemm: unknown form");
}

int gemm_batched(aces_ctx* ctx, GemmForm form, const GemmDesc* descs, int count, int max_tiles_m,
                 int max_tiles_n) {
    if (count <= 0 || max_tiles_m <= 0 || max_tiles_n <= 0) return ACES_OK;
    ACES_TRY(configure(ctx));
    dim3 grid((unsigned)max_tiles_m, (unsigned)max_tiles_n, (unsigned)count);
    switch (form) {
        case GEMM_NT: gemm_batched_kernel<false, false><<<grid, GT, GEMM_SMEM, ctx->stream>>>(descs); break;
        case GEMM_NN: gemm_batched_kernel<false, true><<<grid, GT, GEMM_SMEM, ctx->stream>>>(descs); break;
        case GEMM_TN: gemm_batched_kernel<true, true><<<grid, GT, GEMM_SMEM, ctx->stream>>>(descs); break;
    }
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

int diag_blocks(aces_ctx* ctx, const DiagDesc* descs, int count, double rel_tol) {
    if (count <= 0) return ACES_OK;
    ACES_TRY(configure(ctx));
    DiagDesc none{};
    diag_kernel<<<(unsigned)count, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(descs, none, rel_tol);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

namespace {

// Static schedule of the factorisation DAG for an nb x nb tile matrix on `workers` CTAs: a list schedule
// simulated with estimated task durations (microseconds), priority = (target column, target row,
// first k-block) -- the tile that the next diagonal block is waiting for goes first.  The result is
// the order in which the simulated workers START the tasks, which is a topological order of the DAG.
// Returned: [far list | near list | F(0) .. F(nb-1)]; *n_far = length of the far list.  Both lists keep the
// simulated start order.  `pad` = 1 marks a near task.
std::vector<DagTask> build_dag_schedule(int nb, int workers, int* n_far = nullptr) {
    struct Node {
        DagTask t;
        double cost;
        int indeg = 0;
        std::vector<int> succ;
    };
    std::vector<Node> nodes;
    nodes.reserve((size_t)nb * nb * (nb / 24 + 2));
    auto add = [&](uint8_t type, int slab, int i, int j, int k0, int k1, double cost, int urgent = 0) {
        Node nd;
        nd.t = DagTask{type, (uint8_t)slab, (uint16_t)i, (uint16_t)j, (uint16_t)k0, (uint16_t)k1, (uint16_t)urgent};
        nd.cost = cost;
        nodes.push_back(nd);
        return (int)nodes.size() - 1;
    };
    auto edge = [&](int from, int to) {
        nodes[(size_t)from].succ.push_back(to);
        nodes[(size_t)to].indeg += 1;
    };
    // per tile: the task group (1 task, or 8 slabs) of its last update chunk and of its S / F task
    std::vector<std::vector<int>> last_upd((size_t)nb * nb), solved((size_t)nb * nb);
    std::vector<int> fac((size_t)nb, -1);
    static const int NEAR_U = [] { const char* e = getenv("ACES_POTRF_NEAR_U"); return e ? atoi(e) : 3; }();  // recent chunks of tiles with i - j <= NEAR_U are near
    static const int NEAR_S = [] { const char* e = getenv("ACES_POTRF_NEAR_S"); return e ? atoi(e) : 4; }();  // S(i, j) with i - j <= NEAR_S is near
    static const int CH = [] { const char* e = getenv("ACES_POTRF_CHUNK"); return e ? std::min(4, std::max(1, atoi(e))) : 4; }();  // k-blocks per update chunk (K = 512)
    for (int j = 0; j < nb; ++j) {
        for (int i = j; i < nb; ++i) {
            // Update chunks of tile (i, j), in order.  How soon is the tile needed after its last inputs
            // (column j - 1) exist?  Rows j and j + 1 at once (the next diagonal block and the tile under
            // it): their recent k-blocks are applied as eight 16-row slabs each, the very last one alone.
            // Rows up to j + NEAR within a few steps: full-tile tasks, but urgent.  Everything that only
            // involves k-blocks older than a chunk has several steps of slack: bulk.
            const bool critical = i - j <= 1, close = i - j <= NEAR_U;
            const int recent = close ? std::max(0, (j - 1 - CH) / CH * CH) : j;  // first k-block handled by the near pool
            std::vector<int> prev;
            for (int k0 = 0; k0 < j;) {
                int k1 = std::min(k0 - k0 % CH + CH, j);
                if (critical && k1 == j && k0 < j - 1) k1 = j - 1;  // the last k-block of a critical tile alone
                const bool urgent = k0 >= recent, split = critical && urgent;
                std::vector<int> grp;
                if (split)
                    for (int sl = 0; sl < 8; ++sl)
                        grp.push_back(add(4, sl, i, j, k0, k1, 4.0 + 3.5 * (k1 - k0), 1));
                else
                    grp.push_back(add(2, 0, i, j, k0, k1, 12.0 + 19.0 * (k1 - k0), (critical || urgent) ? 1 : 0));
                for (int g : grp) {
                    for (int p : prev) edge(p, g);
                    for (int k = k0; k < k1; ++k) {
                        for (int sv : solved[(size_t)i * nb + k]) edge(sv, g);
                        if (i != j)
                            for (int sv : solved[(size_t)j * nb + k]) edge(sv, g);
                    }
                }
                prev = grp;
                k0 = k1;
            }
            last_upd[(size_t)i * nb + j] = prev;
            if (i == j) {
                fac[(size_t)j] = add(0, 0, j, j, 0, 0, 52.0);
                for (int p : prev) edge(p, fac[(size_t)j]);
            } else {
                std::vector<int> grp;
                if (i - j <= 2)
                    for (int sl = 0; sl < 8; ++sl) grp.push_back(add(3, sl, i, j, 0, 0, 6.0, 1));
                else
                    grp.push_back(add(1, 0, i, j, 0, 0, 25.0, i - j <= NEAR_S ? 1 : 0));
                for (int g : grp) {
                    edge(fac[(size_t)j], g);
                    for (int p : prev) edge(p, g);
                }
                solved[(size_t)i * nb + j] = grp;
            }
        }
    }
    // list scheduling
    // Priority = bottom level (the longest path of estimated durations from the task to the end of the
    // factorisation, critical-path list scheduling): the tasks that feed the chain of diagonal blocks soon
    // carry the whole remaining chain behind them and go first, however early or late their tile's column
    // is; the far trailing updates, which nothing waits for, fill the gaps.  Every edge goes from a task
    // created earlier to one created later, so one backward sweep computes the levels.
    std::vector<double> level(nodes.size(), 0.0);
    for (int id = (int)nodes.size() - 1; id >= 0; --id) {
        double below = 0.0;
        for (int sc : nodes[(size_t)id].succ) below = std::max(below, level[(size_t)sc]);
        level[(size_t)id] = nodes[(size_t)id].cost + below;
    }
    auto later = [&](int a, int b) {
        if (level[(size_t)a] != level[(size_t)b]) return level[(size_t)a] < level[(size_t)b];
        return a > b;
    };
    std::priority_queue<int, std::vector<int>, decltype(later)> ready(later);
    using Ev = std::pair<double, int>;
    std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> events;
    // (the only task without a dependency is F(0): it goes to the chain worker)
    std::vector<DagTask> order;
    order.reserve(nodes.size());
    double now = 0.0;
    int free_workers = std::max(workers - 1, 1);  // (one pool in the simulation)
    bool chain_free = true;
    std::vector<int> f_ready;  // F tasks whose updates are complete (at most one at a time)
    for (int id = 0; id < (int)nodes.size(); ++id)
        if (nodes[(size_t)id].indeg == 0) {
            if (nodes[(size_t)id].t.type == 0) f_ready.push_back(id);
            else ready.push(id);
        }
    size_t started = 0;
    while (started < nodes.size()) {
        if (chain_free && !f_ready.empty()) {
            const int id = f_ready.back();
            f_ready.pop_back();
            events.push({now + nodes[(size_t)id].cost, id});
            chain_free = false;
            ++started;
        }
        while (free_workers > 0 && !ready.empty()) {
            const int id = ready.top();
            ready.pop();
            order.push_back(nodes[(size_t)id].t);
            events.push({now + nodes[(size_t)id].cost, id});
            --free_workers;
            ++started;
        }
        if (events.empty()) break;  // cannot happen for a DAG
        const Ev ev = events.top();
        events.pop();
        now = ev.first;
        if (nodes[(size_t)ev.second].t.type == 0) chain_free = true;
        else ++free_workers;
        for (int sc : nodes[(size_t)ev.second].succ)
            if (--nodes[(size_t)sc].indeg == 0) {
                if (nodes[(size_t)sc].t.type == 0) f_ready.push_back(sc);
                else ready.push(sc);
            }
    }
    // split the start order into the bulk and the urgent list
    std::vector<DagTask> out;
    out.reserve(order.size() + (size_t)nb);
    for (const DagTask& t : order)
        if (!t.pad) out.push_back(t);
    if (n_far) *n_far = (int)out.size();
    for (const DagTask& t : order)
        if (t.pad) out.push_back(t);
    for (int k = 0; k < nb; ++k) out.push_back(nodes[(size_t)fac[(size_t)k]].t);
    order.swap(out);
    return order;
}

int potrf_dag(aces_ctx* ctx, int64_t n, double* A, int64_t lda, double* dinv, int* flag, const double* diag0,
              double rel_tol, int dead_mode) {
    const int nb = (int)(n / DB);
    ACES_CHECK(ctx, nb < 65536, "dense::potrf: matrix too large for the task encoding");
    static const int NEAR_CTAS = [] { const char* e = getenv("ACES_POTRF_NEAR_CTAS"); return e ? std::max(1, atoi(e)) : 26; }();
    const int grid = std::max(ctx->sm_count, NEAR_CTAS + 2);  // the chain CTA, the near pool and at least one far worker
    if (ctx->dag_nb != nb || ctx->dag_grid != grid) {
        int n_bulk = 0;
        const std::vector<DagTask> order = build_dag_schedule(nb, grid, &n_bulk);
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_dag_tasks, sizeof(DagTask) * order.size()));
        ACES_CUDA(ctx, cudaMemcpyAsync(ctx->d_dag_tasks.ptr, order.data(), sizeof(DagTask) * order.size(), cudaMemcpyHostToDevice,
                                       ctx->stream));
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `order` dies at the end of this block
        ctx->dag_nb = nb;
        ctx->dag_grid = grid;
        ctx->dag_ntasks = n_bulk;
        ctx->dag_nlist = (int)order.size() - nb;  // far + near lists; the nb F tasks follow
    }
    const size_t n_state = 16 + 2 * (size_t)nb * nb + (size_t)nb;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_dag_state, sizeof(int) * n_state));
    int* st = (int*)ctx->d_dag_state.ptr;
    ACES_CUDA(ctx, cudaMemsetAsync(st, 0, sizeof(int) * n_state, ctx->stream));
    DagParams P;
    P.A = A; P.lda = lda; P.dinv = dinv; P.diag0 = diag0; P.flag = flag; P.rel_tol = rel_tol; P.dead_mode = dead_mode;
    P.nb = nb;
    P.tasks = (const DagTask*)ctx->d_dag_tasks.ptr;
    P.n_bulk = ctx->dag_ntasks;
    P.n_list = ctx->dag_nlist;
    P.n_near = NEAR_CTAS;
    P.next = st;
    P.cnt = st + 16;
    P.sdone = P.cnt + (size_t)nb * nb;
    P.fdone = P.sdone + (size_t)nb * nb;
    P.trace = nullptr;
    static const bool trace_on = getenv("ACES_POTRF_TRACE") != nullptr;  // development: per-task timestamps
    if (trace_on) {
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_dag_trace, sizeof(unsigned long long) * 3 * (size_t)(ctx->dag_nlist + nb)));
        P.trace = (unsigned long long*)ctx->d_dag_trace.ptr;
    }
    potrf_dag_kernel<<<(unsigned)grid, GT, DAG_SMEM, ctx->stream>>>(P);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace

int potrf(aces_ctx* ctx, int64_t n, double* A, int64_t lda, double* dinv, int* flag, const double* diag0,
          double rel_tol, int dead_mode) {
    ACES_CHECK(ctx, n % DB == 0, "dense::potrf: n=%lld is not padded", (long long)n);
    ACES_TRY(configure(ctx));
    // ACES_POTRF_STREAMS=1 (development knob) selects the previous multi-stream, launch-per-tile version
    // Which version: the persistent task-DAG kernel wins where the chain of diagonal blocks matters
    // (rotated d = 9, N = 6784: 5.8 ms against 6.7 ms); for large matrices the factorisation is pure GEMM
    // throughput and the multi-stream version with its two-CTA-per-SM 64-row tiles is faster (d = 17,
    // N = 24832: 172 ms against 230 ms, profiles/fit_r02_potrf_variants.txt).  ACES_POTRF_STREAMS = 1 / 0
    // forces one or the other (development knob).
    static const int force = [] { const char* e = getenv("ACES_POTRF_STREAMS"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool use_streams = force >= 0 ? force == 1 : n / DB > 80;
    if (!use_streams) return potrf_dag(ctx, n, A, lda, dinv, flag, diag0, rel_tol, dead_mode);
    constexpr int64_t NB2 = 4 * DB;  // outer panel: the bulk of the flops runs as K = 512 updates
    // development probe: skip parts of the factorisation to time the others (results are then wrong)
#ifdef ACES_POTRF_PROBE  // never in a production build: a stray environment variable must not corrupt results
    const char* skip_env = getenv("ACES_POTRF_SKIP");
    const int skip = skip_env ? atoi(skip_env) : 0;  // 1 diag, 2 critical tiles, 4 second stream, 8 look-ahead, 16 side
#else
    constexpr int skip = 0;
#endif
    const cudaStream_t main_stream = ctx->stream;
    bool side_pending = false;
    // The chain diag(k) -> diag(k+1) is the critical path (one CTA at a time).  Only one tile of the
    // panel and one tile of the update stand between two diagonal blocks: block row k+1 of the panel
    // and the update of block (k+1, k+1).  They run on the main stream; the rest of step k (the
    // panel rows below, the update of the other blocks of the outer panel) runs on a second stream
    // while the next diagonal block is being factorised.
    bool aux2_pending = false;
    auto on_aux2 = [&](auto&& fn) -> int {
        ctx->stream = ctx->aux2_stream;
        const int rc = fn();
        ctx->stream = main_stream;
        return rc;
    };
    for (int64_t K0 = 0; K0 < n; K0 += NB2) {
        const int64_t kb = std::min(NB2, n - K0);
        for (int64_t k = K0; k < K0 + kb; k += DB) {
            double* Akk = A + k * (lda + 1);
            double* dk = dinv + (k / DB) * DB * DB;
            DiagDesc one{Akk, lda, dk, diag0 ? diag0 + k : nullptr, flag, (int32_t)(k / DB) + 1, dead_mode};
            if (!(skip & 1)) diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM, ctx->stream>>>(nullptr, one, rel_tol);
            ACES_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1;
            const int64_t rows = n - (k + DB);
            if (rows <= 0) continue;
            double* panel = Akk + DB;
            const int64_t cols = K0 + kb - (k + DB);  // columns of the outer panel behind block column k
            // block column k below the diagonal got its last update from the second stream (step k-1)
            if (aux2_pending) ACES_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_aux2, 0));
            // critical tile: panel rows [k+128, k+256) <- panel * inv(L_kk)'  (in place), then block (k+1, k+1)
            if (!(skip & 2)) {
            ACES_TRY(gemm(ctx, GEMM_NT, DB, DB, DB, 1.0, panel, lda, dk, DB, 0.0, panel, lda, 0));
            if (cols > 0)
                ACES_TRY(gemm(ctx, GEMM_NT, DB, DB, DB, -1.0, panel, lda, panel, lda, 1.0, Akk + DB * (lda + 1), lda, 1));
            }
            const int64_t rest = rows - DB;
            if (rest > 0 && !(skip & 4)) {
                ACES_CUDA(ctx, cudaEventRecord(ctx->ev_crit, main_stream));
                ACES_CUDA(ctx, cudaStreamWaitEvent(ctx->aux2_stream, ctx->ev_crit, 0));
                double* below = panel + DB;  // panel rows from k + 256 on
                ACES_TRY(on_aux2([&] {
                    ACES_TRY(gemm(ctx, GEMM_NT, rest, DB, DB, 1.0, below, lda, dk, DB, 0.0, below, lda, 0));
                    if (cols > 0)  // block column k+1 below its diagonal block
                        ACES_TRY(gemm(ctx, GEMM_NT, rest, DB, DB, -1.0, below, lda, panel, lda, 1.0,
                                      Akk + DB * (lda + 1) + DB, lda, 0));
                    if (cols > DB)  // the other columns of the outer panel (lower triangle from (k+2, k+2))
                        ACES_TRY(gemm(ctx, GEMM_NT, rest, cols - DB, DB, -1.0, below, lda, below, lda, 1.0,
                                      Akk + 2 * DB * (lda + 1), lda, 1));
                    return (int)ACES_OK;
                }));
                ACES_CUDA(ctx, cudaEventRecord(ctx->ev_aux2, ctx->aux2_stream));
                aux2_pending = true;
            }
        }
        // the outer panel is final once the second stream is done with it
        if (aux2_pending) ACES_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_aux2, 0));
        const int64_t rows = n - (K0 + kb);
        if (rows > 0) {
            // Trailing update with look-ahead: the columns of the NEXT panel are updated on the main
            // stream (the next panel's factorisation needs nothing else), the rest on the side
            // stream, where it overlaps the launch-latency-bound chain of diagonal-block kernels.
            const double* P = A + (K0 + kb) + K0 * lda;  // rows below the panel, its kb columns
            const int64_t next = std::min(NB2, rows);
            ACES_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));  // the panel is final
            if (side_pending) ACES_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
            if (!(skip & 8)) {
                // Only the first 128 columns of the next panel are needed by its first diagonal block
                // and panel; the other column blocks are needed one step later each, so they are
                // updated on the second stream (which also applies the in-panel updates to them: same
                // stream, so the two accumulations into a block are ordered).
                const int64_t first = std::min<int64_t>(DB, next);
                ACES_TRY(gemm(ctx, GEMM_NT, rows, first, kb, -1.0, P, lda, P, lda, 1.0, A + (K0 + kb) * (lda + 1), lda, 1));
                if (next > first) {
                    ACES_CUDA(ctx, cudaEventRecord(ctx->ev_crit, main_stream));
                    ACES_CUDA(ctx, cudaStreamWaitEvent(ctx->aux2_stream, ctx->ev_crit, 0));
                    // rows from the second block row on (the first block row of these columns lies above the diagonal)
                    ACES_TRY(on_aux2([&] {
                        return gemm(ctx, GEMM_NT, rows - first, next - first, kb, -1.0, P + first, lda, P + first, lda, 1.0,
                                    A + (K0 + kb + first) * (lda + 1), lda, 1);
                    }));
                    ACES_CUDA(ctx, cudaEventRecord(ctx->ev_aux2, ctx->aux2_stream));
                    aux2_pending = true;
                }
            }
            const int64_t rest = rows - next;
            if (rest > 0 && !(skip & 16)) {
                ACES_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
                ctx->stream = ctx->aux_stream;
                const int rc = gemm(ctx, GEMM_NT, rest, rest, kb, -1.0, P + next, lda, P + next, lda, 1.0,
                                    A + (K0 + kb + next) * (lda + 1), lda, 1);
                ctx->stream = main_stream;
                ACES_TRY(rc);
                ACES_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
                side_pending = true;
            }
        }
    }
    if (side_pending) ACES_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_join, 0));
    if (aux2_pending) ACES_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_aux2, 0));
    return ACES_OK;
}

namespace {
// flags of the persistent solves: one int per 128-block, zeroed before every launch
int solve_flags(aces_ctx* ctx, int nb, int** flags) {
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_flags, sizeof(int) * (size_t)nb));
    *flags = (int*)ctx->d_flags.ptr;
    ACES_CUDA(ctx, cudaMemsetAsync(*flags, 0, sizeof(int) * (size_t)nb, ctx->stream));
    return ACES_OK;
}
template <typename K>
int launch_persistent(aces_ctx* ctx, K kern, int nb, const double* L, int64_t ldl, const double* dinv,
                      const double* in, double* out, int* flags) {
    const int grid = std::min(nb, ctx->sm_count);
    void* args[] = {(void*)&L, (void*)&ldl, (void*)&dinv, (void*)&in, (void*)&out, (void*)&nb, (void*)&flags};
    ACES_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(TRSV_THREADS), args,
                                               TRSV_SMEM, ctx->stream));
    ctx->launches += 1;
    return ACES_OK;
}
}  // namespace

int trsv_lower(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* b,
               double* y) {
    ACES_CHECK(ctx, n % DB == 0, "dense::trsv_lower: not padded");
    const int nb = (int)(n / DB);
    int* flags = nullptr;
    ACES_TRY(configure(ctx));
    ACES_TRY(solve_flags(ctx, nb, &flags));
    return launch_persistent(ctx, trsv_fwd_persistent, nb, L, ldl, dinv, b, y, flags);
}

int trsv_lower_t(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* y,
                 double* x) {
    ACES_CHECK(ctx, n % DB == 0, "dense::trsv_lower_t: not padded");
    const int nb = (int)(n / DB);
    int* flags = nullptr;
    ACES_TRY(configure(ctx));
    ACES_TRY(solve_flags(ctx, nb, &flags));
    return launch_persistent(ctx, trsv_bwd_persistent, nb, L, ldl, dinv, y, x, flags);
}

int trsm_lower(aces_ctx* ctx, int64_t n, int64_t nrhs, const double* L, int64_t ldl,
               const double* dinv, double* B, int64_t ldb) {
    ACES_CHECK(ctx, n % DB == 0 && nrhs % DB == 0, "dense::trsm_lower: not padded");
    const int64_t nb = n / DB;
    for (int64_t k = 0; k < nb; ++k) {
        double* Bk = B + k * DB;
        // B_k <- inv(L_kk) * B_k   (in place, tile-local)
        ACES_TRY(gemm(ctx, GEMM_NN, DB, nrhs, DB, 1.0, dinv + k * DB * DB, DB, Bk, ldb, 0.0, Bk, ldb, 0));
        const int64_t rows = n - (k + 1) * DB;
        if (rows > 0)
            ACES_TRY(gemm(ctx, GEMM_NN, rows, nrhs, DB, -1.0, L + (k + 1) * DB + k * DB * ldl, ldl, Bk, ldb,
                          1.0, Bk + DB, ldb, 0));
    }
    return ACES_OK;
}

int trtri(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* X,
          int64_t ldx) {
    ACES_CHECK(ctx, n % DB == 0, "dense::trtri: not padded");
    ACES_TRY(set_identity(ctx, n, X, ldx));
    const int64_t nb = n / DB;
    for (int64_t k = 0; k < nb; ++k) {
        double* Xk = X + k * DB;  // block row k, columns 0 .. (k+1)*128
        const int64_t cols = (k + 1) * DB;
        ACES_TRY(gemm(ctx, GEMM_NN, DB, cols, DB, 1.0, dinv + k * DB * DB, DB, Xk, ldx, 0.0, Xk, ldx, 0));
        const int64_t rows = n - (k + 1) * DB;
        if (rows > 0)
            ACES_TRY(gemm(ctx, GEMM_NN, rows, cols, DB, -1.0, L + (k + 1) * DB + k * DB * ldl, ldl, Xk, ldx,
                          1.0, Xk + DB, ldx, 0));
    }
    return ACES_OK;
}

int set_identity(aces_ctx* ctx, int64_t n, double* A, int64_t lda) {
    if (n == 0) return ACES_OK;
    set_identity_kernel<<<(unsigned)std::min<int64_t>((n * n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(n, A, lda);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}
int get_diag(aces_ctx* ctx, int64_t n, const double* A, int64_t lda, double* d) {
    if (n == 0) return ACES_OK;
    get_diag_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, A, lda, d);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}
int fill_diag_range(aces_ctx* ctx, double* A, int64_t lda, int64_t from, int64_t to, double v) {
    if (to <= from) return ACES_OK;
    fill_diag_kernel<<<(unsigned)std::min<int64_t>((to - from + 255) / 256, 1024), 256, 0, ctx->stream>>>(A, lda, from, to, v);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}
int transpose(aces_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb) {
    if (m == 0 || n == 0) return ACES_OK;
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, ctx->stream>>>(m, n, A, lda, B, ldb);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}
int symmetrize_from_lower(aces_ctx* ctx, int64_t n, double* A, int64_t lda) {
    if (n == 0) return ACES_OK;
    const unsigned nt = (unsigned)((n + 31) / 32);
    dim3 grid(nt, nt), block(32, 8);
    symmetrize_kernel<<<grid, block, 0, ctx->stream>>>(n, A, lda);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace dense

extern "C" int32_t aces_potrf_schedule(int32_t nb, int32_t workers, int32_t* out, int64_t capacity, int64_t* n_tasks) {
    if (nb < 1 || nb >= 65536 || workers < 1 || !n_tasks) return ACES_E_INVALID;
    const std::vector<dense::DagTask> order = dense::build_dag_schedule(nb, workers);
    *n_tasks = (int64_t)order.size();
    if (out) {
        if (capacity < (int64_t)order.size()) return ACES_E_INVALID;
        for (size_t t = 0; t < order.size(); ++t) {
            const dense::DagTask& k = order[t];
            int32_t* o = out + 6 * t;
            o[0] = k.type; o[1] = k.slab | (k.pad << 8); o[2] = k.i; o[3] = k.j; o[4] = k.k0; o[5] = k.k1;
        }
    }
    return ACES_OK;
}

// Development hook: per-task timestamps (ns) of the last traced factorisation (ACES_POTRF_TRACE=1) together
// with the task list: 9 int64 per task (type, slab, i, j, k0, k1, t_grab, t_ready, t_done).
extern "C" int32_t aces_potrf_trace(aces_ctx* ctx, int64_t* out, int64_t capacity, int64_t* n_tasks) {
    if (!ctx || !n_tasks) return ACES_E_INVALID;
    *n_tasks = ctx->d_dag_trace.ptr ? ctx->dag_nlist + ctx->dag_nb : 0;
    if (!out || *n_tasks == 0) return ACES_OK;
    if (capacity < *n_tasks) return ACES_E_INVALID;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<dense::DagTask> tasks((size_t)*n_tasks);
    std::vector<unsigned long long> tr(3 * (size_t)*n_tasks);
    ACES_CUDA(ctx, cudaMemcpy(tasks.data(), ctx->d_dag_tasks.ptr, sizeof(dense::DagTask) * tasks.size(), cudaMemcpyDeviceToHost));
    ACES_CUDA(ctx, cudaMemcpy(tr.data(), ctx->d_dag_trace.ptr, sizeof(unsigned long long) * tr.size(), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < tasks.size(); ++t) {
        int64_t* o = out + 9 * t;
        o[0] = tasks[t].type; o[1] = tasks[t].slab; o[2] = tasks[t].i; o[3] = tasks[t].j; o[4] = tasks[t].k0; o[5] = tasks[t].k1;
        o[6] = (int64_t)tr[3 * t]; o[7] = (int64_t)tr[3 * t + 1]; o[8] = (int64_t)tr[3 * t + 2];
    }
    return ACES_OK;
}
