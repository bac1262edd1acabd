// eigen.cu -- all eigenvalues of a dense symmetric FP64 matrix on the GPU.
//
// Replaces LinearAlgebra.eigvals(::Symmetric{Float64, Matrix{Float64}}) (LAPACK dsyevr, JOBZ = 'N') as
// calc_merit calls it nine times per design (src/merit.jl:954-986).  Two stages:
//
//  1. Householder tridiagonalisation Q' A Q = T, blocked like LAPACK's dsytrd / dlatrd: the reflectors of a
//     panel of NB columns are accumulated as A - V W' - W V' and applied to the trailing matrix by one
//     DMMA GEMM per panel (dense::gemm, K = 2 NB); inside the panel every column costs one product of the
//     (not yet updated) trailing matrix with the new reflector -- the memory-bound half of the flops.
//     The trailing matrix is symmetric, so only its 128 x 128 tiles on and below the block diagonal are
//     read: an off-diagonal tile serves both A v and (transposed) its mirror image, sum_j 4 (n - j)^2
//     bytes = 4 n^3 / 3 in total.  A panel is ONE persistent cooperative kernel (a CTA per SM, two grid
//     barriers per column): the rows of V and W a CTA owns stay in its shared memory for the whole
//     panel, the tiles are dealt round robin, and every reduction runs in a fixed order (results are
//     deterministic).
//  2. Eigenvalues of T by multisection on Sturm counts (LAPACK dstebz's recurrence with its pivmin
//     guard): eight lanes x two points per eigenvalue, 17 sections per step; absolute accuracy
//     eps * ||T||, as dsyevr's.
//
// The caller passes the matrix in full (both triangles); the diagonal tiles are used in full.  It is scaled by a power of two to unit max-norm first (exact), as dsyev scales badly scaled input.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "aces_common.h"
#include "dense.cuh"
#include "eigen.cuh"

namespace eigen {

namespace {

constexpr int NB = 64;            // reflectors per panel
constexpr int THREADS = 512;
constexpr int NWARP = THREADS / 32;
constexpr int RB = 128;           // rows of one item of the trailing product
constexpr int KS = 8;             // the panel passes cut the NB reflectors into KS slices of KW
constexpr int KW = NB / KS;
constexpr int P1 = 2 * NB + 2;    // per-CTA partials of phase A: [1 + k] V'x, [1 + NB + k] W'x ([0] unused: |x|^2 rides on the barrier)
constexpr int MAX_SL = 6;         // 32-row slots a CTA can own (shared-memory bound)
#ifndef EIG_MIN_GRID
#define EIG_MIN_GRID 16
#endif

struct LLSlot;
struct PanelArgs {
    const double* A;
    int64_t lda;
    int n, Np;
    int j0, ncols;
    double *P, *Q;      // Np x 2NB each: P = [V | W], Q = [W | V]
    double *d, *e;
    double* ypart;      // [Np / 128][Np]: partial products, [source block][target row]
    double* part1;      // [grid][P1]
    double* pq;         // [2 NB]: V'v then W'v
    LLSlot* ll;         // [2][grid] partial sums + one slot holding the arrival counter
    unsigned epoch0;    // barriers executed by the panels before this one
    int SL;
};

// Grid barrier fused with a reduction: every CTA stores two doubles, arrives on a monotonic counter, waits for
// the whole grid and sums everybody's doubles in a fixed order.  All CTAs are co-resident (cooperative launch).
// bidx is the running index of the barrier within the factorisation (uniform over the grid): the slot array
// alternates with its parity (a CTA that has passed barrier b + 1 knows every CTA has read the slots of
// barrier b), the counter never resets (target (bidx + 1) * grid; zeroed before the first panel).
// Measured alternatives at N = 6779 / 1024 (device-resident, profiles/eig_r02_barrier_variants.txt): all-to-all
// polling of (value, epoch) words packed NCCL-LL style, one round of loads instead of three round trips:
// 358 / 25 ms -- 148 x 148 polled slots congest the L2; the same with back-off 372 / 27 ms; the last arriver
// summing and handing every CTA the result in its own slot 251 / 19 ms; this one 242 / 18 ms.
struct LLSlot {
    unsigned long long w[4];
};
// v0, v1: this CTA's contributions (thread 0's are used); out[0], out[1] (shared memory): the sums over the grid
__device__ __noinline__ void ll_barrier(LLSlot* ll, unsigned bidx, double v0, double v1, double* out) {
    const int G = gridDim.x;
    LLSlot* slots = ll + (bidx & 1u) * G;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned* counter = reinterpret_cast<unsigned*>(ll + 2 * G);
        if (lane == 0) {
            double* mine = reinterpret_cast<double*>(slots[blockIdx.x].w);
            __stcg(mine, v0);
            __stcg(mine + 1, v1);
            __threadfence();  // release: everything this CTA wrote before the barrier
            atomicAdd(counter, 1u);
            const unsigned target = (bidx + 1u) * (unsigned)G;
            unsigned seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            } while (seen < target);
        }
        __syncwarp();
        __threadfence();  // acquire: what the other CTAs wrote before they arrived
        double s0 = 0.0, s1 = 0.0;
        for (int c = lane; c < G; c += 32) {
            const double* p = reinterpret_cast<const double*>(slots[c].w);
            s0 += __ldcg(p);
            s1 += __ldcg(p + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) {
            out[0] = s0;
            out[1] = s1;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// sum over the CTAs of one entry of a per-CTA partial table, by one warp, in a fixed order
__device__ __forceinline__ double sum_over_ctas(const double* tab, int stride, int G, int lane) {
    double s = 0.0;
    for (int c = lane; c < G; c += 32) s += __ldcg(tab + (int64_t)c * stride);
    return warp_sum(s);
}

size_t panel_smem(int SL) {
    const size_t R = (size_t)SL * 32;
    const size_t un = std::max((size_t)(SL < MAX_SL ? 2 : 1) * (NWARP * RB + RB), (size_t)KS * R + std::max((size_t)SL * 2 * NB, (size_t)KS * R));
    return sizeof(double) * (2 * NB * R + R + 4 * NB + 32 + un);
}

__global__ void __launch_bounds__(THREADS, 1) sytrd_panel_kernel(PanelArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int R = a.SL * 32;
    double* Vs = sm;                 // [NB][R] rows of V this CTA owns
    double* Ws = Vs + NB * R;        // [NB][R]
    double* xs = Ws + NB * R;        // [R] the current column (before normalisation)
    double* Vrow = xs + R;           // V[j, 0..jj) of the column being reduced
    double* Wrow = Vrow + NB;
    double* ps = Wrow + NB;          // W'v
    double* qs = ps + NB;            // V'v
    double* scal = qs + NB;          // [32] tau, scale, alpha, beta, alpha2, scratch, barrier sums
    double* un = scal + 32;
    double* red = un;                // [KS][R]
    double* dots = red + KS * R;     // [SL][2 NB] (phase A), [YG <= KS][R] (phase C)
    double* redg = un;               // [1 or 2][NWARP + 1][RB] (phase B only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n = a.n, Np = a.Np;
    unsigned bidx = a.epoch0;  // barriers of the factorisation so far
#ifdef EIG_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pc = clock64();
#define EIG_MARK(k) do { const long long now_ = clock64(); pt[k] += now_ - pc; pc = now_; } while (0)
#else
#define EIG_MARK(k) do { } while (0)
#endif
    double anext = 0.0;   // A[own row, j + 1], fetched one column ahead
    // global row of local row li
    auto grow = [&](int li) -> int64_t { return ((int64_t)((li >> 5) * G + cta)) * 32 + (li & 31); };

    for (int jj = 0; jj < a.ncols; ++jj) {
        const int j = a.j0 + jj;
        const bool reflect = j + 2 < n;
        // ---------------- phase A: column j with the panel's reflectors applied -----------------------
        for (int u = warp; u < a.SL * KS; u += NWARP) {
            const int li = (u / KS) * 32 + lane, ks = u % KS;
            const int k1 = min(jj, (ks + 1) * KW);
            double c = 0.0;
            for (int k = ks * KW; k < k1; ++k) c += Vs[k * R + li] * Wrow[k] + Ws[k * R + li] * Vrow[k];
            red[ks * R + li] = c;
        }
        __syncthreads();
        double nrm = 0.0;
        if (tid < R) {
            const int64_t i = grow(tid);
            double x = 0.0;
            if (i >= j && i < n) {
                double c = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) c += red[ks * R + tid];
                x = (jj > 0 ? anext : a.A[i + (int64_t)j * a.lda]) - c;
                if (i == j) {
                    a.d[j] = x;
                    x = 0.0;
                } else if (reflect) {
                    __stcg(a.P + i + (int64_t)jj * Np, x);
                    if (i >= j + 2) nrm = x * x;
                } else if (i == j + 1) {
                    a.e[j] = x;
                }
            }
            xs[tid] = x;
        }
        if (!reflect) {
            // the last two columns only give d (and e): no reflector (a zero column of V and W).  Uniform
            // over the grid.  The next column needs row j + 1 of V and W, whose newest entry was stored by
            // its owner after the last barrier: one more barrier, then read the row from global memory.
            if (tid < R) Vs[jj * R + tid] = Ws[jj * R + tid] = 0.0;
            ll_barrier(a.ll, bidx++, 0.0, 0.0, scal + 6);
            if (tid < R && j + 1 < n) {
                const int64_t i = grow(tid);
                anext = (i >= j + 1 && i < n) ? a.A[i + (int64_t)(j + 1) * a.lda] : 0.0;
            }
            if (tid <= jj) {
                const bool have = tid < jj && j + 1 < n;
                Vrow[tid] = have ? __ldcg(a.P + (j + 1) + (int64_t)tid * Np) : 0.0;
                Wrow[tid] = have ? __ldcg(a.P + (j + 1) + (int64_t)(NB + tid) * Np) : 0.0;
            }
            __syncthreads();
            continue;
        }
        nrm = warp_sum(nrm);
        if (lane == 0 && warp < a.SL) scal[8 + warp] = nrm;
        if (tid < R && grow(tid) == j + 1) scal[14] = xs[tid];  // alpha, on the CTA that owns row j + 1
        __syncthreads();  // xs complete
        for (int u = warp; u < a.SL * KS; u += NWARP) {
            const int slot = u / KS, ks = u % KS, li = slot * 32 + lane;
            const int k1 = min(jj, (ks + 1) * KW);
            const double x = xs[li];
            for (int k = ks * KW; k < k1; ++k) {
                const double tv = warp_sum(Vs[k * R + li] * x);
                const double tw = warp_sum(Ws[k * R + li] * x);
                if (lane == 0) {
                    dots[slot * 2 * NB + k] = tv;
                    dots[slot * 2 * NB + NB + k] = tw;
                }
            }
        }
        __syncthreads();
        if (tid < 2 * NB && (tid & (NB - 1)) < jj) {
            double s = 0.0;
            for (int slot = 0; slot < a.SL; ++slot) s += dots[slot * 2 * NB + tid];
            __stcg(a.part1 + (int64_t)cta * P1 + 1 + tid, s);
        }
        {
            // |x|^2 (rows >= j + 2) and alpha = x[j + 1] (non-zero on its owner only) ride on the barrier
            double mine = 0.0, alpha_mine = 0.0;
            if (tid == 0) {
                for (int w = 0; w < a.SL; ++w) mine += scal[8 + w];
                const int64_t owner_blk = (j + 1) / 32;
                if ((int)(owner_blk % G) == cta) alpha_mine = scal[14];
            }
            EIG_MARK(0);
            ll_barrier(a.ll, bidx++, mine, alpha_mine, scal + 6);
            EIG_MARK(1);
        }

        // ---------------- phase B: the reflector, then (trailing matrix) * v ---------------------------
        // warp 0: tau, scale; meanwhile warps 1.. sum this CTA's entries of V'x / W'x over the grid
        double pq_sum = 0.0, pq_coef = 0.0;
        int pq_e = -1;
        if (warp == 0) {
            const double nrm2 = scal[6];
            if (lane == 0) {
                const double alpha = scal[7];
                double tau = 0.0, scale = 0.0, beta = alpha;
                if (nrm2 > 0.0) {
                    beta = -copysign(sqrt(alpha * alpha + nrm2), alpha);
                    tau = (beta - alpha) / beta;
                    scale = 1.0 / (alpha - beta);
                }
                scal[0] = tau;
                scal[1] = scale;
                scal[2] = alpha;
                scal[3] = beta;
                if (cta == 0) a.e[j] = beta;
            }
        } else {
            const int e = cta + (warp - 1) * G;  // entry of [V'x | W'x] this warp reduces over the grid
            if (e < 2 * NB && (e & (NB - 1)) < jj) {
                pq_e = e;
                pq_sum = sum_over_ctas(a.part1 + 1 + e, P1, G, lane);
                pq_coef = __ldcg(a.P + (j + 1) + (int64_t)e * Np);
            }
        }
        __syncthreads();
        const double tau = scal[0], scale = scal[1];
        // v = scale * x except v[j+1] = 1:  V'v = scale V'x + V[j+1, :] (1 - scale alpha)
        if (pq_e >= 0 && lane == 0) __stcg(a.pq + pq_e, scale * pq_sum + pq_coef * (1.0 - scale * scal[2]));
        auto vcol = [&](int c) -> double {
            if (c <= j || c >= n) return 0.0;
            if (c == j + 1) return 1.0;
            return __ldcg(a.P + c + (int64_t)jj * Np) * scale;
        };
        {
            // The trailing matrix is symmetric: only its 128 x 128 tiles on and below the block diagonal
            // are read.  Tile (rb, cb), rb > cb, gives rows rb of A v from columns cb AND (transposed) rows
            // cb from columns rb; a diagonal tile (stored in full) only the former.  Partial results go to
            // ypart[source block][target row]: the two uses of a tile never share a slot.
            const int b0 = (j + 1) / RB, m = (n + RB - 1) / RB - b0;
            const int tiles = m * (m + 1) / 2;
            double yv = 0.0;
            // tile t of the lower triangle, row-major: (r, c) with t = r (r + 1) / 2 + c
            auto coords = [&](int t, int& rb, int& cb) {
                int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while (r * (r + 1) / 2 > t) --r;
                while ((r + 1) * (r + 2) / 2 <= t) ++r;
                rb = b0 + r;
                cb = b0 + (t - r * (r + 1) / 2);
            };
            // A warp owns eight columns of a tile (warp + 16 q); they are loaded in two halves of four
            // and the loads run one half ahead of the arithmetic, ACROSS tiles: the first half of the next
            // tile (and its slice of v) is in flight while this tile is reduced through shared memory.
            double2 lo0[4], hi0[4], lo1[4], hi1[4];
            double vvN = 0.0, vrN0 = 0.0, vrN1 = 0.0, vrN2 = 0.0, vrN3 = 0.0;
            int rbN = 0, cbN = 0;
            auto load_half = [&](double2* lo, double2* hi, int rb, int cb, int half) {
                const double* Ab = a.A + (int64_t)rb * RB + 2 * lane + (int64_t)cb * RB * a.lda;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double* p = Ab + (int64_t)(warp + 16 * (4 * half + q)) * a.lda;
                    lo[q] = __ldg(reinterpret_cast<const double2*>(p));
                    hi[q] = __ldg(reinterpret_cast<const double2*>(p + 64));
                }
            };
            auto load_vectors = [&](int rb, int cb) {
                vvN = vcol(cb * RB + warp + 16 * (lane & 7));  // lane q < 8 holds v of column warp + 16 q
                const int r0 = rb * RB + 2 * lane;  // v of the rows this lane holds (transposed product, v'(A v))
                vrN0 = vcol(r0);
                vrN1 = vcol(r0 + 1);
                vrN2 = vcol(r0 + 64);
                vrN3 = vcol(r0 + 65);
            };
            EIG_MARK(2);
            int t = cta, it = 0;
            const int nbuf = a.SL < MAX_SL ? 2 : 1;
            if (t < tiles) {
                coords(t, rbN, cbN);
                load_vectors(rbN, cbN);
                load_half(lo0, hi0, rbN, cbN, 0);
            }
            while (t < tiles) {
                const int rb = rbN, cb = cbN;
                const bool offdiag = rb != cb;
                const double vv = vvN, vr0 = vrN0, vr1 = vrN1, vr2 = vrN2, vr3 = vrN3;
                load_half(lo1, hi1, rb, cb, 1);
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, tmine = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double v = __shfl_sync(0xffffffffu, vv, q);
                    a0 += lo0[q].x * v;
                    a1 += lo0[q].y * v;
                    a2 += hi0[q].x * v;
                    a3 += hi0[q].y * v;
                    if (offdiag) {  // uniform over the CTA
                        const double tq = warp_sum((lo0[q].x * vr0 + lo0[q].y * vr1) + (hi0[q].x * vr2 + hi0[q].y * vr3));
                        if (lane == q) tmine = tq;
                    }
                }
                const int tn = t + G;
                if (tn < tiles) {
                    coords(tn, rbN, cbN);
                    load_vectors(rbN, cbN);
                    load_half(lo0, hi0, rbN, cbN, 0);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double v = __shfl_sync(0xffffffffu, vv, 4 + q);
                    a0 += lo1[q].x * v;
                    a1 += lo1[q].y * v;
                    a2 += hi1[q].x * v;
                    a3 += hi1[q].y * v;
                    if (offdiag) {
                        const double tq = warp_sum((lo1[q].x * vr0 + lo1[q].y * vr1) + (hi1[q].x * vr2 + hi1[q].y * vr3));
                        if (lane == 4 + q) tmine = tq;
                    }
                }
                if (offdiag && lane < 8) {
                    const int c = cb * RB + warp + 16 * lane;
                    __stcg(a.ypart + (int64_t)rb * Np + c, tmine);
                    yv += vv * tmine;
                }
                // cross-warp sum of the direct product through shared memory; with two buffers (all but the
                // largest row-slot count) one barrier per tile is enough: the buffer is written again two
                // tiles later, after the barrier of the tile in between
                double* rbuf = redg + (size_t)(it & (nbuf - 1)) * (NWARP * RB + RB);
                rbuf[warp * RB + 2 * lane] = a0;
                rbuf[warp * RB + 2 * lane + 1] = a1;
                rbuf[warp * RB + 64 + 2 * lane] = a2;
                rbuf[warp * RB + 64 + 2 * lane + 1] = a3;
                if (warp == 0) {
                    double* vsm = rbuf + NWARP * RB;
                    vsm[2 * lane] = vr0;
                    vsm[2 * lane + 1] = vr1;
                    vsm[64 + 2 * lane] = vr2;
                    vsm[64 + 2 * lane + 1] = vr3;
                }
                __syncthreads();
                if (tid < RB) {
                    double sd = 0.0;
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) sd += rbuf[w * RB + tid];
                    __stcg(a.ypart + (int64_t)cb * Np + rb * RB + tid, sd);
                    yv += rbuf[NWARP * RB + tid] * sd;
                }
                if (nbuf == 1) __syncthreads();
                t = tn;
                ++it;
            }
            __syncthreads();
            // v'(A v): this CTA's part rides on the barrier
            yv = warp_sum(yv);
            if (lane == 0) redg[warp] = yv;
            __syncthreads();
            double sy = 0.0;
            if (tid == 0)
                for (int w = 0; w < NWARP; ++w) sy += redg[w];
            EIG_MARK(3);
            ll_barrier(a.ll, bidx++, sy, 0.0, scal + 16);
            EIG_MARK(4);
        }

        // ---------------- phase C: w = tau (A v - V (W'v) - W (V'v)) - (tau/2)(w'v) v --------------------
        // every global load of this phase is issued before the first barrier, so they overlap: V'v / W'v,
        // the partial products of the own rows and of row j + 1, row j + 1 of V and W, and column j + 1 of
        // A for the next phase A
        const int cc0 = (j + 1) / RB, ncc = (n + RB - 1) / RB - cc0;  // source blocks of the partial products
        const int YG = min(THREADS / R, KS);
        double* ysum = dots;  // [YG][R]: dots is dead in this phase
        if (tid < jj) {
            qs[tid] = __ldcg(a.pq + tid);
            ps[tid] = __ldcg(a.pq + NB + tid);
        }
        double vn[2] = {0.0, 0.0}, wn[2] = {0.0, 0.0}, ynext = 0.0;
        if (warp == 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                if (k < jj) {
                    vn[h] = __ldcg(a.P + (j + 1) + (int64_t)k * Np);
                    wn[h] = __ldcg(a.P + (j + 1) + (int64_t)(NB + k) * Np);
                }
            }
            for (int cc = lane; cc < ncc; cc += 32) ynext += __ldcg(a.ypart + (int64_t)(cc0 + cc) * Np + (j + 1));
        }
        if (tid < YG * R) {
            const int g = tid / R, li = tid - g * R;
            const int64_t i = grow(li);
            double y0 = 0.0, y1 = 0.0;
            if (i >= j + 1 && i < n) {
                int cc = g;
                for (; cc + YG < ncc; cc += 2 * YG) {
                    y0 += __ldcg(a.ypart + (int64_t)(cc0 + cc) * Np + i);
                    y1 += __ldcg(a.ypart + (int64_t)(cc0 + cc + YG) * Np + i);
                }
                if (cc < ncc) y0 += __ldcg(a.ypart + (int64_t)(cc0 + cc) * Np + i);
            }
            ysum[g * R + li] = y0 + y1;
        }
        if (tid < R) {
            const int64_t i = grow(tid);
            anext = (i >= j + 1 && i < n) ? a.A[i + (int64_t)(j + 1) * a.lda] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            double t = 0.0;
            for (int k = lane; k < jj; k += 32) t += ps[k] * qs[k];
            t = warp_sum(t);
            if (lane == 0) scal[4] = -0.5 * tau * (tau * (scal[16] - 2.0 * t));
        }
        if (warp == 1) {
            // row j + 1 of V and W including the column being formed: every CTA needs it for the next column
            double c2 = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                if (k < jj) c2 += vn[h] * ps[k] + wn[h] * qs[k];
            }
            c2 = warp_sum(c2);
            ynext = warp_sum(ynext);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                if (k < jj) {
                    Vrow[k] = vn[h];
                    Wrow[k] = wn[h];
                }
            }
            if (lane == 0) {
                Vrow[jj] = 1.0;
                scal[5] = tau * (ynext - c2);  // + alpha2 below, once warp 0 has it
            }
        }
        for (int u = warp; u < a.SL * KS; u += NWARP) {
            const int li = (u / KS) * 32 + lane, ks = u % KS;
            const int k1 = min(jj, (ks + 1) * KW);
            double c = 0.0;
            for (int k = ks * KW; k < k1; ++k) c += Vs[k * R + li] * ps[k] + Ws[k * R + li] * qs[k];
            red[ks * R + li] = c;
        }
        __syncthreads();
        const double alpha2 = scal[4];
        if (tid == 0) Wrow[jj] = scal[5] + alpha2;
        if (tid < R) {
            const int64_t i = grow(tid);
            double v = 0.0, w = 0.0;
            if (i >= j + 1 && i < n) {
                double y = 0.0, c = 0.0;
                for (int g = 0; g < YG; ++g) y += ysum[g * R + tid];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) c += red[ks * R + tid];
                v = (i == j + 1) ? 1.0 : xs[tid] * scale;
                w = tau * (y - c) + alpha2 * v;
                __stcg(a.P + i + (int64_t)jj * Np, v);
                __stcg(a.P + i + (int64_t)(NB + jj) * Np, w);
                __stcg(a.Q + i + (int64_t)jj * Np, w);
                __stcg(a.Q + i + (int64_t)(NB + jj) * Np, v);
            }
            Vs[jj * R + tid] = v;
            Ws[jj * R + tid] = w;
        }
        __syncthreads();
        EIG_MARK(5);
    }
#ifdef EIG_PROFILE
    if (tid == 0 && (cta == 0 || cta == G - 1 || cta == G / 2))
        printf("panel j0=%d cta=%d cycles: A %lld | bar1 %lld | B-pre %lld | tiles %lld | bar2 %lld | C %lld\n", a.j0, cta, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5]);
#endif
}

// ---- preparation ---------------------------------------------------------------------------------------
__global__ void maxabs_kernel(int64_t n, const double* A, int64_t lda, double* part) {
    double m = 0.0;
    for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = fabs(A[i + j * lda]);
            m = (v > m || v != v) ? v : m;  // a NaN sticks
        }
    __shared__ double s[256];
    s[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double b = s[threadIdx.x + o], c = s[threadIdx.x];
            s[threadIdx.x] = (b > c || b != b) ? b : c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

// sc[0] = power of two that brings max|a_ij| into [1, 2), sc[1] = its inverse (1, 1 for a zero matrix)
__global__ void scale_factor_kernel(int nparts, const double* part, double* sc) {
    double m = 0.0;
    for (int i = 0; i < nparts; ++i) m = (part[i] > m || part[i] != part[i]) ? part[i] : m;
    double f = 1.0, fi = 1.0;
    if (m > 0.0 && m < 1.0e308) {
        const int ex = ilogb(m);
        f = scalbn(1.0, -ex);
        fi = scalbn(1.0, ex);
    }
    sc[0] = f;
    sc[1] = fi;
    sc[2] = m;
}

// scales the n x n block and zeroes the padding rows / columns of the Np x Np buffer
__global__ void scale_clear_kernel(int64_t n, int64_t Np, double* A, int64_t lda, const double* sc) {
    const double f = sc[0];
    for (int64_t j = blockIdx.x; j < Np; j += gridDim.x)
        for (int64_t i = threadIdx.x; i < Np; i += blockDim.x) {
            double* p = A + i + j * lda;
            *p = (i < n && j < n) ? *p * f : 0.0;
        }
}

// ---- eigenvalues of the tridiagonal matrix: bisection on Sturm counts ------------------------------------
constexpr int BIS_THREADS = 128;

__global__ void tridiag_bounds_kernel(int n, const double* d, const double* e, double* e2out, double* bnd) {
    // Gershgorin interval, |T|_1 and max e^2 (one block)
    double lo = INFINITY, hi = -INFINITY, tn = 0.0, e2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < n ? fabs(e[i]) : 0.0;
        lo = fmin(lo, d[i] - el - er);
        hi = fmax(hi, d[i] + el + er);
        tn = fmax(tn, fabs(d[i]) + el + er);
        e2 = fmax(e2, er * er);
        e2out[i] = el * el;  // e2out[i] = e[i-1]^2, 0 for the first row
    }
    __shared__ double s[4][256];
    s[0][threadIdx.x] = lo;
    s[1][threadIdx.x] = hi;
    s[2][threadIdx.x] = tn;
    s[3][threadIdx.x] = e2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s[0][threadIdx.x] = fmin(s[0][threadIdx.x], s[0][threadIdx.x + o]);
            s[1][threadIdx.x] = fmax(s[1][threadIdx.x], s[1][threadIdx.x + o]);
            s[2][threadIdx.x] = fmax(s[2][threadIdx.x], s[2][threadIdx.x + o]);
            s[3][threadIdx.x] = fmax(s[3][threadIdx.x], s[3][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double eps = 2.220446049250313e-16, safemin = 2.2250738585072014e-308;
        const double tnorm = s[2][0];
        const double pivmin = safemin * fmax(1.0, s[3][0]);
        bnd[0] = s[0][0] - 2.0 * tnorm * eps * n - 2.0 * pivmin;
        bnd[1] = s[1][0] + 2.0 * tnorm * eps * n + 2.0 * pivmin;
        bnd[2] = tnorm;
        bnd[3] = pivmin;
    }
}

// Eigenvalue k (ascending), scaled back by sc[1].  Eight lanes per eigenvalue, two Sturm counts per lane:
// every step cuts the bracket into 17 parts (14 steps for 54 bits instead of 54).  The new bracket is the
// largest point whose count is <= k and the smallest whose count is > k, whatever the counts in between.
constexpr int LPE = 8;
__global__ void __launch_bounds__(BIS_THREADS) tridiag_bisect_kernel(int n, const double* __restrict__ d,
                                                                     const double* __restrict__ e2,
                                                                     const double* __restrict__ bnd,
                                                                     const double* __restrict__ sc, double* w) {
    const int sub = threadIdx.x & (LPE - 1);
    const int k = (int)((blockIdx.x * (unsigned)BIS_THREADS + threadIdx.x) / LPE);
    const double eps = 2.220446049250313e-16;
    double lo = bnd[0], hi = bnd[1];
    const double tnorm = bnd[2], pivmin = bnd[3];
    const double atol = 0.25 * eps * tnorm + 2.0 * pivmin;
    bool active = k < n;
    for (int it = 0; it < 64; ++it) {
        active = active && (hi - lo > atol + 2.0 * eps * fmax(fabs(lo), fabs(hi)));
        if (!__any_sync(0xffffffffu, active)) break;
        const double h = (hi - lo) * (1.0 / 17.0);
        const double x0 = lo + h * (double)(2 * sub + 1), x1 = lo + h * (double)(2 * sub + 2);
        int c0 = 0, c1 = 0;
        double q0 = 1.0, q1 = 1.0;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const double di = __ldg(d + i), ei = __ldg(e2 + i);
            q0 = di - x0 - ei / q0;
            q1 = di - x1 - ei / q1;
            if (fabs(q0) < pivmin) q0 = -pivmin;
            if (fabs(q1) < pivmin) q1 = -pivmin;
            c0 += q0 < 0.0;
            c1 += q1 < 0.0;
        }
        double nlo = lo, nhi = hi;
        if (c0 <= k) nlo = fmax(nlo, x0); else nhi = fmin(nhi, x0);
        if (c1 <= k) nlo = fmax(nlo, x1); else nhi = fmin(nhi, x1);
#pragma unroll
        for (int o = LPE / 2; o > 0; o >>= 1) {
            nlo = fmax(nlo, __shfl_xor_sync(0xffffffffu, nlo, o));
            nhi = fmin(nhi, __shfl_xor_sync(0xffffffffu, nhi, o));
        }
        if (active) {
            if (!(nlo > lo || nhi < hi) || !(nlo < nhi)) active = false;  // no representable point left inside
            else { lo = nlo; hi = nhi; }
        }
    }
    if (k < n && sub == 0) w[k] = 0.5 * (lo + hi) * sc[1];
}

struct Work {
    double *P, *Q, *ypart, *d, *e, *e2, *part1, *pq, *sc, *bnd, *mx;
    LLSlot* ll;
};

int carve(aces_ctx* ctx, int64_t Np, int grid, Work* w) {
    size_t doubles = (size_t)Np * 2 * NB * 2 + (size_t)(Np / RB) * Np + 3 * (size_t)Np + (size_t)grid * P1 + 2 * NB + 8 + 8 + 1024 + (size_t)grid * 8 + 16;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_eig, sizeof(double) * doubles + 256));
    double* p = (double*)ctx->d_eig.ptr;
    w->P = p; p += (size_t)Np * 2 * NB;
    w->Q = p; p += (size_t)Np * 2 * NB;
    w->ypart = p; p += (size_t)(Np / RB) * Np;
    w->d = p; p += Np;
    w->e = p; p += Np;
    w->e2 = p; p += Np;
    w->part1 = p; p += (size_t)grid * P1;
    w->pq = p; p += 2 * NB;
    w->sc = p; p += 8;
    w->bnd = p; p += 8;
    w->mx = p; p += 1024;
    w->ll = (LLSlot*)p;  // [2][grid], 32 bytes each (p is 8-byte aligned: 64-bit words)
    return ACES_OK;
}

}  // namespace

int64_t max_order(const aces_ctx* ctx) { return (int64_t)MAX_SL * 32 * ctx->sm_count; }

int syev_values(aces_ctx* ctx, int64_t n, int64_t Np, double* A, int64_t lda, double* w_dev) {
    ACES_CHECK(ctx, n >= 1 && Np % dense::DB == 0 && Np >= n && lda >= Np && lda % 2 == 0,
               "eigen::syev_values: the matrix is not padded to a multiple of 128");
    ACES_CHECK(ctx, n <= max_order(ctx), "eigen::syev_values: order %lld exceeds the supported %lld", (long long)n,
               (long long)max_order(ctx));
    // One CTA per SM for large matrices; a small matrix has fewer 128 x 128 tiles than SMs, and a grid barrier
    // among fewer CTAs is cheaper (less skew, fewer arrivals on the counter's L2 slice), and leaves SMs to other
    // contexts (replicas of calc_merit on one GPU).
    const int64_t mt = (n + RB - 1) / RB;
    const int grid = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(mt * (mt + 1) / 2, EIG_MIN_GRID));
    ACES_CHECK(ctx, grid * (NWARP - 1) >= 2 * NB && ctx->sm_count >= grid, "eigen::syev_values: too few SMs");
    int SL = (int)(((n + 31) / 32 + grid - 1) / grid);
    SL = std::max(SL, 1);
    const size_t smem = panel_smem(SL);
    ACES_CHECK(ctx, (int)smem <= ctx->smem_optin, "eigen::syev_values: panel needs %zu bytes of shared memory", smem);
    ACES_CUDA(ctx, cudaFuncSetAttribute(sytrd_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Work w;
    ACES_TRY(carve(ctx, Np, grid, &w));
    cudaStream_t st = ctx->stream;
    // scale to unit max-norm by a power of two, clear the padding
    maxabs_kernel<<<1024, 256, 0, st>>>(n, A, lda, w.mx);
    scale_factor_kernel<<<1, 1, 0, st>>>(1024, w.mx, w.sc);
    scale_clear_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(n, Np, A, lda, w.sc);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    ACES_CUDA(ctx, cudaMemsetAsync(w.ll, 0, sizeof(LLSlot) * (2 * (size_t)grid + 1), st));
    unsigned epoch0 = 0;  // epoch 0 = never published
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
        const int ncols = (int)std::min<int64_t>(NB, n - j0);
        ACES_CUDA(ctx, cudaMemsetAsync(w.P, 0, sizeof(double) * (size_t)Np * 2 * NB * 2, st));  // P and Q are adjacent
        PanelArgs pa;
        pa.A = A; pa.lda = lda; pa.n = (int)n; pa.Np = (int)Np; pa.j0 = (int)j0; pa.ncols = ncols;
        pa.P = w.P; pa.Q = w.Q; pa.d = w.d; pa.e = w.e; pa.ypart = w.ypart; pa.part1 = w.part1;
        pa.pq = w.pq; pa.ll = w.ll; pa.SL = SL; pa.epoch0 = epoch0;
        for (int64_t j = j0; j < j0 + ncols; ++j) epoch0 += (j + 2 < n) ? 2u : 1u;  // barriers of this panel
        void* args[] = {(void*)&pa};
        ACES_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)sytrd_panel_kernel, dim3((unsigned)grid), dim3(THREADS),
                                                   args, smem, st));
        ctx->launches += 1;
        const int64_t t0 = j0 + ncols;
        if (t0 < n) {
            // trailing matrix -= V W' + W V' (from the 128-aligned row at or above t0: the rows in between
            // belong to columns that are already reduced and are never read again)
            const int64_t r0 = t0 / dense::DB * dense::DB;
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NT, Np - r0, Np - r0, 2 * NB, -1.0, w.P + r0, Np, w.Q + r0, Np, 1.0,
                                 A + r0 * (lda + 1), lda, 1));
        }
    }
    tridiag_bounds_kernel<<<1, 256, 0, st>>>((int)n, w.d, w.e, w.e2, w.bnd);
    tridiag_bisect_kernel<<<(unsigned)((n * LPE + BIS_THREADS - 1) / BIS_THREADS), BIS_THREADS, 0, st>>>((int)n, w.d, w.e2, w.bnd,
                                                                                                 w.sc, w_dev);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return ACES_OK;
}

namespace {
// lower triangle <- upper triangle (Julia's Symmetric(A) reads uplo = :U, as does dsyevr called with it)
__global__ void mirror_upper_kernel(int64_t n, double* A, int64_t lda) {
    for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
        for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[i + j * lda] = A[j + i * lda];
}
}  // namespace

int mirror_upper(aces_ctx* ctx, int64_t n, double* A, int64_t lda) {
    mirror_upper_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(n, A, lda);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace eigen

extern "C" int32_t aces_symmetric_eigenvalues(aces_ctx* ctx, int64_t n, const double* a, int64_t lda,
                                              double* eigenvalues) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, n >= 1 && a && eigenvalues && lda >= n, "symmetric_eigenvalues: bad argument");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t Np = dense::pad_up(n);
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_eigA, sizeof(double) * (size_t)Np * Np + sizeof(double) * (size_t)Np));
    double* A = (double*)ctx->d_eigA.ptr;
    double* w = A + (size_t)Np * Np;
    ACES_CUDA(ctx, cudaMemsetAsync(A, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpy2DAsync(A, sizeof(double) * (size_t)Np, a, sizeof(double) * (size_t)lda,
                                     sizeof(double) * (size_t)n, (size_t)n, cudaMemcpyDefault, ctx->stream));
    ACES_TRY(eigen::mirror_upper(ctx, n, A, Np));
    ACES_TRY(eigen::syev_values(ctx, n, Np, A, Np, w));
    ACES_CUDA(ctx, cudaMemcpyAsync(eigenvalues, w, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}
