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
constexpr int P1 = 2 * NB + 2;    // per-CTA partials of phase A: [0] |x|^2, [1 + k] V'x, [1 + NB + k] W'x
constexpr int MAX_SL = 6;         // 32-row slots a CTA can own (shared-memory bound)

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
    LLSlot* ll;         // [2][grid]
    int SL;
};

// Grid barrier fused with a reduction, in the style of NCCL's LL protocol: every CTA publishes two doubles as
// four 64-bit words, each carrying 32 bits of payload and the 32-bit epoch of this barrier, so a reader
// that sees the epoch in a word has the payload of that word too (64-bit stores are single-copy atomic).
// One warp of every CTA polls all slots: the barrier and the cross-CTA sums behind it cost ONE round of
// loads instead of an atomic, a poll and a read.  The sums run over the CTAs in a fixed order.  All CTAs
// are co-resident (cooperative launch).  Two slot arrays alternate (a CTA that has passed barrier b + 1
// knows every CTA has read the slots of barrier b), epochs are unique within one factorisation and the
// arrays are zeroed before it.
struct LLSlot {
    unsigned long long w[4];
};
__device__ __forceinline__ void ll_publish(LLSlot* slot, double v0, double v1, unsigned epoch) {
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long ep = (unsigned long long)epoch << 32;
    const unsigned long long w0 = (b0 & 0xffffffffull) | ep, w1 = (b0 >> 32) | ep;
    const unsigned long long w2 = (b1 & 0xffffffffull) | ep, w3 = (b1 >> 32) | ep;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot->w), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot->w + 2), "l"(w2), "l"(w3) : "memory");
}
__device__ __forceinline__ void ll_wait(const LLSlot* slot, unsigned epoch, double* v0, double* v1) {
    unsigned long long w0, w1, w2, w3;
    do {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot->w) : "memory");
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(slot->w + 2) : "memory");
    } while ((unsigned)(w0 >> 32) != epoch || (unsigned)(w1 >> 32) != epoch || (unsigned)(w2 >> 32) != epoch ||
             (unsigned)(w3 >> 32) != epoch);
    *v0 = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    *v1 = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
}
// v0, v1: this CTA's contributions (thread 0's are used); out[0], out[1] (shared memory): the sums over the grid
__device__ __forceinline__ void ll_barrier(LLSlot* slots, unsigned epoch, double v0, double v1, double* out) {
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x, G = gridDim.x;
        if (lane == 0) {
            __threadfence();  // release: everything this CTA wrote before the barrier
            ll_publish(slots + blockIdx.x, v0, v1, epoch);
        }
        double s0 = 0.0, s1 = 0.0;
        for (int c = lane; c < G; c += 32) {
            double t0, t1;
            ll_wait(slots + c, epoch, &t0, &t1);
            s0 += t0;
            s1 += t1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        __threadfence();  // acquire: what the other CTAs wrote before they published
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
    const size_t un = std::max((size_t)NWARP * RB, (size_t)KS * R + std::max((size_t)SL * 2 * NB, (size_t)KS * R));
    return sizeof(double) * (2 * NB * R + R + 4 * NB + 16 + un);
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
    double* scal = qs + NB;          // tau, scale, alpha, beta, alpha2, scratch
    double* un = scal + 16;
    double* red = un;                // [KS][R]
    double* dots = red + KS * R;     // [SL][2 NB] (phase A), [YG <= KS][R] (phase C)
    double* redg = un;               // [NWARP][RB] (phase B only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n = a.n, Np = a.Np;
    LLSlot* const ll1 = a.ll;
    LLSlot* const ll2 = a.ll + G;
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
            ll_barrier(ll1, (unsigned)(j + 1), 0.0, 0.0, scal + 6);
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
            ll_barrier(ll1, (unsigned)(j + 1), mine, alpha_mine, scal + 6);
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
            for (int t = cta; t < tiles; t += G) {
                int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while (r * (r + 1) / 2 > t) --r;
                while ((r + 1) * (r + 2) / 2 <= t) ++r;
                const int rb = b0 + r, cb = b0 + (t - r * (r + 1) / 2);
                const bool offdiag = rb != cb;
                const double* Ab = a.A + (int64_t)rb * RB + 2 * lane + (int64_t)cb * RB * a.lda;
                // this warp's eight columns of the tile: warp + 16 m (lane m < 8 holds v of column m)
                const double vv = vcol(cb * RB + warp + 16 * (lane & 7));
                double2 lo[8], hi[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double* p = Ab + (int64_t)(warp + 16 * q) * a.lda;
                    lo[q] = __ldg(reinterpret_cast<const double2*>(p));
                    hi[q] = __ldg(reinterpret_cast<const double2*>(p + 64));
                }
                double vr0 = 0.0, vr1 = 0.0, vr2 = 0.0, vr3 = 0.0;
                if (offdiag) {
                    const int r0 = rb * RB + 2 * lane;
                    vr0 = vcol(r0);
                    vr1 = vcol(r0 + 1);
                    vr2 = vcol(r0 + 64);
                    vr3 = vcol(r0 + 65);
                }
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, tmine = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double v = __shfl_sync(0xffffffffu, vv, q);
                    a0 += lo[q].x * v;
                    a1 += lo[q].y * v;
                    a2 += hi[q].x * v;
                    a3 += hi[q].y * v;
                    if (offdiag) {  // uniform over the CTA
                        const double tq = warp_sum((lo[q].x * vr0 + lo[q].y * vr1) + (hi[q].x * vr2 + hi[q].y * vr3));
                        if (lane == q) tmine = tq;
                    }
                }
                if (offdiag && lane < 8) {
                    const int c = cb * RB + warp + 16 * lane;
                    __stcg(a.ypart + (int64_t)rb * Np + c, tmine);
                    yv += vv * tmine;
                }
                redg[warp * RB + 2 * lane] = a0;
                redg[warp * RB + 2 * lane + 1] = a1;
                redg[warp * RB + 64 + 2 * lane] = a2;
                redg[warp * RB + 64 + 2 * lane + 1] = a3;
                __syncthreads();
                if (tid < RB) {
                    double sd = 0.0;
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) sd += redg[w * RB + tid];
                    const int rr = rb * RB + tid;
                    __stcg(a.ypart + (int64_t)cb * Np + rr, sd);
                    yv += vcol(rr) * sd;
                }
                __syncthreads();
            }
            // v'(A v) partial of this CTA
            yv = warp_sum(yv);
            if (lane == 0) redg[warp] = yv;
            __syncthreads();
            if (tid == 0) {
                double sy = 0.0;
                for (int w = 0; w < NWARP; ++w) sy += redg[w];
                __stcg(a.part2 + cta, sy);
            }
        }
        grid_barrier(a.bar, round);

        // ---------------- phase C: w = tau (A v - V (W'v) - W (V'v)) - (tau/2)(w'v) v --------------------
        if (tid < jj) {
            qs[tid] = __ldcg(a.pq + tid);
            ps[tid] = __ldcg(a.pq + NB + tid);
        }
        if (warp == 0) {
            const double yv = sum_over_ctas(a.part2, 1, G, lane);
            if (lane == 0) scal[5] = yv;
        }
        __syncthreads();
        if (warp == 0) {
            double t = 0.0;
            for (int k = lane; k < jj; k += 32) t += ps[k] * qs[k];
            t = warp_sum(t);
            if (lane == 0) scal[4] = -0.5 * tau * (tau * (scal[5] - 2.0 * t));
        }
        const int cc0 = (j + 1) / RB, ncc = (n + RB - 1) / RB - cc0;  // source blocks of the partial products
        __syncthreads();
        const double alpha2 = scal[4];
        if (warp == 1) {
            // row j + 1 of V and W including the column being formed: every CTA needs it for the next column
            double c2 = 0.0, y = 0.0;
            double vn[2], wn[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                vn[h] = wn[h] = 0.0;
                if (k < jj) {
                    vn[h] = __ldcg(a.P + (j + 1) + (int64_t)k * Np);
                    wn[h] = __ldcg(a.P + (j + 1) + (int64_t)(NB + k) * Np);
                    c2 += vn[h] * ps[k] + wn[h] * qs[k];
                }
            }
            for (int cc = lane; cc < ncc; cc += 32) y += __ldcg(a.ypart + (int64_t)(cc0 + cc) * Np + (j + 1));
            c2 = warp_sum(c2);
            y = warp_sum(y);
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
                Wrow[jj] = tau * (y - c2) + alpha2;
            }
        }
        // (A v)[own rows]: the partial products of all source blocks, split over YG thread groups
        const int YG = min(THREADS / R, KS);
        double* ysum = dots;  // [YG][R]: dots is dead in this phase
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
        for (int u = warp; u < a.SL * KS; u += NWARP) {
            const int li = (u / KS) * 32 + lane, ks = u % KS;
            const int k1 = min(jj, (ks + 1) * KW);
            double c = 0.0;
            for (int k = ks * KW; k < k1; ++k) c += Vs[k * R + li] * ps[k] + Ws[k * R + li] * qs[k];
            red[ks * R + li] = c;
        }
        __syncthreads();
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
    }
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
    double *P, *Q, *ypart, *d, *e, *e2, *part1, *part2, *pq, *sc, *bnd, *mx;
    unsigned* bar;
};

int carve(aces_ctx* ctx, int64_t Np, int grid, Work* w) {
    size_t doubles = (size_t)Np * 2 * NB * 2 + (size_t)(Np / RB) * Np + 3 * (size_t)Np + (size_t)grid * P1 + grid + 2 * NB + 8 + 8 + 1024;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_eig, sizeof(double) * doubles + 256));
    double* p = (double*)ctx->d_eig.ptr;
    w->P = p; p += (size_t)Np * 2 * NB;
    w->Q = p; p += (size_t)Np * 2 * NB;
    w->ypart = p; p += (size_t)(Np / RB) * Np;
    w->d = p; p += Np;
    w->e = p; p += Np;
    w->e2 = p; p += Np;
    w->part1 = p; p += (size_t)grid * P1;
    w->part2 = p; p += grid;
    w->pq = p; p += 2 * NB;
    w->sc = p; p += 8;
    w->bnd = p; p += 8;
    w->mx = p; p += 1024;
    w->bar = (unsigned*)p;
    return ACES_OK;
}

}  // namespace

int64_t max_order(const aces_ctx* ctx) { return (int64_t)MAX_SL * 32 * ctx->sm_count; }

int syev_values(aces_ctx* ctx, int64_t n, int64_t Np, double* A, int64_t lda, double* w_dev) {
    ACES_CHECK(ctx, n >= 1 && Np % dense::DB == 0 && Np >= n && lda >= Np && lda % 2 == 0,
               "eigen::syev_values: the matrix is not padded to a multiple of 128");
    ACES_CHECK(ctx, n <= max_order(ctx), "eigen::syev_values: order %lld exceeds the supported %lld", (long long)n,
               (long long)max_order(ctx));
    const int grid = ctx->sm_count;
    ACES_CHECK(ctx, grid * (NWARP - 1) >= 2 * NB, "eigen::syev_values: too few SMs");
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
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
        const int ncols = (int)std::min<int64_t>(NB, n - j0);
        ACES_CUDA(ctx, cudaMemsetAsync(w.P, 0, sizeof(double) * (size_t)Np * 2 * NB * 2, st));  // P and Q are adjacent
        ACES_CUDA(ctx, cudaMemsetAsync(w.bar, 0, sizeof(unsigned), st));
        PanelArgs pa;
        pa.A = A; pa.lda = lda; pa.n = (int)n; pa.Np = (int)Np; pa.j0 = (int)j0; pa.ncols = ncols;
        pa.P = w.P; pa.Q = w.Q; pa.d = w.d; pa.e = w.e; pa.ypart = w.ypart; pa.part1 = w.part1; pa.part2 = w.part2;
        pa.pq = w.pq; pa.bar = w.bar; pa.SL = SL;
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
