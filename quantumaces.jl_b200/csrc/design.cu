// design.cu -- the design-resident fit path: everything estimate_gate_noise needs from the GPU
// with the Design uploaded once.
//
// Reference functions replaced (paths relative to the QuantumACES.jl root):
//   calc_covariance / calc_covariance_log          src/merit.jl:206-291, 356-365   (K2)
//   sparse_covariance_inv_factor                   src/merit.jl:378-427            (K3)
//   estimate_gate_eigenvalues (OLS / WLS / GLS)    src/estimate.jl:492-524         (K4/K5)
//   calc_gls / wls / ols_covariance, nrmse_moments src/merit.jl:556-629, 672-682   (K6)
// as they are called from estimate_gate_noise (src/estimate.jl:703-743, 820-831).
//
// Layout.  The M circuit eigenvalues are grouped in T tuple blocks of m_t rows.  Every block gets
// a dense column-major tile of mp_t x mp_t doubles (mp_t = m_t rounded up to 128) for its
// log-covariance, which is factorised in place (L_t) and inverted (X_t = inv(L_t), the
// reference's F block) by batched launches over all tuples at once.  The "padded row space" of
// dimension Mp = sum mp_t stacks the blocks; clipped rows (src/estimate.jl:435-460) are
// compacted inside their block through row_map and the freed rows become identity padding, so
// the launch sequence does not depend on what was clipped.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "dense.cuh"

using dense::DB;
using dense::DiagDesc;
using dense::GemmDesc;
using dense::pad_up;

struct aces_design {
    int T = 0;
    int64_t M = 0, N = 0, Mp = 0, Np = 0, C = 0, nnz_cov = 0, nnzA = 0;
    int max_nb = 0;
    double times_factor = 0.0;
    std::vector<int64_t> m, mp, row0, prow0, tile_off, dinv_off, key0;
    std::vector<int64_t> h_cov_colptr, h_cov_rowval;  // covariance pattern (0-based), host copy
    std::vector<int32_t> h_blk_of_row;
    // --- device: design data ---
    int64_t *a_colptr = nullptr, *a_rowptr = nullptr;  // CSC / CSR of A
    int32_t *a_rowval = nullptr, *a_colidx = nullptr;
    double *a_nzval = nullptr, *a_vals = nullptr;
    int64_t *cov_colptr = nullptr, *cov_rowval = nullptr;
    int32_t *cov_src = nullptr;    // per stored entry: -1 = variance, else key index
    int32_t *blk_of_row = nullptr;
    double *dfac = nullptr;        // per row: (1 / w_t) * (X_t / c_pp)
    double *kfac = nullptr;        // per key: (1 / w_t) * ((X_t * c_pq) / (c_pp * c_qq))
    int64_t *key_p = nullptr, *key_q = nullptr;  // global rows of each key
    int64_t *d_row0 = nullptr, *d_prow0 = nullptr, *d_tile_off = nullptr, *d_mp = nullptr;
    // --- device: batched descriptor tables (built once) ---
    DiagDesc* dd = nullptr;        // [max_nb][T]
    GemmDesc *g_panel = nullptr, *g_trail = nullptr, *g_tri1 = nullptr, *g_tri2 = nullptr;  // [max_nb][T]
    // --- device: workspaces (allocated on first use) ---
    double *tilesL = nullptr, *tilesX = nullptr, *dinvT = nullptr;
    double *At = nullptr, *G = nullptr, *Xn = nullptr, *Cn = nullptr, *dinvG = nullptr;
    double *vec = nullptr;         // vector workspace
    double *mom = nullptr;         // partial sums of the trace moments
    double *covval = nullptr;      // [nnz_cov]
    int32_t *row_map = nullptr;    // [M] compacted local row or -1
    double *lam_buf = nullptr;     // [M + C + 2] eigenvalue estimates of the current call (owned: the staged
                                   // fit API reads them again in later calls)
    int64_t *kept_dev = nullptr;   // [2 T + 2] kept rows per tuple | export offsets
    int32_t *flags = nullptr;      // [T + 1] not-PD flags (blocks, then the Gram matrix)
    int fallback_blocks = 0;       // blocks that needed the not-PD fallback in the last call
    std::vector<std::pair<std::string, float>> phase_ms;  // phases of the last fit (timing on)
    // --- column windows of the tuple blocks (GLS Gram): block t only touches the columns cols_t ---
    bool windowed = false;         // the windowed Gram does at most half the flops of the dense one
    int max_ct = 0;                // max cp_t / 128
    int64_t n_wcols = 0;           // sum of c_t
    std::vector<int64_t> ct, cpt, cols_off, atc_off, gt_off;
    int32_t *w_blk = nullptr;      // [n_wcols] block of each compact column
    int64_t *w_ptr = nullptr;      // [n_wcols + 1] entries of each compact column
    int32_t *w_row = nullptr;      // global row of each entry
    double *w_val = nullptr;
    int64_t *w_cols = nullptr;     // [n_wcols] global column of each compact column
    int64_t *d_cols_off = nullptr, *d_atc_off = nullptr;
    double *Atc = nullptr, *Gt = nullptr;
    GemmDesc* g_gram = nullptr;    // [T]
    // --- workspaces of aces_design_merit_grad (merit_grad.cuh), allocated on first use ---
    double *W3 = nullptr, *W4 = nullptr;   // two more Np x Np matrices
    double *mg_rows = nullptr, *mg_cu = nullptr;
    int64_t* mg_win = nullptr;     // ct | cpt | cols_off | gt_off
    // --- tuple shard ("one large Gram"): this design works on the tuples that are on;
    //     the host sums the partial Gram matrix and right-hand sides over the shards ---
    std::vector<uint8_t> tuple_on;  // [T], all 1 by default
    uint8_t* d_tuple_on = nullptr;
    bool sharded = false;
    // --- state of a fit split at its reduction points (aces_design_fit_begin .. _end) ---
    int fs_stage = 0, fs_ls_type = 0;
    int est_ls_type = -1;          // estimator whose factors / weights / clipping the workspaces hold (-1: none)
    bool fs_clipped = false;
    double* fs_lam = nullptr;
    std::vector<void*> owned;
};

namespace {

// ---- small device helpers -------------------------------------------------------------------------

template <typename Tp>
int dev_alloc(aces_ctx* ctx, aces_design* d, Tp** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, sizeof(Tp) * std::max<size_t>(count, 1));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return aces_fail(ctx, ACES_E_NOMEM, "design: cudaMalloc of %zu bytes failed: %s", sizeof(Tp) * count,
                         cudaGetErrorString(e));
    }
    d->owned.push_back(q);
    *p = (Tp*)q;
    return ACES_OK;
}
template <typename Tp>
int dev_upload(aces_ctx* ctx, aces_design* d, Tp** p, const std::vector<Tp>& h) {
    ACES_TRY(dev_alloc(ctx, d, p, h.size()));
    if (!h.empty()) ACES_CUDA(ctx, cudaMemcpy(*p, h.data(), sizeof(Tp) * h.size(), cudaMemcpyHostToDevice));
    return ACES_OK;
}

// ---- K2: covariance values ------------------------------------------------------------------------
// One thread per stored entry of the block-diagonal covariance (src/merit.jl:245-279).  The
// arithmetic is written with explicit round-to-nearest intrinsics so that no multiply-add is
// contracted: the values are the ones the reference's Float64 expressions produce.
__global__ void cov_values_kernel(int64_t nnz, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                  const int32_t* __restrict__ src, int64_t M, const double* __restrict__ lam,
                                  const double* __restrict__ lam_cov, const double* __restrict__ dfac,
                                  const double* __restrict__ kfac, const int64_t* __restrict__ key_p,
                                  const int64_t* __restrict__ key_q, double epsilon, double times_factor,
                                  int weight_time, int log_form, double* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rowval[e];
        // column of entry e: last c with colptr[c] <= e
        int64_t lo = 0, hi = M;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (colptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int64_t c = lo;
        const int32_t k = src[e];
        double v;
        if (k < 0) {
            const double l = lam[r];
            v = __dmul_rn(dfac[r], fmax(__dsub_rn(1.0, __dmul_rn(l, l)), epsilon));
        } else {
            v = __dmul_rn(kfac[k], __dsub_rn(lam_cov[k], __dmul_rn(lam[key_p[k]], lam[key_q[k]])));
        }
        if (weight_time) v = __dmul_rn(times_factor, v);
        if (log_form) {
            // sparse(Symmetric(D^-1 * cov * D^-1)): the upper triangle, mirrored (src/merit.jl:361-363)
            const int64_t i = r < c ? r : c, j = r < c ? c : r;
            v = __dmul_rn(__dmul_rn(__ddiv_rn(1.0, lam[i]), v), __ddiv_rn(1.0, lam[j]));
        }
        out[e] = v;
    }
}

// ---- K3: scatter the (clipped) covariance into dense tiles ---------------------------------------
__global__ void tiles_init_kernel(int T, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                  double* __restrict__ tiles, int identity) {
    const int t = blockIdx.y;
    if (t >= T) return;
    const int64_t n = mp[t], total = n * n;
    double* tile = tiles + tile_off[t];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        tile[e] = (identity && (e % n) == (e / n)) ? 1.0 : 0.0;
}
// One thread per column c of the M x M block-diagonal sparse matrix; entries outside the diagonal
// blocks are ignored (the reference slices the blocks out, src/merit.jl:393).  row_map == NULL:
// nothing is clipped.  Rows past the kept ones get a unit diagonal.
__global__ void tiles_scatter_kernel(int64_t M, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                     const double* __restrict__ val, const int32_t* __restrict__ blk_of_row,
                                     const int64_t* __restrict__ row0, const int64_t* __restrict__ tile_off,
                                     const int64_t* __restrict__ mp, const int32_t* __restrict__ row_map,
                                     double* __restrict__ tiles) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < M; c += (int64_t)gridDim.x * blockDim.x) {
        const int t = blk_of_row[c];
        const int64_t lc = row_map ? row_map[c] : c - row0[t];
        if (lc < 0) continue;
        double* tile = tiles + tile_off[t];
        const int64_t n = mp[t];
        for (int64_t e = colptr[c]; e < colptr[c + 1]; ++e) {
            const int64_t r = rowval[e];
            if (blk_of_row[r] != t) continue;
            const int64_t lr = row_map ? row_map[r] : r - row0[t];
            if (lr < 0) continue;
            tile[lr + lc * n] = val[e];
        }
    }
}
// Rebuilds the tile of ONE block from the pattern values for the not-positive-definite fallback
// (src/merit.jl:403-424): elementwise absolute value and `shift` added to the kept diagonal.
__global__ void tile_rebuild_kernel(int64_t c0, int64_t c1, const int64_t* __restrict__ colptr,
                                    const int64_t* __restrict__ rowval, const double* __restrict__ val,
                                    const int32_t* __restrict__ blk_of_row, int t, int64_t row0_t,
                                    const int32_t* __restrict__ row_map, double shift, double* __restrict__ tile,
                                    int64_t n) {
    for (int64_t c = c0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < c1; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lc = row_map ? row_map[c] : c - row0_t;
        if (lc < 0) continue;
        for (int64_t e = colptr[c]; e < colptr[c + 1]; ++e) {
            const int64_t r = rowval[e];
            if (blk_of_row[r] != t) continue;
            const int64_t lr = row_map ? row_map[r] : r - row0_t;
            if (lr < 0) continue;
            tile[lr + lc * n] = fabs(val[e]) + (lr == lc ? shift : 0.0);
        }
    }
}
// unit diagonal on rows kept[t] .. mp[t]-1 of every tile
__global__ void tiles_pad_kernel(int T, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                 const int64_t* __restrict__ kept, double* __restrict__ tiles) {
    const int t = blockIdx.y;
    if (t >= T) return;
    const int64_t n = mp[t];
    double* tile = tiles + tile_off[t];
    for (int64_t i = kept[t] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        tile[i + i * n] = 1.0;
}
// dense lower-triangular tiles -> caller layout (m_t x m_t col-major blocks, concatenated)
__global__ void tiles_export_kernel(int T, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                    const int64_t* __restrict__ kept, const int64_t* __restrict__ out_off,
                                    const double* __restrict__ tiles, double* __restrict__ out) {
    const int t = blockIdx.y;
    if (t >= T) return;
    const int64_t n = mp[t], k = kept[t], total = k * k;
    const double* tile = tiles + tile_off[t];
    double* o = out + out_off[t];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % k, j = e / k;
        o[e] = (i >= j) ? tile[i + j * n] : 0.0;
    }
}

// ---- operators on the padded row space --------------------------------------------------------------
// u[prow] = b_r - sum_j A[r, j] x[j] for kept rows (b_r = -log y_r), zero elsewhere.  x == NULL: u = b.
__global__ void residual_kernel(int64_t M, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const double* __restrict__ vals, const int32_t* __restrict__ blk_of_row,
                                const int64_t* __restrict__ row0, const int64_t* __restrict__ prow0,
                                const int32_t* __restrict__ row_map, const double* __restrict__ y,
                                const double* __restrict__ x, double* __restrict__ u) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
        const int t = blk_of_row[r];
        const int64_t lr = row_map ? row_map[r] : r - row0[t];
        if (lr < 0) continue;
        double s = -log(y[r]);
        if (x)
            for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) s -= vals[e] * x[colidx[e]];
        u[prow0[t] + lr] = s;
    }
}
// g[j] = sum_r A[r, j] v[prow(r)]
__global__ void at_times_kernel(int64_t N, const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                const double* __restrict__ nzval, const int32_t* __restrict__ blk_of_row,
                                const int64_t* __restrict__ row0, const int64_t* __restrict__ prow0,
                                const int32_t* __restrict__ row_map, const double* __restrict__ v,
                                double* __restrict__ g) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < N; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int64_t e = colptr[j]; e < colptr[j + 1]; ++e) {
            const int64_t r = rowval[e];
            const int t = blk_of_row[r];
            const int64_t lr = row_map ? row_map[r] : r - row0[t];
            if (lr >= 0) s += nzval[e] * v[prow0[t] + lr];
        }
        g[j] = s;
    }
}
// out = F u with F = blockdiag(X_t) (lower triangular tiles).  A block of 256 threads owns 32 rows
// of one tile: thread (r, q) sums the columns c = q, q + 8, ... <= row (for a fixed column the 32
// rows are consecutive in memory), the eight partial sums of a row are added in a fixed order
// (deterministic: concurrent fits stay bit-identical).
__global__ void __launch_bounds__(256) apply_f_kernel(int T, const int64_t* __restrict__ prow0,
                                                      const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                                      const double* __restrict__ tiles, const uint8_t* __restrict__ on,
                                                      const double* __restrict__ u, double* __restrict__ out) {
    __shared__ double part[8][33];
    const int t = blockIdx.y;
    if (t >= T) return;
    const int64_t n = mp[t];
    const double* X = tiles + tile_off[t];
    const double* ut = u + prow0[t];
    const bool act = on[t] != 0;  // a tuple of another shard contributes nothing here
    const int r = threadIdx.x & 31, q = threadIdx.x >> 5;
    for (int64_t i0 = 32 * (int64_t)blockIdx.x; i0 < n; i0 += 32 * (int64_t)gridDim.x) {
        const int64_t i = i0 + r;
        double s = 0.0;
        if (act) {
            // four independent chains hide the load latency
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int64_t c = q;
            for (; c + 24 <= i; c += 32) {
                s0 += X[i + c * n] * ut[c];
                s1 += X[i + (c + 8) * n] * ut[c + 8];
                s2 += X[i + (c + 16) * n] * ut[c + 16];
                s3 += X[i + (c + 24) * n] * ut[c + 24];
            }
            for (; c <= i; c += 8) s0 += X[i + c * n] * ut[c];
            s = (s0 + s1) + (s2 + s3);
        }
        part[q][r] = s;
        __syncthreads();
        if (q == 0) {
            double v = part[0][r];
#pragma unroll
            for (int k = 1; k < 8; ++k) v += part[k][r];
            out[prow0[t] + i] = v;
        }
        __syncthreads();
    }
}
// out = F' v: one warp per padded column
__global__ void apply_ft_kernel(int T, const int64_t* __restrict__ prow0, const int64_t* __restrict__ tile_off,
                                const int64_t* __restrict__ mp, const double* __restrict__ tiles,
                                const uint8_t* __restrict__ on, const double* __restrict__ v, double* __restrict__ out) {
    const int t = blockIdx.y;
    if (t >= T) return;
    const int64_t n = mp[t];
    const double* X = tiles + tile_off[t];
    const double* vt = v + prow0[t];
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool act = on[t] != 0;
    for (int64_t c = warp; c < n; c += nwarps) {
        double s = 0.0;
        if (act)
            for (int64_t i = c + lane; i < n; i += 32) s += X[i + c * n] * vt[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[prow0[t] + c] = s;
    }
}
// diagonal F (WLS): f[prow] = cov_log[r, r]^(-1/2) for kept rows (src/estimate.jl:729-731), 0 elsewhere
__global__ void diag_factor_kernel(int64_t M, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                   const double* __restrict__ val, const int32_t* __restrict__ blk_of_row,
                                   const int64_t* __restrict__ row0, const int64_t* __restrict__ prow0,
                                   const int32_t* __restrict__ row_map, const uint8_t* __restrict__ on, int unit,
                                   double* __restrict__ f) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
        const int t = blk_of_row[r];
        const int64_t lr = row_map ? row_map[r] : r - row0[t];
        if (lr < 0) continue;
        double d = 1.0;
        if (!on[t])
            d = 0.0;  // rows of another shard's tuples: weight 0 (f is zeroed before, kept explicit)
        else if (!unit)
            for (int64_t e = colptr[r]; e < colptr[r + 1]; ++e)
                if (rowval[e] == r) d = pow(val[e], -0.5);
        f[prow0[t] + lr] = d;
    }
}
__global__ void mul_vec_kernel(int64_t n, const double* __restrict__ f, const double* __restrict__ u, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = f[i] * u[i];
}
__global__ void axpy_kernel(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] += a * x[i];
}
__global__ void exp_neg_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ g, int constrain) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (constrain && v < 0.0) v = 0.0;  // src/estimate.jl:504-506
        g[i] = exp(-v);
    }
}

// out[0] = max |x_i|, out[1] = max |d_i| (one CTA): the convergence test of the refinement
__global__ void max_abs_pair_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ dlt,
                                    double* __restrict__ out) {
    __shared__ double sx[256], sd[256];
    double mx = 0.0, md = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        mx = fmax(mx, fabs(x[i]));
        md = fmax(md, fabs(dlt[i]));
    }
    sx[threadIdx.x] = mx;
    sd[threadIdx.x] = md;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sx[threadIdx.x] = fmax(sx[threadIdx.x], sx[threadIdx.x + o]);
            sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sx[0];
        out[1] = sd[0];
    }
}

// ---- whitened design matrix (GLS) --------------------------------------------------------------------
// At[:, j] = sum over the entries (r, a) of column j of A of a * X_t[:, lr]  (one CTA per column j;
// row i of a block is always handled by thread i % blockDim, so no two threads touch one entry).
__global__ void whiten_tiles_kernel(int64_t N, const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval,
                                    const double* __restrict__ nzval, const int32_t* __restrict__ blk_of_row,
                                    const int64_t* __restrict__ row0, const int64_t* __restrict__ prow0,
                                    const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                    const int32_t* __restrict__ row_map, const double* __restrict__ tiles,
                                    const uint8_t* __restrict__ on, double* __restrict__ At, int64_t ld) {
    const int64_t j = blockIdx.x;
    if (j >= N) return;
    double* col = At + j * ld;
    const int bd = blockDim.x, tid = threadIdx.x;
    for (int64_t e = colptr[j]; e < colptr[j + 1]; ++e) {
        const int64_t r = rowval[e];
        const int t = blk_of_row[r];
        const int64_t lr = row_map ? row_map[r] : r - row0[t];
        if (lr < 0 || !on[t]) continue;
        const double a = nzval[e];
        const int64_t n = mp[t];
        const double* X = tiles + tile_off[t] + lr * n;  // column lr of X_t: nonzero for rows >= lr
        double* out = col + prow0[t];
        int64_t i = lr - (lr % bd) + tid;
        if (i < lr) i += bd;
        for (; i < n; i += bd) out[i] += a * X[i];
    }
}

// Compact whitened blocks for the windowed Gram: compact column jc of block t (global column
// w_cols[jc]) = sum over its entries (r, a) of a * X_t[:, lr].  One CTA per compact column.
__global__ void whiten_compact_kernel(int64_t n_wcols, const int32_t* __restrict__ w_blk,
                                      const int64_t* __restrict__ w_ptr, const int32_t* __restrict__ w_row,
                                      const double* __restrict__ w_val, const int64_t* __restrict__ cols_off,
                                      const int64_t* __restrict__ atc_off, const int64_t* __restrict__ row0,
                                      const int64_t* __restrict__ tile_off, const int64_t* __restrict__ mp,
                                      const int32_t* __restrict__ row_map, const double* __restrict__ tiles,
                                      const uint8_t* __restrict__ on, double* __restrict__ Atc) {
    const int64_t jc = blockIdx.x;
    if (jc >= n_wcols) return;
    const int t = w_blk[jc];
    if (!on[t]) return;
    const int64_t n = mp[t];
    double* out = Atc + atc_off[t] + (jc - cols_off[t]) * n;
    const int bd = blockDim.x, tid = threadIdx.x;
    for (int64_t e = w_ptr[jc]; e < w_ptr[jc + 1]; ++e) {
        const int64_t r = w_row[e];
        const int64_t lr = row_map ? row_map[r] : r - row0[t];
        if (lr < 0) continue;
        const double a = w_val[e];
        const double* X = tiles + tile_off[t] + lr * n;
        int64_t i = lr - (lr % bd) + tid;
        if (i < lr) i += bd;
        for (; i < n; i += bd) out[i] += a * X[i];
    }
}
// G[cols[i], cols[j]] += Gt[i, j] for i >= j (cols ascending, so the target stays in the lower triangle)
__global__ void gram_scatter_kernel(int64_t c, int64_t cp, const int64_t* __restrict__ cols,
                                    const double* __restrict__ Gt, double* __restrict__ G, int64_t ldg) {
    const int64_t j = blockIdx.y;
    if (j >= c) return;
    const int64_t gj = cols[j];
    for (int64_t i = j + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < c; i += (int64_t)gridDim.x * blockDim.x)
        G[cols[i] + gj * ldg] += Gt[i + j * cp];
}

// ---- sparse sandwich C = A' diag(w) S diag(w) A (lower triangle of C only) --------------------------
// S is an M x M sparse matrix in CSC (the covariance pattern, or NULL for the identity); w is
// given on the padded row space (0 for clipped rows).  One thread per column b of C.
__global__ void sandwich_kernel(int64_t N, const int64_t* __restrict__ a_colptr, const int32_t* __restrict__ a_rowval,
                                const double* __restrict__ a_nzval, const int64_t* __restrict__ a_rowptr,
                                const int32_t* __restrict__ a_colidx, const double* __restrict__ a_vals,
                                const int64_t* __restrict__ s_colptr, const int64_t* __restrict__ s_rowval,
                                const double* __restrict__ s_val, const int32_t* __restrict__ blk_of_row,
                                const int64_t* __restrict__ row0, const int64_t* __restrict__ prow0,
                                const int32_t* __restrict__ row_map, const double* __restrict__ w,
                                double* __restrict__ Cm, int64_t ldc) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < N; b += (int64_t)gridDim.x * blockDim.x) {
        double* col = Cm + b * ldc;
        for (int64_t e = a_colptr[b]; e < a_colptr[b + 1]; ++e) {
            const int64_t j = a_rowval[e];
            const int tj = blk_of_row[j];
            const int64_t lj = row_map ? row_map[j] : j - row0[tj];
            if (lj < 0) continue;
            const double ajb = a_nzval[e] * w[prow0[tj] + lj];
            const int64_t f0 = s_colptr ? s_colptr[j] : 0, f1 = s_colptr ? s_colptr[j + 1] : 1;
            for (int64_t f = f0; f < f1; ++f) {
                const int64_t i = s_colptr ? s_rowval[f] : j;
                const int ti = blk_of_row[i];
                const int64_t li = row_map ? row_map[i] : i - row0[ti];
                if (li < 0) continue;
                const double sij = (s_colptr ? s_val[f] : 1.0) * w[prow0[ti] + li] * ajb;
                for (int64_t q = a_rowptr[i]; q < a_rowptr[i + 1]; ++q) {
                    const int64_t a = a_colidx[q];
                    if (a >= b) col[a] += a_vals[q] * sij;
                }
            }
        }
    }
}

// Per-gate diagonal blocks of the precision matrix D^-1 G D^-1 (src/merit.jl:754-770) from the lower
// triangle of the Gram matrix: one CTA per gate, block (a, b) = (1 / g_a) * G[i_a, i_b] * (1 / g_b).
__global__ void precision_blocks_kernel(int G, const int64_t* __restrict__ gptr, const int64_t* __restrict__ boff,
                                        const int32_t* __restrict__ gidx, const double* __restrict__ Gm, int64_t ldg,
                                        const double* __restrict__ scale, double* __restrict__ out) {
    const int g = blockIdx.x;
    if (g >= G) return;
    const int64_t e0 = gptr[g];
    const int n = (int)(gptr[g + 1] - e0);
    double* o = out + boff[g];
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int a = e / n, b = e % n;
        const int64_t ia = gidx[e0 + a], ib = gidx[e0 + b];
        const double v = Gm[(ia > ib ? ia : ib) + (ia > ib ? ib : ia) * ldg];
        const double da = scale ? __ddiv_rn(1.0, scale[ia]) : 1.0, db = scale ? __ddiv_rn(1.0, scale[ib]) : 1.0;
        o[e] = __dmul_rn(__dmul_rn(da, v), db);
    }
}

__global__ void scale_sym_kernel(int64_t n, double* S, int64_t lds, const double* d) {
    const int64_t total = n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % n, j = e / n;
        S[i + j * lds] = S[i + j * lds] * (d[i] * d[j]);  // (di*dj) commutes: the result stays exactly symmetric
    }
}
// tr(S) and tr(S^2) = sum_ij S_ij^2 of a symmetric n x n matrix (src/merit.jl:674-675)
__global__ void moments_kernel(int64_t n, const double* __restrict__ S, int64_t lds, double* __restrict__ out) {
    double tr = 0.0, sq = 0.0;
    const int64_t total = n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % n, j = e / n;
        const double v = S[i + j * lds];
        sq += v * v;
        if (i == j) tr += v;
    }
    __shared__ double s_tr[256], s_sq[256];
    s_tr[threadIdx.x] = tr;
    s_sq[threadIdx.x] = sq;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_tr[threadIdx.x] += s_tr[threadIdx.x + o];
            s_sq[threadIdx.x] += s_sq[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // per-CTA partial sums: the final order of additions is fixed
        out[2 + 2 * blockIdx.x] = s_tr[0];
        out[3 + 2 * blockIdx.x] = s_sq[0];
    }
}
__global__ void moments_final_kernel(int nparts, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tr = 0.0, sq = 0.0;
        for (int i = 0; i < nparts; ++i) {
            tr += out[2 + 2 * i];
            sq += out[3 + 2 * i];
        }
        out[0] = tr;
        out[1] = sq;
    }
}


// CUDA-event timing of the phases of a fit: recorded when timing is on (aces_set_timing) and read
// back with aces_design_phase_ms; ACES_FIT_TRACE=1 also prints them to stderr.
struct PhaseTimer {
    aces_ctx* ctx;
    aces_design* d;
    bool on, print;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    PhaseTimer(aces_ctx* c, aces_design* dd)
        : ctx(c), d(dd), on(false), print(getenv("ACES_FIT_TRACE") != nullptr) {
        on = print || c->timing;
        mark("start");
    }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        ev.push_back(e);
        names.push_back(name);
    }
    ~PhaseTimer() {
        if (!on) return;
        cudaEventSynchronize(ev.back());
        d->phase_ms.clear();
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            d->phase_ms.push_back({names[i], ms});
            if (print) fprintf(stderr, "[aces fit] %-18s %8.3f ms\n", names[i], ms);
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

inline unsigned grid_for(int64_t n, int threads, int64_t cap = 148 * 16) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, cap));
}

// ---- host orchestration -------------------------------------------------------------------------------

int ensure_workspaces(aces_ctx* ctx, aces_design* d, bool need_tiles, bool need_at, bool need_inverse, bool need_c) {
    const size_t tiles = (size_t)(d->tile_off.back());
    if (need_tiles && !d->tilesL) {
        ACES_TRY(dev_alloc(ctx, d, &d->tilesL, tiles));
        ACES_TRY(dev_alloc(ctx, d, &d->tilesX, tiles));
        ACES_TRY(dev_alloc(ctx, d, &d->dinvT, (size_t)d->dinv_off.back()));
    }
    if (need_at && d->windowed && !d->Atc) {
        ACES_TRY(dev_alloc(ctx, d, &d->Atc, (size_t)d->atc_off.back()));
        ACES_TRY(dev_alloc(ctx, d, &d->Gt, (size_t)d->gt_off.back()));
    }
    if (need_at && d->windowed && !d->g_gram) {
        std::vector<GemmDesc> gg((size_t)d->T);
        for (int t = 0; t < d->T; ++t) {
            const double* A = d->Atc + d->atc_off[t];
            const int tiles = (int)(d->cpt[t] / DB);
            gg[(size_t)t] = GemmDesc{A, A, d->Gt + d->gt_off[t], d->mp[t], d->mp[t], d->cpt[t], (int32_t)d->mp[t], tiles, tiles,
                                     1, 1.0, 0.0};
            if (!d->tuple_on[(size_t)t]) gg[(size_t)t] = GemmDesc{};  // k = 0: inactive entry
        }
        ACES_TRY(dev_upload(ctx, d, &d->g_gram, gg));
    }
    if (need_at && !d->windowed && !d->At) ACES_TRY(dev_alloc(ctx, d, &d->At, (size_t)d->Mp * d->Np));
    if (!d->G) {
        ACES_TRY(dev_alloc(ctx, d, &d->G, (size_t)d->Np * d->Np));
        ACES_TRY(dev_alloc(ctx, d, &d->dinvG, (size_t)d->Np * DB));
        ACES_TRY(dev_alloc(ctx, d, &d->vec, (size_t)(8 * std::max(d->Mp, d->Np) + 64)));
        ACES_TRY(dev_alloc(ctx, d, &d->covval, (size_t)d->nnz_cov));
        ACES_TRY(dev_alloc(ctx, d, &d->mom, (size_t)(2 + 2 * 148 * 4)));
        ACES_TRY(dev_alloc(ctx, d, &d->row_map, (size_t)d->M));
        ACES_TRY(dev_alloc(ctx, d, &d->flags, (size_t)d->T + 2));
        ACES_TRY(dev_alloc(ctx, d, &d->lam_buf, (size_t)(d->M + d->C + 2)));
        ACES_TRY(dev_alloc(ctx, d, &d->kept_dev, (size_t)(2 * d->T + 2)));
    }
    if (need_inverse && !d->Xn) ACES_TRY(dev_alloc(ctx, d, &d->Xn, (size_t)d->Np * d->Np));
    if (need_c && !d->Cn) ACES_TRY(dev_alloc(ctx, d, &d->Cn, (size_t)d->Np * d->Np));
    return ACES_OK;
}

// Cholesky factors L_t (tilesL, in place) and their inverses X_t (tilesX) of every tile, batched.
int factor_tiles(aces_ctx* ctx, aces_design* d) {
    const int T = d->T;
    for (int k = 0; k < d->max_nb; ++k) {
        ACES_TRY(dense::diag_blocks(ctx, d->dd + (size_t)k * T, T, 0.0));
        if (k + 1 < d->max_nb) {
            ACES_TRY(dense::gemm_batched(ctx, dense::GEMM_NT, d->g_panel + (size_t)k * T, T, d->max_nb - k - 1, 1));
            ACES_TRY(dense::gemm_batched(ctx, dense::GEMM_NT, d->g_trail + (size_t)k * T, T, d->max_nb - k - 1,
                                         d->max_nb - k - 1));
        }
    }
    // X_t = inv(L_t): forward substitution on the identity, block row by block row
    {
        dim3 grid(148, (unsigned)T);
        tiles_init_kernel<<<grid, 256, 0, ctx->stream>>>(T, d->d_tile_off, d->d_mp, d->tilesX, 1);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    for (int k = 0; k < d->max_nb; ++k) {
        ACES_TRY(dense::gemm_batched(ctx, dense::GEMM_NN, d->g_tri1 + (size_t)k * T, T, 1, k + 1));
        if (k + 1 < d->max_nb)
            ACES_TRY(dense::gemm_batched(ctx, dense::GEMM_NN, d->g_tri2 + (size_t)k * T, T, d->max_nb - k - 1, k + 1));
    }
    return ACES_OK;
}

// Uploads row_map / kept counts for a kept-row list (NULL = all rows) and returns the device kept array.
int upload_clipping(aces_ctx* ctx, aces_design* d, const int64_t* kept_rows, int64_t n_kept, int index_base,
                    std::vector<int64_t>* kept_count, bool* clipped) {
    kept_count->assign(d->m.begin(), d->m.end());
    *clipped = false;
    if (!kept_rows) return ACES_OK;
    std::vector<int32_t> map((size_t)d->M, -1);
    std::vector<int64_t> cnt((size_t)d->T, 0);
    int64_t prev = -1;
    for (int64_t i = 0; i < n_kept; ++i) {
        const int64_t r = kept_rows[i] - index_base;
        ACES_CHECK(ctx, r > prev && r < d->M, "design: kept rows must be sorted, unique and inside 0..M-1");
        prev = r;
        const int t = d->h_blk_of_row[(size_t)r];
        map[(size_t)r] = (int32_t)cnt[(size_t)t]++;
    }
    *clipped = n_kept != d->M;
    if (!*clipped) return ACES_OK;
    *kept_count = cnt;
    ACES_CUDA(ctx, cudaMemcpyAsync(d->row_map, map.data(), sizeof(int32_t) * (size_t)d->M, cudaMemcpyHostToDevice,
                                   ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `map` dies at return
    return ACES_OK;
}

int launch_cov_values(aces_ctx* ctx, aces_design* d, const double* lam_dev, const double* lamcov_dev, double epsilon,
                      int weight_time, int log_form, double* out_dev) {
    if (d->nnz_cov == 0) return ACES_OK;
    cov_values_kernel<<<grid_for(d->nnz_cov, 256), 256, 0, ctx->stream>>>(
        d->nnz_cov, d->cov_colptr, d->cov_rowval, d->cov_src, d->M, lam_dev, lamcov_dev, d->dfac, d->kfac, d->key_p,
        d->key_q, epsilon, d->times_factor, weight_time, log_form, out_dev);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

// Fills tilesL with the (clipped) log-covariance blocks held in `val` (pattern order), factorises.
int build_and_factor_tiles(aces_ctx* ctx, aces_design* d, const double* val_dev, const int32_t* row_map,
                           const std::vector<int64_t>& kept_count) {
    const int T = d->T;
    dim3 grid(148, (unsigned)T);
    tiles_init_kernel<<<grid, 256, 0, ctx->stream>>>(T, d->d_tile_off, d->d_mp, d->tilesL, 0);
    tiles_scatter_kernel<<<grid_for(d->M, 128), 128, 0, ctx->stream>>>(d->M, d->cov_colptr, d->cov_rowval, val_dev,
                                                                      d->blk_of_row, d->d_row0, d->d_tile_off, d->d_mp,
                                                                      row_map, d->tilesL);
    int64_t* kept_dev = d->kept_dev;
    ACES_CUDA(ctx, cudaMemcpyAsync(kept_dev, kept_count.data(), sizeof(int64_t) * (size_t)T, cudaMemcpyHostToDevice,
                                   ctx->stream));
    dim3 gpad(4, (unsigned)T);
    tiles_pad_kernel<<<gpad, 128, 0, ctx->stream>>>(T, d->d_tile_off, d->d_mp, kept_dev, d->tilesL);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // kept_count may be a temporary of the caller
    return factor_tiles(ctx, d);
}

struct FitVectors {
    double *u, *v, *w, *g, *x, *t1, *t2, *f;
};
FitVectors fit_vectors(aces_design* d) {
    const int64_t n = std::max(d->Mp, d->Np);
    FitVectors fv;
    fv.u = d->vec; fv.v = d->vec + n; fv.w = d->vec + 2 * n; fv.g = d->vec + 3 * n; fv.x = d->vec + 4 * n;
    fv.t1 = d->vec + 5 * n; fv.t2 = d->vec + 6 * n; fv.f = d->vec + 7 * n;
    return fv;
}

// out (padded row space) = F u  /  out = F' v for the chosen estimator
int apply_F(aces_ctx* ctx, aces_design* d, int ls_type, const FitVectors& fv, const double* u, double* out, bool transpose) {
    if (ls_type == 2) {
        dim3 grid(16, (unsigned)d->T);
        if (!transpose)
            apply_f_kernel<<<dim3(48, (unsigned)d->T), 256, 0, ctx->stream>>>(d->T, d->d_prow0, d->d_tile_off, d->d_mp, d->tilesX, d->d_tuple_on, u, out);
        else
            apply_ft_kernel<<<grid, 256, 0, ctx->stream>>>(d->T, d->d_prow0, d->d_tile_off, d->d_mp, d->tilesX, d->d_tuple_on, u, out);
    } else {
        mul_vec_kernel<<<grid_for(d->Mp, 256), 256, 0, ctx->stream>>>(d->Mp, fv.f, u, out);
    }
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

// G = (F A)' (F A) (lower triangle) for the chosen estimator; fv.f must hold the diagonal factor
// for OLS / WLS, tilesX the block factors for GLS.
int build_gram(aces_ctx* ctx, aces_design* d, int ls_type, const FitVectors& fv, const int32_t* row_map) {
    const int64_t Np = d->Np, Mp = d->Mp;
    if (ls_type == 2 && d->windowed) {
        // G = sum_t P_t' (At_t' At_t) P_t with At_t = X_t A_t restricted to the columns block t touches
        ACES_CUDA(ctx, cudaMemsetAsync(d->Atc, 0, sizeof(double) * (size_t)d->atc_off.back(), ctx->stream));
        ACES_CUDA(ctx, cudaMemsetAsync(d->G, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        whiten_compact_kernel<<<(unsigned)d->n_wcols, 256, 0, ctx->stream>>>(
            d->n_wcols, d->w_blk, d->w_ptr, d->w_row, d->w_val, d->d_cols_off, d->d_atc_off, d->d_row0, d->d_tile_off,
            d->d_mp, row_map, d->tilesX, d->d_tuple_on, d->Atc);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        if (ctx->timing) ACES_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
        ACES_TRY(dense::gemm_batched(ctx, dense::GEMM_TN, d->g_gram, d->T, d->max_ct, d->max_ct));
        if (ctx->timing) {
            ACES_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
            ctx->timed = 1;
        }
        for (int t = 0; t < d->T; ++t) {  // one launch per block: the order of the additions is fixed
            if (d->ct[t] == 0 || !d->tuple_on[(size_t)t]) continue;
            dim3 grid((unsigned)std::min<int64_t>((d->ct[t] + 255) / 256, 8), (unsigned)d->ct[t]);
            gram_scatter_kernel<<<grid, 256, 0, ctx->stream>>>(d->ct[t], d->cpt[t], d->w_cols + d->cols_off[t],
                                                              d->Gt + d->gt_off[t], d->G, Np);
            ctx->launches += 1;
        }
        ACES_CUDA(ctx, cudaGetLastError());
    } else if (ls_type == 2) {
        ACES_CUDA(ctx, cudaMemsetAsync(d->At, 0, sizeof(double) * (size_t)Mp * Np, ctx->stream));
        whiten_tiles_kernel<<<(unsigned)d->N, 256, 0, ctx->stream>>>(d->N, d->a_colptr, d->a_rowval, d->a_nzval,
                                                                    d->blk_of_row, d->d_row0, d->d_prow0, d->d_tile_off,
                                                                    d->d_mp, row_map, d->tilesX, d->d_tuple_on, d->At, Mp);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        if (ctx->timing) ACES_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Mp, 1.0, d->At, Mp, d->At, Mp, 0.0, d->G, Np, 1));
        if (ctx->timing) {
            ACES_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
            ctx->timed = 1;
        }
    } else {
        ACES_CUDA(ctx, cudaMemsetAsync(d->G, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        sandwich_kernel<<<grid_for(d->N, 64), 64, 0, ctx->stream>>>(
            d->N, d->a_colptr, d->a_rowval, d->a_nzval, d->a_rowptr, d->a_colidx, d->a_vals, nullptr, nullptr, nullptr,
            d->blk_of_row, d->d_row0, d->d_prow0, row_map, fv.f, d->G, Np);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    ACES_TRY(dense::fill_diag_range(ctx, d->G, Np, d->N, Np, 1.0));
    return ACES_OK;
}

// The reference's fallback for a block whose Cholesky factorisation fails (src/merit.jl:403-424):
// take the elementwise absolute value and, if the smallest eigenvalue is then below epsilon, add
// (epsilon - eig_min) * I.  The reference gets eig_min from ARPACK; here it is located by bisection
// on the shift mu for which the block + mu * I is still factorisable (accurate to about
// n * eps * |block|, the same order as epsilon = 1e-12 -- this path is "parity unpinned").
int fallback_tiles(aces_ctx* ctx, aces_design* d, const double* val_dev, const int32_t* row_map,
                   const std::vector<int64_t>& kept_count, double epsilon) {
    std::vector<int32_t> fl((size_t)d->T);
    ACES_CUDA(ctx, cudaMemcpyAsync(fl.data(), d->flags, sizeof(int32_t) * fl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    d->fallback_blocks = 0;
    std::vector<double> hval;
    std::vector<int32_t> hmap;
    for (int t = 0; t < d->T; ++t) {
        if (!fl[(size_t)t]) continue;
        d->fallback_blocks += 1;
        if (hval.empty()) {
            hval.resize((size_t)d->nnz_cov);
            ACES_CUDA(ctx, cudaMemcpy(hval.data(), val_dev, sizeof(double) * hval.size(), cudaMemcpyDeviceToHost));
        }
        // infinity norm of the block bounds every eigenvalue
        // (kept rows and columns only: a clipped row can hold inf or huge values from 1 / lambda)
        if (row_map && hmap.empty()) {
            hmap.resize((size_t)d->M);
            ACES_CUDA(ctx, cudaMemcpy(hmap.data(), row_map, sizeof(int32_t) * hmap.size(), cudaMemcpyDeviceToHost));
        }
        double norm = 0.0;
        for (int64_t c = d->row0[t]; c < d->row0[t + 1]; ++c) {
            if (row_map && hmap[(size_t)c] < 0) continue;
            double sum = 0.0;
            for (int64_t e = d->h_cov_colptr[(size_t)c]; e < d->h_cov_colptr[(size_t)c + 1]; ++e) {
                if (row_map && hmap[(size_t)d->h_cov_rowval[(size_t)e]] < 0) continue;
                sum += std::fabs(hval[(size_t)e]);
            }
            norm = std::max(norm, sum);
        }
        const int64_t n = d->mp[t];
        double* tile = d->tilesL + d->tile_off[t];
        double* dinv = d->dinvT + d->dinv_off[t];
        int32_t* flag = d->flags + t;
        auto attempt = [&](double mu, bool* ok) -> int {
            ACES_CUDA(ctx, cudaMemsetAsync(tile, 0, sizeof(double) * (size_t)n * n, ctx->stream));
            ACES_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int32_t), ctx->stream));
            tile_rebuild_kernel<<<grid_for(d->m[t], 128), 128, 0, ctx->stream>>>(
                d->row0[t], d->row0[t + 1], d->cov_colptr, d->cov_rowval, val_dev, d->blk_of_row, t, d->row0[t], row_map,
                mu, tile, n);
            ACES_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1;
            ACES_TRY(dense::fill_diag_range(ctx, tile, n, kept_count[(size_t)t], n, 1.0));
            ACES_TRY(dense::potrf(ctx, n, tile, n, dinv, flag));
            int32_t f = 0;
            ACES_CUDA(ctx, cudaMemcpyAsync(&f, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            *ok = f == 0;
            return ACES_OK;
        };
        bool ok = false;
        double shift = 0.0;
        ACES_TRY(attempt(-epsilon, &ok));  // factorisable: the smallest eigenvalue is already >= epsilon
        if (!ok) {
            double lo = -epsilon, hi = 2.0 * norm + epsilon;
            for (int it = 0; it < 64 && hi - lo > 1e-16 * norm; ++it) {
                const double mid = 0.5 * (lo + hi);
                ACES_TRY(attempt(mid, &ok));
                if (ok) hi = mid; else lo = mid;
            }
            shift = hi + epsilon;  // hi ~ -eig_min
        }
        ACES_TRY(attempt(shift, &ok));
        if (!ok)
            return aces_fail(ctx, ACES_E_NOT_PD, "covariance block of tuple %d is not positive definite even after the "
                             "absolute-value / eigenvalue-shift fallback", t);
        ACES_TRY(dense::trtri(ctx, n, tile, n, dinv, d->tilesX + d->tile_off[t], n));
    }
    return ACES_OK;
}

// gram_dead_ok: the Gram matrix was factorised in dead-pivot mode (a fit): its flag counts the
// eliminated columns and is reported through aces_last_rank_deficiency instead of failing.
int check_flags(aces_ctx* ctx, aces_design* d, const char* what, bool gram_dead_ok = false) {
    std::vector<int32_t> fl((size_t)d->T + 1);
    ACES_CUDA(ctx, cudaMemcpyAsync(fl.data(), d->flags, sizeof(int32_t) * fl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < d->T; ++t)
        if (fl[(size_t)t])
            return aces_fail(ctx, ACES_E_NOT_PD, "%s: the covariance block of tuple %d is not positive definite", what, t);
    if (gram_dead_ok) {
        ctx->rank_deficiency = fl[(size_t)d->T];
        return ACES_OK;
    }
    if (fl[(size_t)d->T])
        return aces_fail(ctx, ACES_E_NOT_PD,
                         "%s: the Gram matrix is not positive definite (diagonal block %d): the whitened design "
                         "matrix is rank deficient", what, fl[(size_t)d->T] - 1);
    return ACES_OK;
}

// Prepares the estimator: covariance values (from eigenvalue estimates, or given), factor (GLS) or
// diagonal weights (OLS / WLS), Gram matrix and its Cholesky factor.
int factor_gram_matrix(aces_ctx* ctx, aces_design* d, const FitVectors& fv, PhaseTimer* pt, int dead_mode);
int prepare_estimator(aces_ctx* ctx, aces_design* d, int ls_type, const double* val_dev, const int32_t* row_map,
                      const std::vector<int64_t>& kept_count, const FitVectors& fv, PhaseTimer* pt = nullptr,
                      bool factor_gram = true, int dead_mode = 0) {
    d->fallback_blocks = 0;
    ACES_CUDA(ctx, cudaMemsetAsync(d->flags, 0, sizeof(int32_t) * ((size_t)d->T + 2), ctx->stream));
    ACES_CUDA(ctx, cudaMemsetAsync(d->vec, 0, sizeof(double) * (size_t)(8 * std::max(d->Mp, d->Np)), ctx->stream));
    if (ls_type == 2) {
        ACES_TRY(build_and_factor_tiles(ctx, d, val_dev, row_map, kept_count));
        // sparse_covariance_inv_factor's default epsilon (src/merit.jl:381)
        ACES_TRY(fallback_tiles(ctx, d, val_dev, row_map, kept_count, 1e-12));
        if (pt) pt->mark("block factor");
    } else {
        diag_factor_kernel<<<grid_for(d->M, 256), 256, 0, ctx->stream>>>(d->M, d->cov_colptr, d->cov_rowval, val_dev,
                                                                        d->blk_of_row, d->d_row0, d->d_prow0, row_map,
                                                                        d->d_tuple_on, ls_type == 0 ? 1 : 0, fv.f);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    ACES_TRY(build_gram(ctx, d, ls_type, fv, row_map));
    if (pt) pt->mark("gram");
    if (!factor_gram) return ACES_OK;  // a shard: the host sums the partial Gram matrices first
    return factor_gram_matrix(ctx, d, fv, pt, dead_mode);
}

// Cholesky factor of the Gram matrix (in place, lower) with the relative pivot test against the
// original diagonal.  dead_mode (fits): a column that fails the test is eliminated, which gives the
// basic solution of a rank-deficient problem (the reference's sparse QR also returns a basic
// solution there, src/estimate.jl:500-502; WHICH columns it zeroes depends on SPQR's column
// ordering, so only the fitted values A x are comparable in that case).
int factor_gram_matrix(aces_ctx* ctx, aces_design* d, const FitVectors& fv, PhaseTimer* pt, int dead_mode) {
    ACES_TRY(dense::get_diag(ctx, d->Np, d->G, d->Np, fv.t2));
    ACES_CUDA(ctx, cudaMemcpyAsync(fv.t1, fv.t2, sizeof(double) * (size_t)d->Np, cudaMemcpyDeviceToDevice, ctx->stream));
    ACES_TRY(dense::potrf(ctx, d->Np, d->G, d->Np, d->dinvG, d->flags + d->T, fv.t1, 1e-13, dead_mode));
    if (pt) pt->mark("potrf");
    return ACES_OK;
}

}  // namespace

extern "C" {

int32_t aces_design_create(aces_ctx* ctx, int32_t T, const int64_t* mapping_lengths, int64_t N,
                           const int32_t* A_colptr, const int32_t* A_rowval, const int32_t* A_nzval,
                           int32_t index_base, const int64_t* experiment_numbers, const double* shot_weights,
                           const double* tuple_times, const int64_t* diag_counts, const int64_t* cov_ptr,
                           const int64_t* cov_p, const int64_t* cov_q, const int64_t* cov_counts,
                           aces_design** out) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, out != nullptr, "design_create: out is NULL");
    *out = nullptr;
    ACES_CHECK(ctx, T >= 1 && N >= 1, "design_create: T=%d N=%lld", T, (long long)N);
    ACES_CHECK(ctx, mapping_lengths && A_colptr && A_rowval && A_nzval && experiment_numbers && shot_weights &&
                        tuple_times && diag_counts, "design_create: NULL argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "design_create: index_base must be 0 or 1");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    aces_design* d = new aces_design();
    auto fail = [&](int rc) {
        for (void* p : d->owned) cudaFree(p);
        delete d;
        return rc;
    };
    d->T = T;
    d->N = N;
    d->Np = pad_up(N);
    d->m.assign(mapping_lengths, mapping_lengths + T);
    d->mp.resize(T); d->row0.resize(T + 1); d->prow0.resize(T + 1); d->tile_off.resize(T + 1);
    d->dinv_off.resize(T + 1); d->key0.resize(T + 1);
    d->row0[0] = d->prow0[0] = d->tile_off[0] = d->dinv_off[0] = 0;
    for (int t = 0; t < T; ++t) {
        if (d->m[t] < 0) return fail(aces_fail(ctx, ACES_E_INVALID, "design_create: negative mapping length"));
        d->mp[t] = std::max<int64_t>(pad_up(d->m[t]), DB);
        d->row0[t + 1] = d->row0[t] + d->m[t];
        d->prow0[t + 1] = d->prow0[t] + d->mp[t];
        d->tile_off[t + 1] = d->tile_off[t] + d->mp[t] * d->mp[t];
        d->dinv_off[t + 1] = d->dinv_off[t] + d->mp[t] * DB;
        d->max_nb = std::max<int>(d->max_nb, (int)(d->mp[t] / DB));
        d->times_factor += shot_weights[t] * tuple_times[t];  // sum(d.shot_weights .* d.tuple_times)
    }
    d->M = d->row0[T];
    d->Mp = d->prow0[T];
    const int64_t M = d->M;
    d->h_blk_of_row.resize((size_t)M);
    for (int t = 0; t < T; ++t)
        for (int64_t r = d->row0[t]; r < d->row0[t + 1]; ++r) d->h_blk_of_row[(size_t)r] = t;

    // ---- design matrix: CSC (as given) and CSR -------------------------------------------------
    const int64_t nnzA = (int64_t)A_colptr[N] - index_base;
    d->nnzA = nnzA;
    std::vector<int64_t> colptr((size_t)N + 1), rowptr((size_t)M + 1, 0);
    std::vector<int32_t> rowval((size_t)nnzA), colidx((size_t)nnzA);
    std::vector<double> nzval((size_t)nnzA), vals((size_t)nnzA);
    for (int64_t j = 0; j <= N; ++j) colptr[(size_t)j] = (int64_t)A_colptr[j] - index_base;
    for (int64_t e = 0; e < nnzA; ++e) {
        const int64_t r = (int64_t)A_rowval[e] - index_base;
        if (r < 0 || r >= M) return fail(aces_fail(ctx, ACES_E_INVALID, "design_create: design-matrix row %lld outside 0..%lld", (long long)r, (long long)M - 1));
        rowval[(size_t)e] = (int32_t)r;
        nzval[(size_t)e] = (double)A_nzval[e];
        rowptr[(size_t)r + 1]++;
    }
    for (int64_t r = 0; r < M; ++r) rowptr[(size_t)r + 1] += rowptr[(size_t)r];
    {
        std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
        for (int64_t j = 0; j < N; ++j)
            for (int64_t e = colptr[(size_t)j]; e < colptr[(size_t)j + 1]; ++e) {
                const int64_t p = fill[(size_t)rowval[(size_t)e]]++;
                colidx[(size_t)p] = (int32_t)j;
                vals[(size_t)p] = nzval[(size_t)e];
            }
    }
    // ---- covariance pattern: the diagonal plus both orientations of every key ----------------------
    const int64_t C = cov_ptr ? cov_ptr[T] : 0;
    d->C = C;
    std::vector<int64_t> key_p((size_t)C), key_q((size_t)C);
    std::vector<double> dfac((size_t)M), kfac((size_t)C);
    for (int t = 0; t < T; ++t) {
        const double w = shot_weights[t];
        const double X = (double)experiment_numbers[t];
        for (int64_t r = d->row0[t]; r < d->row0[t + 1]; ++r) {
            if (diag_counts[r] < 1) return fail(aces_fail(ctx, ACES_E_INVALID, "design_create: diag_counts[%lld] < 1", (long long)r));
            dfac[(size_t)r] = (1.0 / w) * (X / (double)diag_counts[r]);  // src/merit.jl:250-251
        }
        d->key0[t] = cov_ptr ? cov_ptr[t] : 0;
    }
    d->key0[T] = C;
    std::vector<std::vector<std::pair<int64_t, int32_t>>> cols((size_t)M);  // per column: (row, src)
    for (int64_t r = 0; r < M; ++r) cols[(size_t)r].push_back({r, -1});
    for (int t = 0; t < T; ++t) {
        const double w = shot_weights[t];
        for (int64_t k = d->key0[t]; k < d->key0[t + 1]; ++k) {
            const int64_t p = cov_p[k] - index_base, q = cov_q[k] - index_base;
            if (p < 0 || q < 0 || p >= d->m[t] || q >= d->m[t] || p == q)
                return fail(aces_fail(ctx, ACES_E_INVALID, "design_create: covariance key %lld = (%lld, %lld) invalid for tuple %d", (long long)k, (long long)p, (long long)q, t));
            const int64_t gp = d->row0[t] + p, gq = d->row0[t] + q;
            key_p[(size_t)k] = gp;
            key_q[(size_t)k] = gq;
            // src/merit.jl:269-273: (1 / w) * ((X * c_pq) / (c_pp * c_qq)), integer products first
            kfac[(size_t)k] = (1.0 / w) * ((double)(experiment_numbers[t] * cov_counts[k]) /
                                           (double)(diag_counts[gp] * diag_counts[gq]));
            cols[(size_t)gq].push_back({gp, (int32_t)k});
            cols[(size_t)gp].push_back({gq, (int32_t)k});
        }
    }
    std::vector<int64_t> cov_colptr((size_t)M + 1, 0), cov_rowval;
    std::vector<int32_t> cov_src;
    for (int64_t c = 0; c < M; ++c) {
        auto& v = cols[(size_t)c];
        std::sort(v.begin(), v.end());
        for (size_t i = 0; i < v.size(); ++i) {
            if (i > 0 && v[i].first == v[i - 1].first)
                return fail(aces_fail(ctx, ACES_E_INVALID, "design_create: duplicate covariance key at (%lld, %lld)", (long long)v[i].first, (long long)c));
            cov_rowval.push_back(v[i].first);
            cov_src.push_back(v[i].second);
        }
        cov_colptr[(size_t)c + 1] = (int64_t)cov_rowval.size();
    }
    d->nnz_cov = (int64_t)cov_rowval.size();
    d->h_cov_colptr = cov_colptr;
    d->h_cov_rowval = cov_rowval;

    // ---- column windows of the blocks (for the GLS Gram) -----------------------------------------
    {
        d->ct.assign(T, 0); d->cpt.assign(T, 0);
        d->cols_off.assign(T + 1, 0); d->atc_off.assign(T + 1, 0); d->gt_off.assign(T + 1, 0);
        std::vector<int64_t> w_cols, w_ptr(1, 0);
        std::vector<int32_t> w_blk, w_row;
        std::vector<double> w_val;
        // entries of A grouped by (block, column): CSC order already groups rows of a column by block
        std::vector<std::vector<int64_t>> cols_of((size_t)T);
        for (int64_t j = 0; j < N; ++j) {
            int last = -1;
            for (int64_t e = colptr[(size_t)j]; e < colptr[(size_t)j + 1]; ++e) {
                const int t = d->h_blk_of_row[(size_t)rowval[(size_t)e]];
                if (t != last) { cols_of[(size_t)t].push_back(j); last = t; }  // rows are sorted within a column
            }
        }
        double flops_w = 0.0;
        for (int t = 0; t < T; ++t) {
            auto& cs = cols_of[(size_t)t];
            std::sort(cs.begin(), cs.end());
            cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
            d->ct[t] = (int64_t)cs.size();
            d->cpt[t] = std::max<int64_t>(pad_up(d->ct[t]), DB);
            d->cols_off[t + 1] = d->cols_off[t] + d->ct[t];
            d->atc_off[t + 1] = d->atc_off[t] + d->mp[t] * d->cpt[t];
            d->gt_off[t + 1] = d->gt_off[t] + d->cpt[t] * d->cpt[t];
            d->max_ct = std::max<int>(d->max_ct, (int)(d->cpt[t] / DB));
            flops_w += (double)d->mp[t] * (double)d->cpt[t] * (double)d->cpt[t];
            for (int64_t j : cs) {
                w_cols.push_back(j);
                w_blk.push_back(t);
                for (int64_t e = colptr[(size_t)j]; e < colptr[(size_t)j + 1]; ++e) {
                    const int64_t r = rowval[(size_t)e];
                    if (d->h_blk_of_row[(size_t)r] != t) continue;
                    w_row.push_back((int32_t)r);
                    w_val.push_back(nzval[(size_t)e]);
                }
                w_ptr.push_back((int64_t)w_row.size());
            }
        }
        d->n_wcols = (int64_t)w_cols.size();
        d->windowed = flops_w < 0.5 * (double)d->Mp * (double)d->Np * (double)d->Np;
        {   // uploaded even when the fit uses the dense Gram: aces_design_merit_grad needs the per-tuple blocks
            int rcw;
#define UPW(field, vecname) if ((rcw = dev_upload(ctx, d, &d->field, vecname)) != ACES_OK) return fail(rcw)
            UPW(w_blk, w_blk); UPW(w_ptr, w_ptr); UPW(w_row, w_row); UPW(w_val, w_val); UPW(w_cols, w_cols);
            UPW(d_cols_off, d->cols_off); UPW(d_atc_off, d->atc_off);
#undef UPW
        }
    }
    int rc;
#define UP(field, vecname) if ((rc = dev_upload(ctx, d, &d->field, vecname)) != ACES_OK) return fail(rc)
    UP(a_colptr, colptr); UP(a_rowval, rowval); UP(a_nzval, nzval);
    UP(a_rowptr, rowptr); UP(a_colidx, colidx); UP(a_vals, vals);
    UP(cov_colptr, cov_colptr); UP(cov_rowval, cov_rowval); UP(cov_src, cov_src);
    UP(blk_of_row, d->h_blk_of_row); UP(dfac, dfac); UP(kfac, kfac); UP(key_p, key_p); UP(key_q, key_q);
    UP(d_row0, d->row0); UP(d_prow0, d->prow0); UP(d_tile_off, d->tile_off); UP(d_mp, d->mp);
    d->tuple_on.assign((size_t)T, 1);
    UP(d_tuple_on, d->tuple_on);
#undef UP
    *out = d;
    return ACES_OK;
}

int32_t aces_design_destroy(aces_ctx* ctx, aces_design* d) {
    if (!ctx) return ACES_E_INVALID;
    if (!d) return ACES_OK;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (void* p : d->owned) cudaFree(p);
    delete d;
    return ACES_OK;
}

int64_t aces_design_covariance_nnz(const aces_design* d) { return d ? d->nnz_cov : -1; }
int32_t aces_design_fallback_count(const aces_design* d) { return d ? d->fallback_blocks : -1; }
int32_t aces_design_phase_ms(const aces_design* d, int32_t index, char* name, int32_t name_len, float* ms) {
    if (!d || index < 0 || (size_t)index >= d->phase_ms.size()) return ACES_E_INVALID;
    if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", d->phase_ms[(size_t)index].first.c_str());
    if (ms) *ms = d->phase_ms[(size_t)index].second;
    return ACES_OK;
}
double aces_design_gram_flops(const aces_design* d) {
    if (!d) return -1.0;
    double tiles_k = 0.0;  // sum over 128 x 128 output tiles of their inner dimension
    if (d->windowed) {
        for (int t = 0; t < d->T; ++t) {
            const double ct = (double)(d->cpt[t] / DB);
            tiles_k += 0.5 * ct * (ct + 1.0) * (double)d->mp[t];
        }
    } else {
        const double nt = (double)(d->Np / DB);
        tiles_k = 0.5 * nt * (nt + 1.0) * (double)d->Mp;
    }
    return 2.0 * DB * DB * tiles_k;
}

int32_t aces_design_covariance_pattern(const aces_design* d, int64_t* colptr, int64_t* rowval) {
    if (!d || !colptr || !rowval) return ACES_E_INVALID;
    std::copy(d->h_cov_colptr.begin(), d->h_cov_colptr.end(), colptr);
    std::copy(d->h_cov_rowval.begin(), d->h_cov_rowval.end(), rowval);
    return ACES_OK;
}

}  // extern "C"

namespace {

// Descriptor tables of the batched tile factorisation (built on first use: they hold workspace addresses).
int build_tile_descs(aces_ctx* ctx, aces_design* d) {
    if (d->dd) return ACES_OK;
    const int T = d->T, nbm = d->max_nb;
    std::vector<DiagDesc> dd((size_t)nbm * T);
    std::vector<GemmDesc> gp((size_t)nbm * T), gt((size_t)nbm * T), g1((size_t)nbm * T), g2((size_t)nbm * T);
    for (int k = 0; k < nbm; ++k)
        for (int t = 0; t < T; ++t) {
            const size_t idx = (size_t)k * T + t;
            const int64_t n = d->mp[t];
            const int nb = (int)(n / DB);
            DiagDesc& a = dd[idx];
            GemmDesc z{};
            gp[idx] = gt[idx] = g1[idx] = g2[idx] = z;
            a = DiagDesc{};
            if (k >= nb || !d->tuple_on[(size_t)t]) continue;  // inactive entry
            double* L = d->tilesL + d->tile_off[t];
            double* X = d->tilesX + d->tile_off[t];
            double* dk = d->dinvT + d->dinv_off[t] + (int64_t)k * DB * DB;
            double* Akk = L + (int64_t)k * DB * (n + 1);
            a.A = Akk; a.lda = n; a.dinv = dk; a.diag0 = nullptr; a.flag = d->flags + t; a.code = 1;
            const int rows_t = nb - k - 1;  // 128-row tiles below the diagonal block
            if (rows_t > 0) {
                double* panel = Akk + DB;
                gp[idx] = GemmDesc{panel, dk, panel, n, DB, n, DB, rows_t, 1, 0, 1.0, 0.0};
                gt[idx] = GemmDesc{panel, panel, Akk + (int64_t)DB * (n + 1), n, n, n, DB, rows_t, rows_t, 1, -1.0, 1.0};
            }
            double* Xk = X + (int64_t)k * DB;  // block row k of X, columns 0 .. (k+1)*128
            g1[idx] = GemmDesc{dk, Xk, Xk, DB, n, n, DB, 1, k + 1, 0, 1.0, 0.0};
            if (rows_t > 0)
                g2[idx] = GemmDesc{L + (int64_t)(k + 1) * DB + (int64_t)k * DB * n, Xk, Xk + DB, n, n, n, DB, rows_t, k + 1, 0, -1.0, 1.0};
        }
    ACES_TRY(dev_upload(ctx, d, &d->dd, dd));
    ACES_TRY(dev_upload(ctx, d, &d->g_panel, gp));
    ACES_TRY(dev_upload(ctx, d, &d->g_trail, gt));
    ACES_TRY(dev_upload(ctx, d, &d->g_tri1, g1));
    ACES_TRY(dev_upload(ctx, d, &d->g_tri2, g2));
    return ACES_OK;
}

int check_design(aces_ctx* ctx, const aces_design* d, const char* what) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, d != nullptr, "%s: design is NULL", what);
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    return ACES_OK;
}

// device copies of the eigenvalue vectors; returns pointers inside the vector workspace tail
int upload_eigenvalues(aces_ctx* ctx, aces_design* d, const double* eig, const double* cov_eig, double** lam_dev,
                       double** lamcov_dev) {
    // a buffer of the design, not context scratch: fit_begin keeps the pointer for the later stages
    // of a staged fit, and other calls on the same context may run in between
    double* base = d->lam_buf;
    *lam_dev = base;
    *lamcov_dev = base + d->M;
    ACES_CUDA(ctx, cudaMemcpyAsync(*lam_dev, eig, sizeof(double) * (size_t)d->M, cudaMemcpyDefault, ctx->stream));
    if (d->C > 0) {
        ACES_CHECK(ctx, cov_eig != nullptr, "design: the design has covariance keys but covariance eigenvalues are NULL");
        ACES_CUDA(ctx, cudaMemcpyAsync(*lamcov_dev, cov_eig, sizeof(double) * (size_t)d->C, cudaMemcpyDefault, ctx->stream));
    }
    return ACES_OK;
}

}  // namespace

extern "C" {

int32_t aces_calc_covariance(aces_ctx* ctx, aces_design* d, const double* eigenvalues,
                             const double* covariance_eigenvalues, double epsilon, int32_t weight_time,
                             int32_t log_form, double* nzval) {
    ACES_TRY(check_design(ctx, d, "calc_covariance"));
    ACES_CHECK(ctx, eigenvalues && nzval, "calc_covariance: NULL argument");
    ACES_TRY(ensure_workspaces(ctx, d, false, false, false, false));
    double *lam, *lamcov;
    ACES_TRY(upload_eigenvalues(ctx, d, eigenvalues, covariance_eigenvalues, &lam, &lamcov));
    ACES_TRY(launch_cov_values(ctx, d, lam, lamcov, epsilon, weight_time, log_form, d->covval));
    ACES_CUDA(ctx, cudaMemcpyAsync(nzval, d->covval, sizeof(double) * (size_t)d->nnz_cov, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}

int32_t aces_design_inv_factor(aces_ctx* ctx, aces_design* d, const double* covariance_log_nzval,
                               const int64_t* kept_rows, int64_t n_kept, int32_t index_base,
                               double fallback_epsilon, double* factor_blocks, int64_t* block_lengths) {
    ACES_TRY(check_design(ctx, d, "design_inv_factor"));
    ACES_CHECK(ctx, !d->sharded, "design_inv_factor: the design is a tuple shard");
    ACES_CHECK(ctx, covariance_log_nzval && factor_blocks, "design_inv_factor: NULL argument");
    d->est_ls_type = -1;
    ACES_TRY(ensure_workspaces(ctx, d, true, false, false, false));
    ACES_TRY(build_tile_descs(ctx, d));
    std::vector<int64_t> kept;
    bool clipped;
    ACES_TRY(upload_clipping(ctx, d, kept_rows, n_kept, index_base, &kept, &clipped));
    ACES_CUDA(ctx, cudaMemsetAsync(d->flags, 0, sizeof(int32_t) * ((size_t)d->T + 2), ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d->covval, covariance_log_nzval, sizeof(double) * (size_t)d->nnz_cov, cudaMemcpyDefault,
                                   ctx->stream));
    d->fallback_blocks = 0;
    ACES_TRY(build_and_factor_tiles(ctx, d, d->covval, clipped ? d->row_map : nullptr, kept));
    if (fallback_epsilon > 0.0)
        ACES_TRY(fallback_tiles(ctx, d, d->covval, clipped ? d->row_map : nullptr, kept, fallback_epsilon));
    // export: blocks of kept[t] x kept[t], concatenated
    std::vector<int64_t> out_off((size_t)d->T + 1, 0);
    for (int t = 0; t < d->T; ++t) out_off[(size_t)t + 1] = out_off[(size_t)t] + kept[(size_t)t] * kept[(size_t)t];
    int64_t* meta = d->kept_dev;  // [kept | out_off]
    ACES_CUDA(ctx, cudaMemcpyAsync(meta, kept.data(), sizeof(int64_t) * (size_t)d->T, cudaMemcpyHostToDevice, ctx->stream));
    const size_t hdr = ((size_t)d->T + 2) & ~(size_t)1;  // int64 slots before the (8-byte aligned) blocks
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[1], sizeof(int64_t) * hdr + sizeof(double) * (size_t)out_off.back()));
    int64_t* off_dev = (int64_t*)ctx->d_work[1].ptr;
    double* out_dev = (double*)(off_dev + hdr);
    ACES_CUDA(ctx, cudaMemcpyAsync(off_dev, out_off.data(), sizeof(int64_t) * ((size_t)d->T + 1), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid(148, (unsigned)d->T);
    tiles_export_kernel<<<grid, 256, 0, ctx->stream>>>(d->T, d->d_tile_off, d->d_mp, meta, off_dev, d->tilesX, out_dev);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(factor_blocks, out_dev, sizeof(double) * (size_t)out_off.back(), cudaMemcpyDefault, ctx->stream));
    if (block_lengths) std::copy(kept.begin(), kept.end(), block_lengths);
    return check_flags(ctx, d, "design_inv_factor");
}

// ---- the fit, split at its reduction points -----------------------------------------------------------
// aces_design_fit runs the stages back to back on one design.  A tuple-sharded fit (the
// Gram matrix is a sum over the tuple blocks because the covariance is block diagonal by tuple,
// src/merit.jl:284) runs them on every shard and sums the buffers of aces_design_fit_buffers over
// the shards between the stages (ncclAllReduce on the device pointers); the Cholesky factorisation
// and the triangular solves are replicated.

namespace {

// right-hand side of the (corrected semi-) normal equations: g = A' F' F (b - A x), x may be NULL
int fit_normal_rhs(aces_ctx* ctx, aces_design* d, const double* x_or_null) {
    FitVectors fv = fit_vectors(d);
    const int32_t* row_map = d->fs_clipped ? d->row_map : nullptr;
    const unsigned gm = grid_for(d->M, 128), gn = grid_for(d->N, 128);
    residual_kernel<<<gm, 128, 0, ctx->stream>>>(d->M, d->a_rowptr, d->a_colidx, d->a_vals, d->blk_of_row, d->d_row0,
                                                d->d_prow0, row_map, d->fs_lam, x_or_null, fv.u);
    ctx->launches += 1;
    ACES_TRY(apply_F(ctx, d, d->fs_ls_type, fv, fv.u, fv.v, false));
    ACES_TRY(apply_F(ctx, d, d->fs_ls_type, fv, fv.v, fv.w, true));
    at_times_kernel<<<gn, 128, 0, ctx->stream>>>(d->N, d->a_colptr, d->a_rowval, d->a_nzval, d->blk_of_row, d->d_row0,
                                                d->d_prow0, row_map, fv.w, fv.g);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

int fit_begin(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* est_eigenvalues,
              const double* est_covariance_eigenvalues, const int64_t* kept_rows, int64_t n_kept,
              int32_t index_base, double epsilon, PhaseTimer* pt, bool factor_gram) {
    ACES_CHECK(ctx, ls_type >= 0 && ls_type <= 2, "design_fit: ls_type must be 0 (OLS), 1 (WLS) or 2 (GLS)");
    ACES_CHECK(ctx, est_eigenvalues != nullptr, "design_fit: NULL argument");
    d->fs_stage = 0;
    const bool gls = ls_type == 2;
    ACES_TRY(ensure_workspaces(ctx, d, gls, gls, false, false));
    if (gls) ACES_TRY(build_tile_descs(ctx, d));
    std::vector<int64_t> kept;
    bool clipped;
    ACES_TRY(upload_clipping(ctx, d, kept_rows, n_kept, index_base, &kept, &clipped));
    const int32_t* row_map = clipped ? d->row_map : nullptr;
    double *lam, *lamcov;
    ACES_TRY(upload_eigenvalues(ctx, d, est_eigenvalues, est_covariance_eigenvalues, &lam, &lamcov));
    FitVectors fv = fit_vectors(d);
    // K2: log-covariance of the estimates, weight_time = false (src/estimate.jl:713-720)
    if (ls_type != 0) ACES_TRY(launch_cov_values(ctx, d, lam, lamcov, epsilon, 0, 1, d->covval));
    if (pt) pt->mark("covariance");
    d->fs_ls_type = ls_type;
    d->fs_clipped = clipped;
    d->fs_lam = lam;
    d->est_ls_type = -1;
    ACES_TRY(prepare_estimator(ctx, d, ls_type, d->covval, row_map, kept, fv, pt, factor_gram, 1));
    d->est_ls_type = ls_type;
    ACES_TRY(fit_normal_rhs(ctx, d, nullptr));  // g = A' F' F b
    d->fs_stage = 1;
    return ACES_OK;
}

int fit_solve(aces_ctx* ctx, aces_design* d, bool factor_gram, PhaseTimer* pt) {
    ACES_CHECK(ctx, d->fs_stage == 1, "design_fit_solve: call aces_design_fit_begin first");
    FitVectors fv = fit_vectors(d);
    if (factor_gram) {
        // padding rows: the shards each wrote 1.0 on the padding diagonal, the sum is their number
        ACES_TRY(factor_gram_matrix(ctx, d, fv, pt, 1));
    }
    ACES_TRY(dense::trsv_lower(ctx, d->Np, d->G, d->Np, d->dinvG, fv.g, fv.t1));
    ACES_TRY(dense::trsv_lower_t(ctx, d->Np, d->G, d->Np, d->dinvG, fv.t1, fv.x));
    if (pt) pt->mark("first solve");
    d->fs_stage = 2;
    return ACES_OK;
}

// x += inv(G) g; returns max|x| and max|delta| (host) for the stopping rule
int fit_correct(aces_ctx* ctx, aces_design* d, double* x_max, double* delta_max) {
    FitVectors fv = fit_vectors(d);
    const unsigned gn = grid_for(d->N, 128);
    ACES_TRY(dense::trsv_lower(ctx, d->Np, d->G, d->Np, d->dinvG, fv.g, fv.t1));
    ACES_TRY(dense::trsv_lower_t(ctx, d->Np, d->G, d->Np, d->dinvG, fv.t1, fv.t2));
    axpy_kernel<<<gn, 128, 0, ctx->stream>>>(d->N, 1.0, fv.t2, fv.x);
    ctx->launches += 1;
    if (x_max) {
        max_abs_pair_kernel<<<1, 256, 0, ctx->stream>>>(d->N, fv.x, fv.t2, d->mom);
        ctx->launches += 1;
        double h[2] = {0.0, 0.0};
        ACES_CUDA(ctx, cudaMemcpyAsync(h, d->mom, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *x_max = h[0];
        *delta_max = h[1];
    }
    return ACES_OK;
}

int fit_end(aces_ctx* ctx, aces_design* d, int32_t constrain, double* gate_eigenvalues) {
    ACES_CHECK(ctx, d->fs_stage == 2, "design_fit_end: call aces_design_fit_solve first");
    ACES_CHECK(ctx, gate_eigenvalues != nullptr, "design_fit: NULL argument");
    FitVectors fv = fit_vectors(d);
    exp_neg_kernel<<<grid_for(d->N, 128), 128, 0, ctx->stream>>>(d->N, fv.x, fv.t1, constrain);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(gate_eigenvalues, fv.t1, sizeof(double) * (size_t)d->N, cudaMemcpyDefault, ctx->stream));
    d->fs_stage = 0;
    return check_flags(ctx, d, "design_fit", true);
}

}  // namespace

int32_t aces_design_fit(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* est_eigenvalues,
                        const double* est_covariance_eigenvalues, const int64_t* kept_rows, int64_t n_kept,
                        int32_t index_base, double epsilon, int32_t constrain, double* gate_eigenvalues) {
    ACES_TRY(check_design(ctx, d, "design_fit"));
    ACES_CHECK(ctx, !d->sharded, "design_fit: the design is a tuple shard; use aces_design_fit_begin .. _end");
    ACES_CHECK(ctx, gate_eigenvalues != nullptr, "design_fit: NULL argument");
    PhaseTimer pt(ctx, d);
    ACES_TRY(fit_begin(ctx, d, ls_type, est_eigenvalues, est_covariance_eigenvalues, kept_rows, n_kept, index_base,
                       epsilon, &pt, true));
    ACES_TRY(fit_solve(ctx, d, false, &pt));
    for (int it = 0; it < 2; ++it) {
        // corrected semi-normal equations: r = F (b - A x) in FP64, x += inv(G) A' F' r
        ACES_TRY(fit_normal_rhs(ctx, d, fit_vectors(d).x));
        double xm = 0.0, dm = 0.0;
        ACES_TRY(fit_correct(ctx, d, it == 0 ? &xm : nullptr, &dm));
        // The correction contracts the error by about |delta| / |x| per step: when the first one is
        // already below 1e-7 relative, a second step would change x by < 1e-14.
        if (it == 0 && dm <= 1e-7 * xm) break;
    }
    pt.mark("refinement");
    return fit_end(ctx, d, constrain, gate_eigenvalues);
}

int32_t aces_design_set_tuple_shard(aces_ctx* ctx, aces_design* d, const uint8_t* tuple_on) {
    ACES_TRY(check_design(ctx, d, "design_set_tuple_shard"));
    std::vector<uint8_t> on((size_t)d->T, 1);
    bool all = true;
    if (tuple_on)
        for (int t = 0; t < d->T; ++t) {
            on[(size_t)t] = tuple_on[t] ? 1 : 0;
            all = all && on[(size_t)t];
        }
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    d->tuple_on = on;
    d->sharded = !all;
    ACES_CUDA(ctx, cudaMemcpy(d->d_tuple_on, on.data(), (size_t)d->T, cudaMemcpyHostToDevice));
    // the batched descriptor tables carry the mask (inactive entries): rebuilt on the next use
    d->dd = nullptr;
    d->g_panel = d->g_trail = d->g_tri1 = d->g_tri2 = nullptr;
    d->g_gram = nullptr;
    d->fs_stage = 0;
    d->est_ls_type = -1;
    return ACES_OK;
}

int32_t aces_design_fit_begin(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* est_eigenvalues,
                              const double* est_covariance_eigenvalues, const int64_t* kept_rows, int64_t n_kept,
                              int32_t index_base, double epsilon) {
    ACES_TRY(check_design(ctx, d, "design_fit_begin"));
    ACES_TRY(fit_begin(ctx, d, ls_type, est_eigenvalues, est_covariance_eigenvalues, kept_rows, n_kept, index_base,
                       epsilon, nullptr, false));
    // a failed block factorisation must surface on the shard that owns the block, before the sum
    return check_flags(ctx, d, "design_fit_begin");
}

int32_t aces_design_fit_buffers(aces_design* d, double** gram, int64_t* gram_elems, double** rhs, int64_t* rhs_elems) {
    if (!d || !d->G) return ACES_E_INVALID;
    if (gram) *gram = d->G;
    if (gram_elems) *gram_elems = d->Np * d->Np;
    if (rhs) *rhs = fit_vectors(d).g;
    if (rhs_elems) *rhs_elems = d->Np;
    return ACES_OK;
}

int32_t aces_design_fit_solve(aces_ctx* ctx, aces_design* d) {
    ACES_TRY(check_design(ctx, d, "design_fit_solve"));
    return fit_solve(ctx, d, true, nullptr);
}

// ---- Cholesky of the summed Gram matrix distributed over the shards (last row) ---------------
// Every shard holds the whole summed Gram matrix after the all-reduce.  Block columns of DIST_PW columns are
// owned cyclically; step p: the owner factorises block column p in place (diagonal block by dense::potrf,
// the rows below it with the explicit inverse of that block on the tensor pipe), the host broadcasts the
// DIST_PW full columns (one contiguous range of the column-major matrix) and the inverses of their
// 128-blocks to every shard, and every shard applies the panel to the block columns IT owns.  Columns a
// shard does not own are never updated there: they arrive, final, with their owner's broadcast.  After the
// last step every shard holds the complete factor, and the triangular solves run replicated as before.
int32_t aces_design_fit_dist_begin(aces_ctx* ctx, aces_design* d) {
    ACES_TRY(check_design(ctx, d, "design_fit_dist_begin"));
    ACES_CHECK(ctx, d->fs_stage == 1, "design_fit_dist_begin: call aces_design_fit_begin (and sum the buffers) first");
    FitVectors fv = fit_vectors(d);
    // the diagonal before the factorisation: reference of the relative pivot test, as in factor_gram_matrix
    ACES_TRY(dense::get_diag(ctx, d->Np, d->G, d->Np, fv.t2));
    ACES_CUDA(ctx, cudaMemcpyAsync(fv.t1, fv.t2, sizeof(double) * (size_t)d->Np, cudaMemcpyDeviceToDevice, ctx->stream));
    return ACES_OK;
}

int32_t aces_design_fit_dist_factor(aces_ctx* ctx, aces_design* d, int64_t col0, int64_t ncols) {
    ACES_TRY(check_design(ctx, d, "design_fit_dist_factor"));
    const int64_t Np = d->Np;
    ACES_CHECK(ctx, d->fs_stage == 1 && col0 >= 0 && ncols > 0 && col0 % DB == 0 && ncols % DB == 0 && col0 + ncols <= Np,
               "design_fit_dist_factor: bad block column [%lld, %lld)", (long long)col0, (long long)(col0 + ncols));
    FitVectors fv = fit_vectors(d);
    double* Akk = d->G + col0 * (Np + 1);
    double* dinv = d->dinvG + (col0 / DB) * DB * DB;
    ACES_TRY(dense::potrf(ctx, ncols, Akk, Np, dinv, d->flags + d->T, fv.t1 + col0, 1e-13, 1));
    const int64_t rows = Np - col0 - ncols;
    if (rows > 0) {
        // rows below: L_below = A_below * inv(L_kk)': X = inv(L_kk), a copy of A_below, one NT GEMM back in place
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[7], sizeof(double) * (size_t)(ncols * ncols + rows * ncols)));
        double* X = (double*)ctx->d_work[7].ptr;
        double* Tm = X + ncols * ncols;
        ACES_TRY(dense::trtri(ctx, ncols, Akk, Np, dinv, X, ncols));
        double* below = Akk + ncols;
        ACES_CUDA(ctx, cudaMemcpy2DAsync(Tm, sizeof(double) * (size_t)rows, below, sizeof(double) * (size_t)Np,
                                         sizeof(double) * (size_t)rows, (size_t)ncols, cudaMemcpyDeviceToDevice, ctx->stream));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NT, rows, ncols, ncols, 1.0, Tm, rows, X, ncols, 0.0, below, Np, 0));
    }
    return ACES_OK;
}

int32_t aces_design_fit_dist_update(aces_ctx* ctx, aces_design* d, int64_t pcol0, int64_t pncols, int32_t n_targets,
                                    const int64_t* target_col0, const int64_t* target_ncols) {
    ACES_TRY(check_design(ctx, d, "design_fit_dist_update"));
    const int64_t Np = d->Np;
    ACES_CHECK(ctx, d->fs_stage == 1 && pcol0 >= 0 && pncols > 0 && pcol0 % DB == 0 && pncols % 16 == 0 && pcol0 + pncols <= Np,
               "design_fit_dist_update: bad panel");
    ACES_CHECK(ctx, n_targets >= 0 && (n_targets == 0 || (target_col0 && target_ncols)), "design_fit_dist_update: NULL targets");
    for (int q = 0; q < n_targets; ++q) {
        const int64_t t0 = target_col0[q], tn = target_ncols[q];
        ACES_CHECK(ctx, t0 >= pcol0 + pncols && tn > 0 && t0 % DB == 0 && tn % DB == 0 && t0 + tn <= Np,
                   "design_fit_dist_update: target block column [%lld, %lld) is not behind the panel", (long long)t0,
                   (long long)(t0 + tn));
        // G[t0:, t0:t0+tn] -= L[t0:, panel] * L[t0:t0+tn, panel]'
        const double* P = d->G + t0 + pcol0 * Np;
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NT, Np - t0, tn, pncols, -1.0, P, Np, P, Np, 1.0, d->G + t0 * (Np + 1), Np, 0));
    }
    return ACES_OK;
}

int32_t aces_design_fit_dist_buffers(aces_design* d, double** block_inverses, int64_t* block_inverse_elems,
                                     int32_t** dead_columns) {
    if (!d || !d->G) return ACES_E_INVALID;
    if (block_inverses) *block_inverses = d->dinvG;
    if (block_inverse_elems) *block_inverse_elems = d->Np * DB;
    if (dead_columns) *dead_columns = d->flags + d->T;
    return ACES_OK;
}

int32_t aces_design_fit_dist_solve(aces_ctx* ctx, aces_design* d) {
    ACES_TRY(check_design(ctx, d, "design_fit_dist_solve"));
    return fit_solve(ctx, d, false, nullptr);
}

int32_t aces_design_fit_refine_begin(aces_ctx* ctx, aces_design* d) {
    ACES_TRY(check_design(ctx, d, "design_fit_refine_begin"));
    ACES_CHECK(ctx, d->fs_stage == 2, "design_fit_refine_begin: call aces_design_fit_solve first");
    return fit_normal_rhs(ctx, d, fit_vectors(d).x);
}

int32_t aces_design_fit_refine_end(aces_ctx* ctx, aces_design* d, double* x_max, double* delta_max) {
    ACES_TRY(check_design(ctx, d, "design_fit_refine_end"));
    ACES_CHECK(ctx, d->fs_stage == 2 && x_max && delta_max, "design_fit_refine_end: bad state or NULL argument");
    return fit_correct(ctx, d, x_max, delta_max);
}

int32_t aces_design_fit_end(aces_ctx* ctx, aces_design* d, int32_t constrain, double* gate_eigenvalues) {
    ACES_TRY(check_design(ctx, d, "design_fit_end"));
    return fit_end(ctx, d, constrain, gate_eigenvalues);
}

int32_t aces_design_precision_blocks(aces_ctx* ctx, aces_design* d, int32_t ls_type, int32_t G,
                                     const int64_t* group_ptr, const int32_t* group_idx, int32_t index_base,
                                     const double* gate_eigenvalues, double* blocks) {
    ACES_TRY(check_design(ctx, d, "design_precision_blocks"));
    ACES_CHECK(ctx, !d->sharded, "design_precision_blocks: the design is a tuple shard");
    ACES_CHECK(ctx, G >= 1 && group_ptr && group_idx && blocks, "design_precision_blocks: bad argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "design_precision_blocks: index_base must be 0 or 1");
    ACES_CHECK(ctx, d->est_ls_type == ls_type && d->fs_stage == 0,
               "design_precision_blocks: call aces_design_fit with the same ls_type on this design first (the blocks "
               "are those of the Gram matrix of that fit: its clipping, weights and block factors)");
    ACES_CHECK(ctx, group_ptr[0] == 0, "design_precision_blocks: group_ptr must start at 0");
    const int64_t nidx = group_ptr[G];
    std::vector<int32_t> idx((size_t)nidx);
    std::vector<int64_t> boff((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) {
        const int64_t n = group_ptr[g + 1] - group_ptr[g];
        ACES_CHECK(ctx, n >= 0 && n <= 4096, "design_precision_blocks: bad group size");
        boff[(size_t)g + 1] = boff[(size_t)g] + n * n;
        for (int64_t e = group_ptr[g]; e < group_ptr[g + 1]; ++e) {
            const int64_t i = (int64_t)group_idx[e] - index_base;
            ACES_CHECK(ctx, i >= 0 && i < d->N, "design_precision_blocks: index %lld outside 0..N-1", (long long)i);
            idx[(size_t)e] = (int32_t)i;
        }
    }
    FitVectors fv = fit_vectors(d);
    // the Gram matrix of the last fit again (its Cholesky factor overwrote it): same clipping and factors
    ACES_TRY(build_gram(ctx, d, ls_type, fv, d->fs_clipped ? d->row_map : nullptr));
    d->est_ls_type = ls_type;
    const size_t b_ptr = sizeof(int64_t) * ((size_t)G + 1), b_idx = (sizeof(int32_t) * (size_t)nidx + 7) & ~(size_t)7;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[6], 2 * b_ptr + b_idx + sizeof(double) * (size_t)(boff.back() + d->N) + 64));
    uint8_t* base = (uint8_t*)ctx->d_work[6].ptr;
    int64_t* d_ptr = (int64_t*)base;
    int64_t* d_boff = (int64_t*)(base + b_ptr);
    int32_t* d_idx = (int32_t*)(base + 2 * b_ptr);
    double* d_scale = (double*)(base + 2 * b_ptr + b_idx);
    double* d_out = d_scale + d->N;
    ACES_CUDA(ctx, cudaMemcpyAsync(d_ptr, group_ptr, b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_boff, boff.data(), b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int32_t) * (size_t)nidx, cudaMemcpyHostToDevice, ctx->stream));
    if (gate_eigenvalues)
        ACES_CUDA(ctx, cudaMemcpyAsync(d_scale, gate_eigenvalues, sizeof(double) * (size_t)d->N, cudaMemcpyDefault, ctx->stream));
    precision_blocks_kernel<<<(unsigned)G, 64, 0, ctx->stream>>>(G, d_ptr, d_boff, d_idx, d->G, d->Np,
                                                                gate_eigenvalues ? d_scale : nullptr, d_out);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(blocks, d_out, sizeof(double) * (size_t)boff.back(), cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}

}  // extern "C"

// calc_gls / wls / ols_covariance on the device: *S_out is the N x N estimator covariance inside an
// Np x Np buffer of the design (d->G for GLS, d->Cn otherwise), exactly symmetric.
int ls_covariance_device(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* gate_eigenvalues,
                         const double* covariance_log_nzval, double** S_out) {
    d->est_ls_type = -1;
    const bool gls = ls_type == 2;
    ACES_TRY(ensure_workspaces(ctx, d, gls, gls, true, !gls));
    if (gls) ACES_TRY(build_tile_descs(ctx, d));
    FitVectors fv = fit_vectors(d);
    const int64_t Np = d->Np, N = d->N;
    std::vector<int64_t> kept(d->m.begin(), d->m.end());
    if (covariance_log_nzval != d->covval)
        ACES_CUDA(ctx, cudaMemcpyAsync(d->covval, covariance_log_nzval, sizeof(double) * (size_t)d->nnz_cov,
                                       cudaMemcpyDefault, ctx->stream));
    ACES_TRY(prepare_estimator(ctx, d, ls_type, d->covval, nullptr, kept, fv));
    // inv(G) = X' X with X = inv(L): lower tiles on the tensor pipe, then mirrored
    ACES_TRY(dense::trtri(ctx, Np, d->G, Np, d->dinvG, d->Xn, Np));
    ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Np, 1.0, d->Xn, Np, d->Xn, Np, 0.0, d->G, Np, 1));
    ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->G, Np));
    double* S = d->G;
    if (!gls) {
        // sandwich: est * cov_log * est' = H (A' W cov_log W A) H with H = inv(A' W A)
        // (src/merit.jl:584-596, 614-622); the middle matrix is assembled sparsely, W = f^2
        mul_vec_kernel<<<grid_for(d->Mp, 256), 256, 0, ctx->stream>>>(d->Mp, fv.f, fv.f, fv.v);
        ACES_CUDA(ctx, cudaMemsetAsync(d->Cn, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        sandwich_kernel<<<grid_for(N, 64), 64, 0, ctx->stream>>>(
            N, d->a_colptr, d->a_rowval, d->a_nzval, d->a_rowptr, d->a_colidx, d->a_vals, d->cov_colptr, d->cov_rowval,
            d->covval, d->blk_of_row, d->d_row0, d->d_prow0, nullptr, fv.v, d->Cn, Np);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 2;
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->Cn, Np));
        // Xn <- H * C ; Cn <- (H C) * H
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, d->G, Np, d->Cn, Np, 0.0, d->Xn, Np, 0));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, d->Xn, Np, d->G, Np, 0.0, d->Cn, Np, 1));
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->Cn, Np));
        S = d->Cn;
    }
    double* gdev = fv.t2;
    ACES_CUDA(ctx, cudaMemcpyAsync(gdev, gate_eigenvalues, sizeof(double) * (size_t)N, cudaMemcpyDefault, ctx->stream));
    scale_sym_kernel<<<148 * 8, 256, 0, ctx->stream>>>(N, S, Np, gdev);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    *S_out = S;
    return ACES_OK;
}

extern "C" {

int32_t aces_design_ls_covariance(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* gate_eigenvalues,
                                  const double* covariance_log_nzval, double* sigma, double* trace,
                                  double* trace_sq) {
    ACES_TRY(check_design(ctx, d, "design_ls_covariance"));
    ACES_CHECK(ctx, !d->sharded, "design_ls_covariance: the design is a tuple shard");
    ACES_CHECK(ctx, ls_type >= 0 && ls_type <= 2, "design_ls_covariance: ls_type must be 0 (OLS), 1 (WLS) or 2 (GLS)");
    ACES_CHECK(ctx, gate_eigenvalues && covariance_log_nzval, "design_ls_covariance: NULL argument");
    const int64_t Np = d->Np, N = d->N;
    double* S = nullptr;
    ACES_TRY(ls_covariance_device(ctx, d, ls_type, gate_eigenvalues, covariance_log_nzval, &S));
    double* mom = d->mom;
    moments_kernel<<<148 * 4, 256, 0, ctx->stream>>>(N, S, Np, mom);
    moments_final_kernel<<<1, 32, 0, ctx->stream>>>(148 * 4, mom);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    double hm[2] = {0.0, 0.0};
    ACES_CUDA(ctx, cudaMemcpyAsync(hm, mom, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sigma)
        ACES_CUDA(ctx, cudaMemcpy2DAsync(sigma, sizeof(double) * (size_t)N, S, sizeof(double) * (size_t)Np,
                                         sizeof(double) * (size_t)N, (size_t)N, cudaMemcpyDefault, ctx->stream));
    ACES_TRY(check_flags(ctx, d, "design_ls_covariance"));
    if (trace) *trace = hm[0];
    if (trace_sq) *trace_sq = hm[1];
    return ACES_OK;
}

}  // extern "C"

#include "merit_grad.cuh"
#include "merit_tail.cuh"
#include "project_full.cuh"
