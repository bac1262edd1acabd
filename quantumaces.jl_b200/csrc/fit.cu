// fit.cu -- least-squares fit of the gate eigenvalues and the estimator covariances (K2-K6).
//
// Reference functions replaced (paths relative to the QuantumACES.jl root):
//   calc_covariance / calc_covariance_log        src/merit.jl:206-291, 356-365     (K2)
//   sparse_covariance_inv_factor / _inv           src/merit.jl:378-448              (K3)
//   estimate_gate_eigenvalues                     src/estimate.jl:492-524           (K4/K5)
//   calc_precision_matrix                         src/merit.jl:754-770
//   calc_gls/wls/ols_covariance                   src/merit.jl:556-629              (K5/K6)
//
// Numerical method of the fit.  The reference solves min |F (A x - b)| with a sparse QR
// (SPQR).  Here the whitened matrix At = F*A is formed densely on the device, the Gram matrix
// G = At' At is accumulated on the FP64 tensor pipe (DMMA), factorised by a blocked Cholesky,
// and the normal-equation solution is refined with residuals r = bt - At x evaluated in FP64
// (corrected semi-normal equations), which restores QR-level accuracy for the condition
// numbers of ACES design matrices (checked against a QR solve to 1e-9 in tests/).
#include <algorithm>
#include <cmath>
#include <vector>

#include "dense.cuh"

namespace {

using dense::DB;
using dense::pad_up;

// ---- sparse -> dense ------------------------------------------------------------------------

// At[:, j] = sum_i A[i, j] * F[:, i]   (one CTA per column j of A; F == NULL: At = A)
__global__ void whiten_kernel(int64_t N, const int64_t* __restrict__ a_colptr,
                              const int64_t* __restrict__ a_rowval, const double* __restrict__ a_nzval,
                              const int64_t* __restrict__ f_colptr, const int64_t* __restrict__ f_rowval,
                              const double* __restrict__ f_nzval, double* __restrict__ At, int64_t ld) {
    const int64_t j = blockIdx.x;
    if (j >= N) return;
    double* col = At + j * ld;
    for (int64_t e = a_colptr[j]; e < a_colptr[j + 1]; ++e) {
        const int64_t i = a_rowval[e];
        const double a = a_nzval[e];
        if (f_colptr == nullptr) {
            if (threadIdx.x == 0) atomicAdd(&col[i], a);
        } else {
            for (int64_t f = f_colptr[i] + threadIdx.x; f < f_colptr[i + 1]; f += blockDim.x)
                atomicAdd(&col[f_rowval[f]], a * f_nzval[f]);
        }
    }
}
// bt = F * b (F == NULL: bt = b), b_i = -log(y_i)
__global__ void whiten_rhs_kernel(int64_t M, const double* __restrict__ y,
                                  const int64_t* __restrict__ f_colptr, const int64_t* __restrict__ f_rowval,
                                  const double* __restrict__ f_nzval, double* __restrict__ bt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double b = -log(y[i]);
        if (f_colptr == nullptr) {
            atomicAdd(&bt[i], b);
        } else {
            for (int64_t f = f_colptr[i]; f < f_colptr[i + 1]; ++f) atomicAdd(&bt[f_rowval[f]], b * f_nzval[f]);
        }
    }
}

// ---- dense matrix-vector products (memory bound) -------------------------------------------------

// y[j] (+)= sum_i A[i + j*ld] * x[i]   one warp per column
__global__ void gemv_t_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t ld,
                              const double* __restrict__ x, double* __restrict__ y, double alpha, double beta) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < n; j += nwarps) {
        const double* col = A + j * ld;
        double s = 0.0;
        for (int64_t i = lane; i < m; i += 32) s += col[i] * x[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[j] = alpha * s + (beta != 0.0 ? beta * y[j] : 0.0);
    }
}
// y[i] += alpha * sum_{j in chunk} A[i + j*ld] * x[j]    (y pre-scaled by the caller)
__global__ void gemv_n_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t ld,
                              const double* __restrict__ x, double* __restrict__ y, double alpha, int64_t jchunk) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t j0 = blockIdx.y * jchunk, j1 = min(j0 + jchunk, n);
    if (i >= m) return;
    double s = 0.0;
    for (int64_t j = j0; j < j1; ++j) s += A[i + j * ld] * x[j];
    atomicAdd(&y[i], alpha * s);
}
__global__ void scale_vec_kernel(int64_t n, double* y, double beta) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = beta == 0.0 ? 0.0 : beta * y[i];
}
__global__ void axpy_kernel(int64_t n, double a, const double* x, double* y) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] += a * x[i];
}
__global__ void exp_neg_kernel(int64_t n, const double* x, double* g, int constrain) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (constrain && v < 0.0) v = 0.0;  // src/estimate.jl:504-506
        g[i] = exp(-v);
    }
}
// S[i + j*lds] = d[i] * S[i + j*lds] * d[j]  (d == NULL: untouched), only i,j < n
__global__ void scale_sym_kernel(int64_t n, double* S, int64_t lds, const double* d, int inverse) {
    const int64_t total = n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % n, j = e / n;
        const double di = inverse ? 1.0 / d[i] : d[i], dj = inverse ? 1.0 / d[j] : d[j];
        S[i + j * lds] = S[i + j * lds] * (di * dj);  // (di*dj) commutes: the result stays exactly symmetric
    }
}

int gemv_t(aces_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t ld, const double* x, double* y,
           double alpha, double beta) {
    gemv_t_kernel<<<(unsigned)std::min<int64_t>((n + 7) / 8, 148 * 8), 256, 0, ctx->stream>>>(m, n, A, ld, x, y, alpha, beta);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}
int gemv_n(aces_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t ld, const double* x, double* y,
           double alpha, double beta) {
    scale_vec_kernel<<<(unsigned)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, ctx->stream>>>(m, y, beta);
    const int64_t jchunk = 256;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)((n + jchunk - 1) / jchunk));
    gemv_n_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, A, ld, x, y, alpha, jchunk);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return ACES_OK;
}

struct SparseDev {  // CSC on the device, 0-based
    int64_t* colptr = nullptr;
    int64_t* rowval = nullptr;
    double* nzval = nullptr;
};

// Uploads a CSC matrix given with index base `base` (0: C, 1: Julia) into `buf`.
int upload_csc(aces_ctx* ctx, DevBuf& buf, int64_t ncols, const int64_t* colptr, const int64_t* rowval,
               const double* nzval, int base, SparseDev* out) {
    const int64_t nnz = colptr[ncols] - base;
    std::vector<int64_t> cp((size_t)ncols + 1), rv((size_t)std::max<int64_t>(nnz, 1));
    for (int64_t j = 0; j <= ncols; ++j) cp[(size_t)j] = colptr[j] - base;
    for (int64_t e = 0; e < nnz; ++e) rv[(size_t)e] = rowval[e] - base;
    const size_t b_cp = sizeof(int64_t) * ((size_t)ncols + 1), b_rv = sizeof(int64_t) * (size_t)std::max<int64_t>(nnz, 1),
                 b_nz = sizeof(double) * (size_t)std::max<int64_t>(nnz, 1);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    ACES_TRY(aces_reserve_dev(ctx, buf, al(b_cp) + al(b_rv) + al(b_nz)));
    uint8_t* d = (uint8_t*)buf.ptr;
    out->colptr = (int64_t*)d;
    out->rowval = (int64_t*)(d + al(b_cp));
    out->nzval = (double*)(d + al(b_cp) + al(b_rv));
    ACES_CUDA(ctx, cudaMemcpyAsync(out->colptr, cp.data(), b_cp, cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
        ACES_CUDA(ctx, cudaMemcpyAsync(out->rowval, rv.data(), sizeof(int64_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        ACES_CUDA(ctx, cudaMemcpyAsync(out->nzval, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the staging vectors die at return: wait for the copies
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}

// Workspace slots of ctx->d_work
enum { W_A = 0, W_F = 1, W_AT = 2, W_G = 3, W_X = 4, W_DINV = 5, W_VEC = 6, W_AUX = 7 };

struct LsqState {
    int64_t M, N, Mp, Np;
    double* At;    // Mp x Np whitened design matrix
    double* G;     // Np x Np Gram matrix -> its Cholesky factor (lower)
    double* X;     // Np x Np inverse of the factor
    double* dinv;  // diagonal-block inverses
    double* bt;    // Mp
    double* g;     // Np   At' bt
    double* x;     // Np   solution
    double* t1;    // max(Mp, Np) scratch
    double* t2;    // Np scratch
    double* t3;    // Np scratch
    int* flag;
};

// Builds At = F*A (dense), bt = F*(-log y), G = At'At, factorises it and inverts the factor.
int lsq_setup(aces_ctx* ctx, int64_t M, int64_t N, const int64_t* A_colptr, const int64_t* A_rowval,
              const double* A_nzval, const int64_t* F_colptr, const int64_t* F_rowval, const double* F_nzval,
              const double* y, int base, LsqState* st, bool with_inverse) {
    ACES_CHECK(ctx, M >= 1 && N >= 1, "lsq: M=%lld N=%lld", (long long)M, (long long)N);
    ACES_CHECK(ctx, A_colptr && A_rowval && A_nzval, "lsq: NULL design matrix");
    const int64_t Mp = pad_up(M), Np = pad_up(N);
    st->M = M; st->N = N; st->Mp = Mp; st->Np = Np;
    SparseDev A, F;
    ACES_TRY(upload_csc(ctx, ctx->d_work[W_A], N, A_colptr, A_rowval, A_nzval, base, &A));
    const bool hasF = F_colptr != nullptr;
    if (hasF) ACES_TRY(upload_csc(ctx, ctx->d_work[W_F], M, F_colptr, F_rowval, F_nzval, base, &F));
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_AT], sizeof(double) * (size_t)Mp * Np));
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_G], sizeof(double) * (size_t)Np * Np));
    if (with_inverse) ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_X], sizeof(double) * (size_t)Np * Np));
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_DINV], sizeof(double) * (size_t)Np * DB));
    const size_t vec = (size_t)std::max(Mp, Np);
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_VEC], sizeof(double) * (vec * 7 + 64)));
    st->At = (double*)ctx->d_work[W_AT].ptr;
    st->G = (double*)ctx->d_work[W_G].ptr;
    st->X = (double*)ctx->d_work[W_X].ptr;
    st->dinv = (double*)ctx->d_work[W_DINV].ptr;
    double* v = (double*)ctx->d_work[W_VEC].ptr;
    st->bt = v; st->g = v + vec; st->x = v + 2 * vec; st->t1 = v + 3 * vec; st->t2 = v + 4 * vec;
    st->t3 = v + 5 * vec;
    double* ydev = v + 6 * vec;
    st->flag = (int*)(v + 7 * vec);
    ACES_CUDA(ctx, cudaMemsetAsync(st->At, 0, sizeof(double) * (size_t)Mp * Np, ctx->stream));
    ACES_CUDA(ctx, cudaMemsetAsync(v, 0, sizeof(double) * (vec * 7 + 64), ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(ydev, y, sizeof(double) * (size_t)M, cudaMemcpyDefault, ctx->stream));
    whiten_kernel<<<(unsigned)N, 128, 0, ctx->stream>>>(N, A.colptr, A.rowval, A.nzval, hasF ? F.colptr : nullptr,
                                                        hasF ? F.rowval : nullptr, hasF ? F.nzval : nullptr, st->At, Mp);
    whiten_rhs_kernel<<<(unsigned)std::min<int64_t>((M + 127) / 128, 2048), 128, 0, ctx->stream>>>(
        M, ydev, hasF ? F.colptr : nullptr, hasF ? F.rowval : nullptr, hasF ? F.nzval : nullptr, st->bt);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    // Gram matrix on the FP64 tensor pipe (lower tiles), unit diagonal on the padding
    if (ctx->timing) ACES_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Mp, 1.0, st->At, Mp, st->At, Mp, 0.0, st->G, Np, 1));
    if (ctx->timing) {
        ACES_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
        ctx->timed = 1;
    }
    ACES_TRY(dense::fill_diag_range(ctx, st->G, Np, N, Np, 1.0));
    ACES_TRY(gemv_t(ctx, Mp, Np, st->At, Mp, st->bt, st->g, 1.0, 0.0));
    ACES_TRY(dense::get_diag(ctx, Np, st->G, Np, st->t2));
    // A fit eliminates the columns that fail the pivot test (basic solution of a rank-deficient
    // problem, as the reference's sparse QR returns; the count is read back with the result).  The
    // inverse of a singular Gram matrix is an error, as inv(cholesky(.)) is in the reference.
    ACES_TRY(dense::potrf(ctx, Np, st->G, Np, st->dinv, st->flag, st->t2, 1e-13, with_inverse ? 0 : 1));
    if (with_inverse) {
        int flag = 0;
        ACES_CUDA(ctx, cudaMemcpyAsync(&flag, st->flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (flag != 0)
            return aces_fail(ctx, ACES_E_NOT_PD,
                             "lsq: the Gram matrix is not positive definite (diagonal block %d): the whitened "
                             "design matrix is rank deficient", flag - 1);
        ACES_TRY(dense::trtri(ctx, Np, st->G, Np, st->dinv, st->X, Np));
    }
    return ACES_OK;
}

// out = inv(G) v = inv(L') inv(L) v by two blocked triangular solves (v is kept)
int lsq_apply_inverse(aces_ctx* ctx, const LsqState& st, const double* v, double* out) {
    ACES_CUDA(ctx, cudaMemcpyAsync(st.t2, v, sizeof(double) * (size_t)st.Np, cudaMemcpyDeviceToDevice, ctx->stream));
    ACES_TRY(dense::trsv_lower(ctx, st.Np, st.G, st.Np, st.dinv, st.t2, st.t3));
    ACES_TRY(dense::trsv_lower_t(ctx, st.Np, st.G, st.Np, st.dinv, st.t3, out));
    return ACES_OK;
}

int lsq_solve(aces_ctx* ctx, const LsqState& st, int refine) {
    ACES_TRY(lsq_apply_inverse(ctx, st, st.g, st.x));
    for (int it = 0; it < refine; ++it) {
        // r = bt - At x ; g2 = At' r ; x += inv(G) g2
        ACES_CUDA(ctx, cudaMemcpyAsync(st.t1, st.bt, sizeof(double) * (size_t)st.Mp, cudaMemcpyDeviceToDevice, ctx->stream));
        ACES_TRY(gemv_n(ctx, st.Mp, st.Np, st.At, st.Mp, st.x, st.t1, -1.0, 1.0));
        ACES_TRY(gemv_t(ctx, st.Mp, st.Np, st.At, st.Mp, st.t1, st.g, 1.0, 0.0));
        ACES_TRY(lsq_apply_inverse(ctx, st, st.g, st.t1));
        axpy_kernel<<<(unsigned)std::min<int64_t>((st.Np + 255) / 256, 1024), 256, 0, ctx->stream>>>(st.Np, 1.0, st.t1, st.x);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return ACES_OK;
}

}  // namespace

extern "C" {

int32_t aces_estimate_gate_eigenvalues(aces_ctx* ctx, int64_t M, int64_t N, const int64_t* A_colptr,
                                       const int64_t* A_rowval, const double* A_nzval,
                                       const int64_t* F_colptr, const int64_t* F_rowval,
                                       const double* F_nzval, int32_t index_base,
                                       const double* est_eigenvalues, int32_t constrain,
                                       double* gate_eigenvalues) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, est_eigenvalues && gate_eigenvalues, "estimate_gate_eigenvalues: NULL argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "estimate_gate_eigenvalues: index_base must be 0 or 1");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    LsqState st;
    ACES_TRY(lsq_setup(ctx, M, N, A_colptr, A_rowval, A_nzval, F_colptr, F_rowval, F_nzval, est_eigenvalues,
                       index_base, &st, false));
    ACES_TRY(lsq_solve(ctx, st, 2));
    exp_neg_kernel<<<(unsigned)std::min<int64_t>((N + 255) / 256, 1024), 256, 0, ctx->stream>>>(N, st.x, st.t1, constrain);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(gate_eigenvalues, st.t1, sizeof(double) * (size_t)N, cudaMemcpyDefault, ctx->stream));
    int dead = 0;
    ACES_CUDA(ctx, cudaMemcpyAsync(&dead, st.flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->rank_deficiency = dead;
    return ACES_OK;
}

int32_t aces_gram_inverse(aces_ctx* ctx, int64_t M, int64_t N, const int64_t* A_colptr,
                          const int64_t* A_rowval, const double* A_nzval, const int64_t* F_colptr,
                          const int64_t* F_rowval, const double* F_nzval, int32_t index_base,
                          const double* scale, int32_t scale_inverse, int32_t want_gram, double* out) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, out != nullptr, "gram_inverse: NULL output");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "gram_inverse: index_base must be 0 or 1");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    LsqState st;
    const int64_t Np = pad_up(N), Mp = pad_up(M);
    if (want_gram) {
        // precision-matrix form: scale * (At' At) * scale, no factorisation needed
        std::vector<double> ones((size_t)M, 1.0);
        // lsq_setup factorises G in place, so build the pieces by hand here
        SparseDev A, F;
        ACES_TRY(upload_csc(ctx, ctx->d_work[W_A], N, A_colptr, A_rowval, A_nzval, index_base, &A));
        const bool hasF = F_colptr != nullptr;
        if (hasF) ACES_TRY(upload_csc(ctx, ctx->d_work[W_F], M, F_colptr, F_rowval, F_nzval, index_base, &F));
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_AT], sizeof(double) * (size_t)Mp * Np));
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_G], sizeof(double) * (size_t)Np * Np));
        double* At = (double*)ctx->d_work[W_AT].ptr;
        double* G = (double*)ctx->d_work[W_G].ptr;
        ACES_CUDA(ctx, cudaMemsetAsync(At, 0, sizeof(double) * (size_t)Mp * Np, ctx->stream));
        whiten_kernel<<<(unsigned)N, 128, 0, ctx->stream>>>(N, A.colptr, A.rowval, A.nzval, hasF ? F.colptr : nullptr,
                                                            hasF ? F.rowval : nullptr, hasF ? F.nzval : nullptr, At, Mp);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Mp, 1.0, At, Mp, At, Mp, 0.0, G, Np, 1));
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, G, Np));
        st.G = G;
        st.X = G;
    } else {
        std::vector<double> ones((size_t)M, 1.0);
        ACES_TRY(lsq_setup(ctx, M, N, A_colptr, A_rowval, A_nzval, F_colptr, F_rowval, F_nzval, ones.data(),
                           index_base, &st, true));
        // inv(G) = X' X  (lower tiles on the tensor pipe, then mirrored) -> reuse the G buffer
        ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Np, 1.0, st.X, Np, st.X, Np, 0.0, st.G, Np, 1));
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, st.G, Np));
    }
    double* S = st.G;
    if (scale) {
        ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[W_AUX], sizeof(double) * (size_t)N));
        double* d = (double*)ctx->d_work[W_AUX].ptr;
        ACES_CUDA(ctx, cudaMemcpyAsync(d, scale, sizeof(double) * (size_t)N, cudaMemcpyDefault, ctx->stream));
        scale_sym_kernel<<<148 * 8, 256, 0, ctx->stream>>>(N, S, Np, d, scale_inverse);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    ACES_CUDA(ctx, cudaMemcpy2DAsync(out, sizeof(double) * (size_t)N, S, sizeof(double) * (size_t)Np,
                                     sizeof(double) * (size_t)N, (size_t)N, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACES_OK;
}

}  // extern "C"
