// merit_grad.cuh -- figure of merit of a design and its gradient with respect to the logarithms of the
// shot weights: calc_gls_merit_grad_log / calc_wls_merit_grad_log /
// calc_ols_merit_grad_log of src/optimise_weights.jl:103-152, 327-391, 563-606, the inner step of
// gls_ / wls_ / ols_optimise_weights (one call per gradient-descent step, up to 1000 steps).
//
// Included at the end of design.cu (it uses that file's internal helpers).
//
// The reference forms dense M x N and M x M intermediates (gls_matrix, wls_gram, wls_sq_gram,
// ols_gram_covariance) and then keeps only per-tuple sums of one diagonal.  Those sums are inner
// products with N x N matrices, so nothing larger than N x N is built here.  With
//   A = design matrix (rows a_i), Ou = unweighted covariance (block diagonal by tuple),
//   s_t = (sum_u w_u tau_u) / w_t the shot-weight factor of tuple t, O = Ou * diag(s),
//   TT = T'T for the gate transform T (N' x N),
// and per estimator
//   GLS: H = inv(A' inv(O) A),            B_t = A_t' inv(O_t) A_t
//   WLS: H = inv(A' W A), W = inv(diag O), Bm = A' W O W A
//   OLS: H = inv(A' A),                   Bm = A' O A
//   P = H TT H,   P2 = P Bm P (GLS: P TT H),   Q1 = P Bm H,   Q2 = P2 Bm H,
// the quantities of the reference are
//   tr(Sigma) = <P, Bm>, tr(Sigma^2) = <P2, Bm>                         (GLS: Bm = sum_t B_t = inv(H))
//   GLS  sigma_tr_partial[t]    = (1 / s_t) <P, B_t>,  sigma_sq_tr_partial[t] = (2 / s_t) <P2, B_t>
//   OLS  sigma_tr_partial[t]    = sum_{(j,i) in block t} Ou[j,i] a_i' P a_j            (P2: twice that)
//   WLS  sigma_tr_partial[t]    = sum_{(j,i)} w_i w_j Ou[j,i] a_i' P a_j
//                                + sum_i 2 Ou[i,i] w_i^2 (a_i' Q1 a_i - sum_j w_j O[j,i] a_i' P a_j)
//                                                                       (P2, Q2: twice that)
// (derivation: wls_gram[i,j] = w_i w_j a_i' P a_j, wls_sq_gram[i,j] = w_i w_j a_i' P2 a_j, and the
// commutant of src/optimise_weights.jl:351-352 contracted against them).  tests/test_optimise_gpu.py
// checks all of it against a dense transcription of the reference's formulas.
//
// Cost per call at rotated d = 9: the Gram matrix, its Cholesky factor and inverse, then two (GLS),
// three (OLS) or six (WLS) N^3 DMMA GEMMs; the gathers over the covariance pattern are negligible.
#pragma once

namespace {

// Ou values -> O = Ou * s_t (column scaling; the pattern is block diagonal, so row and column share the tuple)
__global__ void scale_by_tuple_kernel(int64_t M, const int64_t* __restrict__ colptr, const int32_t* __restrict__ blk_of_row,
                                      const double* __restrict__ sf, const double* __restrict__ in, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double s = sf[blk_of_row[i]];
        for (int64_t e = colptr[i]; e < colptr[i + 1]; ++e) out[e] = in[e] * s;
    }
}

// per row: the diagonal entry of Ou and the WLS weight 1 / (Ou[i,i] s_t) (1 for OLS)
__global__ void row_weights_kernel(int64_t M, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                   const double* __restrict__ cu, const int32_t* __restrict__ blk_of_row,
                                   const double* __restrict__ sf, int wls, double* __restrict__ udiag, double* __restrict__ wvec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        double u = 0.0;
        for (int64_t e = colptr[i]; e < colptr[i + 1]; ++e)
            if (rowval[e] == i) u = cu[e];
        udiag[i] = u;
        wvec[i] = wls ? 1.0 / (u * sf[blk_of_row[i]]) : 1.0;
    }
}

// U = S * X for a sparse symmetric S (CSC == CSR) and a dense X (n columns of length ldx)
__global__ void spmm_sym_kernel(int64_t N, int64_t ncols, const int64_t* __restrict__ ptr, const int64_t* __restrict__ idx,
                                const double* __restrict__ val, const double* __restrict__ X, int64_t ldx,
                                double* __restrict__ U, int64_t ldu) {
    const int64_t total = N * ncols;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e % N, c = e / N;
        double acc = 0.0;
        for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * X[idx[k] + c * ldx];
        U[r + c * ldu] = acc;
    }
}

__device__ __forceinline__ double bilinear_sym(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                               const double* __restrict__ vals, int64_t i, int64_t j,
                                               const double* __restrict__ X, int64_t ldx) {
    double bf = 0.0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const int64_t cp = colidx[p];
        double inner = 0.0;
        for (int64_t q = rowptr[j]; q < rowptr[j + 1]; ++q) {
            const int64_t cq = colidx[q];
            inner += vals[q] * X[(cp > cq ? cp : cq) + (cp > cq ? cq : cp) * ldx];  // X symmetric: lower triangle
        }
        bf += vals[p] * inner;
    }
    return bf;
}

// One thread per row i of the design: the OLS / WLS sums over the covariance-pattern column i.
//   out[2 i]     contribution of row i to sigma_tr_partial (X = P) or half of sigma_sq_tr_partial (X = P2)
//   out[2 i + 1] contribution to tr(Sigma) (X = P) or tr(Sigma^2) (X = P2)
// Q (WLS only) = P Bm H or P2 Bm H, full storage.
__global__ void pattern_rows_kernel(int64_t M, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                    const double* __restrict__ cu, const int64_t* __restrict__ a_rowptr,
                                    const int32_t* __restrict__ a_colidx, const double* __restrict__ a_vals,
                                    const int32_t* __restrict__ blk_of_row, const double* __restrict__ sf,
                                    const double* __restrict__ udiag, const double* __restrict__ wvec, int wls,
                                    const double* __restrict__ X, int64_t ldx, const double* __restrict__ Q, int64_t ldq,
                                    double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const double s = sf[blk_of_row[i]];
        const double wi = wvec[i], uii = udiag[i];
        double acc_p = 0.0, acc_tr = 0.0, acc_c = 0.0;
        for (int64_t e = colptr[i]; e < colptr[i + 1]; ++e) {
            const int64_t j = rowval[e];
            const double u = cu[e];
            const double bf = bilinear_sym(a_rowptr, a_colidx, a_vals, i, j, X, ldx);
            const double wj = wvec[j];
            acc_p += wi * wj * u * bf;          // (wls_gram * Ou)[i, i]; weights are 1 for OLS
            acc_tr += wi * wj * (u * s) * bf;   // <X, Bm>
            acc_c += wj * (u * s) * bf;         // sum_j w_j O[j, i] a_i' X a_j
        }
        if (wls) {
            double qf = 0.0;
            for (int64_t p = a_rowptr[i]; p < a_rowptr[i + 1]; ++p) {
                double inner = 0.0;
                for (int64_t q = a_rowptr[i]; q < a_rowptr[i + 1]; ++q) inner += Q[a_colidx[p] + (int64_t)a_colidx[q] * ldq] * a_vals[q];
                qf += a_vals[p] * inner;
            }
            acc_p += 2.0 * uii * wi * wi * (qf - acc_c);
        }
        out[2 * i] = acc_p;
        out[2 * i + 1] = acc_tr;
    }
}

// out[2 t + k] = sum over the rows of tuple t of in[2 i + k], in a fixed order (one CTA per tuple)
__global__ void tuple_sums_kernel(const int64_t* __restrict__ row0, const double* __restrict__ in, double* __restrict__ out) {
    const int t = blockIdx.x;
    double a = 0.0, b = 0.0;
    for (int64_t i = row0[t] + threadIdx.x; i < row0[t + 1]; i += blockDim.x) {
        a += in[2 * i];
        b += in[2 * i + 1];
    }
    __shared__ double sa[256], sb[256];
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sa[threadIdx.x] += sa[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * t] = sa[0];
        out[2 * t + 1] = sb[0];
    }
}

// GLS: <X, B_t> over the column window of tuple t.  Gt holds the lower triangle of the compact block
// (cp x cp, column-major), cols its global columns (ascending), X the lower triangle of a symmetric
// matrix.  CTA (cx, t) takes the window columns b = cx, cx + gridDim.x, ...; parts[t * gridDim.x + cx].
__global__ void window_inner_kernel(const int64_t* __restrict__ ct, const int64_t* __restrict__ cpt,
                                    const int64_t* __restrict__ cols_off, const int64_t* __restrict__ gt_off,
                                    const int64_t* __restrict__ w_cols, const double* __restrict__ Gt,
                                    const double* __restrict__ X, int64_t ldx, double* __restrict__ parts) {
    const int t = blockIdx.y;
    const int64_t c = ct[t], cp = cpt[t];
    const int64_t* cols = w_cols + cols_off[t];
    const double* G = Gt + gt_off[t];
    double acc = 0.0;
    for (int64_t b = blockIdx.x; b < c; b += gridDim.x) {
        const int64_t gb = cols[b];
        for (int64_t a = b + threadIdx.x; a < c; a += blockDim.x) {
            const double v = G[a + b * cp] * X[cols[a] + gb * ldx];
            acc += (a == b) ? v : 2.0 * v;
        }
    }
    __shared__ double sa[256];
    sa[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sa[threadIdx.x] += sa[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) parts[(int64_t)t * gridDim.x + blockIdx.x] = sa[0];
}
__global__ void window_parts_kernel(int nparts, const double* __restrict__ parts, double* __restrict__ out) {
    const int t = blockIdx.x;
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < nparts; ++i) a += parts[(int64_t)t * nparts + i];
        out[t] = a;
    }
}

}  // namespace

extern "C" {

int32_t aces_design_merit_grad(aces_ctx* ctx, aces_design* d, int32_t ls_type, const double* covariance_log_unweighted_nzval,
                               const double* shot_weights, const double* tuple_times, int64_t n_transformed,
                               const int64_t* tt_colptr, const int64_t* tt_rowval, const double* tt_nzval,
                               int32_t index_base, double* merit, double* merit_grad_log, double* sigma_tr,
                               double* sigma_sq_tr) {
    ACES_TRY(check_design(ctx, d, "design_merit_grad"));
    ACES_CHECK(ctx, !d->sharded, "design_merit_grad: the design is a tuple shard");
    ACES_CHECK(ctx, ls_type >= 0 && ls_type <= 2, "design_merit_grad: ls_type must be 0 (OLS), 1 (WLS) or 2 (GLS)");
    ACES_CHECK(ctx, covariance_log_unweighted_nzval && shot_weights && tuple_times && tt_colptr && tt_rowval && tt_nzval,
               "design_merit_grad: NULL argument");
    ACES_CHECK(ctx, n_transformed >= 1 && (index_base == 0 || index_base == 1), "design_merit_grad: n_transformed=%lld index_base=%d",
               (long long)n_transformed, index_base);
    const int T = d->T;
    const int64_t N = d->N, Np = d->Np, M = d->M;
    const bool gls = ls_type == 2, wls = ls_type == 1;
    double times_factor = 0.0;
    for (int t = 0; t < T; ++t) {
        ACES_CHECK(ctx, shot_weights[t] > 0.0, "design_merit_grad: shot weight %d is not positive", t);
        times_factor += shot_weights[t] * tuple_times[t];
    }
    std::vector<double> sf((size_t)T);
    for (int t = 0; t < T; ++t) sf[(size_t)t] = times_factor * (1.0 / shot_weights[t]);   // get_shot_weights_factor
    // T'T (symmetric) as 0-based CSC, validated
    const int64_t tt_nnz = tt_colptr[N] - index_base;
    ACES_CHECK(ctx, tt_colptr[0] == index_base && tt_nnz >= 0, "design_merit_grad: malformed transform colptr");
    std::vector<int64_t> h_ptr((size_t)N + 1), h_idx((size_t)tt_nnz);
    for (int64_t c = 0; c <= N; ++c) h_ptr[(size_t)c] = tt_colptr[c] - index_base;
    for (int64_t e = 0; e < tt_nnz; ++e) {
        h_idx[(size_t)e] = tt_rowval[e] - index_base;
        ACES_CHECK(ctx, h_idx[(size_t)e] >= 0 && h_idx[(size_t)e] < N, "design_merit_grad: transform row index out of range");
    }
    d->est_ls_type = -1;
    // the windowed Gram path also leaves the per-tuple blocks B_t behind, which the GLS sums need
    const bool was_windowed = d->windowed;
    if (gls) d->windowed = true;
    struct Restore {
        aces_design* d; bool w;
        ~Restore() { d->windowed = w; }
    } restore{d, was_windowed};
    ACES_TRY(ensure_workspaces(ctx, d, gls, gls, true, true));
    if (gls) ACES_TRY(build_tile_descs(ctx, d));
    if (!d->W3) ACES_TRY(dev_alloc(ctx, d, &d->W3, (size_t)Np * Np));
    if (wls && !d->W4) ACES_TRY(dev_alloc(ctx, d, &d->W4, (size_t)Np * Np));
    if (!d->mg_rows) {
        ACES_TRY(dev_alloc(ctx, d, &d->mg_rows, (size_t)(4 * M + 2 * M + 64 * T + 8 * T + 16)));
        ACES_TRY(dev_alloc(ctx, d, &d->mg_cu, (size_t)d->nnz_cov));
    }
    // small inputs through one staging buffer of the context
    const size_t tt_bytes = sizeof(int64_t) * (h_ptr.size() + h_idx.size()) + sizeof(double) * (size_t)tt_nnz;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[2], tt_bytes + sizeof(double) * (size_t)T + 64));
    uint8_t* base = (uint8_t*)ctx->d_work[2].ptr;
    int64_t* tt_ptr_d = (int64_t*)base;
    int64_t* tt_idx_d = tt_ptr_d + h_ptr.size();
    double* tt_val_d = (double*)(tt_idx_d + h_idx.size());
    double* sf_d = tt_val_d + tt_nnz;
    ACES_CUDA(ctx, cudaMemcpyAsync(tt_ptr_d, h_ptr.data(), sizeof(int64_t) * h_ptr.size(), cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(tt_idx_d, h_idx.data(), sizeof(int64_t) * h_idx.size(), cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(tt_val_d, tt_nzval, sizeof(double) * (size_t)tt_nnz, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(sf_d, sf.data(), sizeof(double) * (size_t)T, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d->mg_cu, covariance_log_unweighted_nzval, sizeof(double) * (size_t)d->nnz_cov,
                                   cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
    double* rows_out = d->mg_rows;            // [2 M]
    double* udiag = rows_out + 2 * M;         // [M]
    double* wvec = udiag + M;                 // [M]
    double* parts = wvec + M + 2 * M;         // [64 T]
    double* tsum = parts + 64 * T;            // [8 T]
    FitVectors fv = fit_vectors(d);
    std::vector<int64_t> kept(d->m.begin(), d->m.end());
    // O = Ou * s_t on the design's pattern, then the estimator: block factors / weights, Gram matrix, Cholesky
    scale_by_tuple_kernel<<<grid_for(M, 256), 256, 0, ctx->stream>>>(M, d->cov_colptr, d->blk_of_row, sf_d, d->mg_cu, d->covval);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_TRY(prepare_estimator(ctx, d, ls_type, d->covval, nullptr, kept, fv));
    // H = inv(G) = X'X with X = inv(L)
    ACES_TRY(dense::trtri(ctx, Np, d->G, Np, d->dinvG, d->Xn, Np));
    ACES_TRY(dense::gemm(ctx, dense::GEMM_TN, Np, Np, Np, 1.0, d->Xn, Np, d->Xn, Np, 0.0, d->G, Np, 1));
    ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->G, Np));
    // the padding block of G is the identity: keep it out of the products
    ACES_TRY(dense::fill_diag_range(ctx, d->G, Np, N, Np, 0.0));
    double* H = d->G;
    auto spmm = [&](const double* X, double* U) -> int {   // U = TT * X (padding rows / columns zero)
        ACES_CUDA(ctx, cudaMemsetAsync(U, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        spmm_sym_kernel<<<148 * 8, 256, 0, ctx->stream>>>(N, N, tt_ptr_d, tt_idx_d, tt_val_d, X, Np, U, Np);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        return ACES_OK;
    };
    std::vector<double> h1((size_t)2 * T, 0.0), h2((size_t)2 * T, 0.0);   // per tuple: (partial, trace) for P and P2
    if (gls) {
        double* U = d->Xn;
        double* P = d->Cn;
        double* P2 = d->W3;
        ACES_TRY(spmm(H, U));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, H, Np, U, Np, 0.0, P, Np, 1));
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, P, Np));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, P, Np, U, Np, 0.0, P2, Np, 1));
        if (!d->mg_win) {
            std::vector<int64_t> w;
            w.insert(w.end(), d->ct.begin(), d->ct.end());
            w.insert(w.end(), d->cpt.begin(), d->cpt.end());
            w.insert(w.end(), d->cols_off.begin(), d->cols_off.begin() + T);
            w.insert(w.end(), d->gt_off.begin(), d->gt_off.begin() + T);
            ACES_TRY(dev_upload(ctx, d, &d->mg_win, w));
        }
        for (int pass = 0; pass < 2; ++pass) {
            window_inner_kernel<<<dim3(64, (unsigned)T), 256, 0, ctx->stream>>>(d->mg_win, d->mg_win + T, d->mg_win + 2 * T,
                                                                               d->mg_win + 3 * T, d->w_cols, d->Gt,
                                                                               pass ? P2 : P, Np, parts);
            window_parts_kernel<<<(unsigned)T, 32, 0, ctx->stream>>>(64, parts, tsum + pass * T);
            ACES_CUDA(ctx, cudaGetLastError());
            ctx->launches += 2;
        }
        std::vector<double> q((size_t)2 * T);
        ACES_CUDA(ctx, cudaMemcpyAsync(q.data(), tsum, sizeof(double) * q.size(), cudaMemcpyDeviceToHost, ctx->stream));
        ACES_TRY(check_flags(ctx, d, "design_merit_grad"));
        for (int t = 0; t < T; ++t) {
            const double sfi = 1.0 / sf[(size_t)t];
            h1[(size_t)2 * t] = sfi * q[(size_t)t];
            h1[(size_t)2 * t + 1] = q[(size_t)t];
            h2[(size_t)2 * t] = sfi * q[(size_t)T + t];
            h2[(size_t)2 * t + 1] = q[(size_t)T + t];
        }
    } else {
        row_weights_kernel<<<grid_for(M, 256), 256, 0, ctx->stream>>>(M, d->cov_colptr, d->cov_rowval, d->mg_cu, d->blk_of_row,
                                                                     sf_d, wls ? 1 : 0, udiag, wvec);
        // Bm = A' W O W A (W = f^2 on the padded row space), assembled sparsely into Cn
        mul_vec_kernel<<<grid_for(d->Mp, 256), 256, 0, ctx->stream>>>(d->Mp, fv.f, fv.f, fv.v);
        ACES_CUDA(ctx, cudaMemsetAsync(d->Cn, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        sandwich_kernel<<<grid_for(N, 64), 64, 0, ctx->stream>>>(
            N, d->a_colptr, d->a_rowval, d->a_nzval, d->a_rowptr, d->a_colidx, d->a_vals, d->cov_colptr, d->cov_rowval,
            d->covval, d->blk_of_row, d->d_row0, d->d_prow0, nullptr, fv.v, d->Cn, Np);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 3;
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->Cn, Np));
        double* Bm = d->Cn;
        double* B1 = d->Xn;
        double* P = d->W3;
        auto pattern_pass = [&](const double* X, const double* Q, std::vector<double>& h) -> int {
            pattern_rows_kernel<<<grid_for(M, 128), 128, 0, ctx->stream>>>(
                M, d->cov_colptr, d->cov_rowval, d->mg_cu, d->a_rowptr, d->a_colidx, d->a_vals, d->blk_of_row, sf_d, udiag, wvec,
                wls ? 1 : 0, X, Np, Q, Np, rows_out);
            tuple_sums_kernel<<<(unsigned)T, 256, 0, ctx->stream>>>(d->d_row0, rows_out, tsum);
            ACES_CUDA(ctx, cudaGetLastError());
            ctx->launches += 2;
            ACES_CUDA(ctx, cudaMemcpyAsync(h.data(), tsum, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
            ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return ACES_OK;
        };
        ACES_TRY(spmm(H, B1));                                                                               // U = TT H
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, H, Np, B1, Np, 0.0, P, Np, 1));           // P = H U
        ACES_TRY(dense::symmetrize_from_lower(ctx, Np, P, Np));
        ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, P, Np, Bm, Np, 0.0, B1, Np, 0));          // Z = P Bm
        if (wls) {
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, B1, Np, H, Np, 0.0, d->W4, Np, 0));   // Q1 = Z H
            ACES_TRY(pattern_pass(P, d->W4, h1));
        } else {
            ACES_TRY(pattern_pass(P, nullptr, h1));
        }
        double* P2 = wls ? d->W4 : d->Xn;
        if (wls) {
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, B1, Np, P, Np, 0.0, P2, Np, 1));      // P2 = Z P
            ACES_TRY(dense::symmetrize_from_lower(ctx, Np, P2, Np));
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, P2, Np, Bm, Np, 0.0, B1, Np, 0));     // Z2 = P2 Bm
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, B1, Np, H, Np, 0.0, P, Np, 0));       // Q2 = Z2 H (P is dead)
            ACES_TRY(pattern_pass(P2, P, h2));
        } else {
            // OLS: P2 = Z P needs a buffer other than Z's: Z lives in Xn, P in W3, Bm in Cn (dead after Z)
            ACES_TRY(dense::gemm(ctx, dense::GEMM_NN, Np, Np, Np, 1.0, B1, Np, P, Np, 0.0, d->Cn, Np, 1));   // P2 = Z P
            ACES_TRY(pattern_pass(d->Cn, nullptr, h2));
        }
        ACES_TRY(check_flags(ctx, d, "design_merit_grad"));
    }
    // ---- host: src/optimise_weights.jl:128-151 (the tail common to the three estimators) -------------------
    double tr = 0.0, sq = 0.0;
    std::vector<double> p1((size_t)T), p2((size_t)T);
    for (int t = 0; t < T; ++t) {
        tr += h1[(size_t)2 * t + 1];
        sq += h2[(size_t)2 * t + 1];
        p1[(size_t)t] = h1[(size_t)2 * t];
        p2[(size_t)t] = 2.0 * h2[(size_t)2 * t];
    }
    const double Nn = (double)n_transformed;
    double s1 = 0.0, s2 = 0.0;
    for (int t = 0; t < T; ++t) {
        s1 += p1[(size_t)t] / shot_weights[t];
        s2 += p2[(size_t)t] / shot_weights[t];
    }
    std::vector<double> mg((size_t)T);
    const double c1 = ((tr * tr / 2.0) + (3.0 * sq / 8.0)) / std::pow(tr, 2.5), c2 = 1.0 / (4.0 * std::pow(tr, 1.5));
    for (int t = 0; t < T; ++t) {
        const double local = -times_factor / (shot_weights[t] * shot_weights[t]);   // get_shot_weights_local_grad
        const double g1 = s1 * tuple_times[t] + p1[(size_t)t] * local;
        const double g2 = s2 * tuple_times[t] + p2[(size_t)t] * local;
        mg[(size_t)t] = (g1 * c1 - g2 * c2) / std::sqrt(Nn);                        // get_merit_grad
    }
    if (merit) *merit = std::sqrt(tr) * (1.0 - sq / (4.0 * tr * tr)) / std::sqrt(Nn);
    if (sigma_tr) *sigma_tr = tr;
    if (sigma_sq_tr) *sigma_sq_tr = sq;
    if (merit_grad_log) {
        // (w w' - Diagonal(w)) * merit_grad (get_shot_weights_log_matrix)
        double dot = 0.0;
        for (int t = 0; t < T; ++t) dot += shot_weights[t] * mg[(size_t)t];
        for (int t = 0; t < T; ++t) merit_grad_log[t] = shot_weights[t] * dot - shot_weights[t] * mg[(size_t)t];
    }
    return ACES_OK;
}

}  // extern "C"
