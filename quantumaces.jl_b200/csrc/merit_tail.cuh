// merit_tail.cuh -- the numeric core of calc_merit(d::Design) (src/merit.jl:925-1006), SURVEY 8f row f2:
// the three estimator covariances, their marginal / relative transforms T Sigma T' (src/merit.jl:477-509)
// and the nine spectra (eigen.cu), plus the extreme eigenvalues of d.matrix' * d.matrix for the condition
// number and the pseudoinverse norm.  Everything stays on the device between the stages; only the
// spectra come back.
//
// Included at the end of design.cu (it uses that file's internal helpers).
#pragma once

#include "eigen.cuh"

namespace {

struct HostCsr {
    std::vector<int64_t> rowptr;
    std::vector<int32_t> colidx;
    std::vector<double> vals;
};

// Julia SparseMatrixCSC (rows x N) -> CSR, entries of a row in column order
int csc_to_csr(aces_ctx* ctx, const char* what, int64_t rows, int64_t N, const int64_t* colptr, const int64_t* rowval,
               const double* nzval, int base, HostCsr* out) {
    ACES_CHECK(ctx, colptr && rowval && nzval, "design_merit: NULL %s transform", what);
    const int64_t nnz = colptr[N] - base;
    ACES_CHECK(ctx, nnz >= 0, "design_merit: bad %s colptr", what);
    out->rowptr.assign((size_t)rows + 1, 0);
    out->colidx.resize((size_t)nnz);
    out->vals.resize((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t r = rowval[e] - base;
        ACES_CHECK(ctx, r >= 0 && r < rows, "design_merit: %s transform row %lld outside 0..%lld", what, (long long)r,
                   (long long)rows - 1);
        out->rowptr[(size_t)r + 1]++;
    }
    for (int64_t r = 0; r < rows; ++r) out->rowptr[(size_t)r + 1] += out->rowptr[(size_t)r];
    std::vector<int64_t> fill(out->rowptr.begin(), out->rowptr.end() - 1);
    for (int64_t j = 0; j < N; ++j)
        for (int64_t e = colptr[j] - base; e < colptr[j + 1] - base; ++e) {
            const int64_t p = fill[(size_t)(rowval[e] - base)]++;
            out->colidx[(size_t)p] = (int32_t)j;
            out->vals[(size_t)p] = nzval[e];
        }
    return ACES_OK;
}

// out (np x np, zero padded) = T S T' for a sparse T (rows x N, CSR) and a symmetric S.  One thread per
// entry of the lower triangle, mirrored, so the result is exactly symmetric.
__global__ void transform_sandwich_kernel(int64_t rows, int64_t np, const int64_t* __restrict__ rowptr,
                                          const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                                          const double* __restrict__ S, int64_t lds, double* __restrict__ out) {
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // row (fast index: coalesced stores)
    const int64_t b = blockIdx.y;
    if (a >= np) return;
    if (a < b) return;
    double s = 0.0;
    if (a < rows && b < rows) {
        for (int64_t eb = rowptr[b]; eb < rowptr[b + 1]; ++eb) {
            const double* col = S + (int64_t)colidx[eb] * lds;
            double t = 0.0;
            for (int64_t ea = rowptr[a]; ea < rowptr[a + 1]; ++ea) t += vals[ea] * col[colidx[ea]];
            s += t * vals[eb];
        }
    }
    out[a + b * np] = s;
    out[b + a * np] = s;
}

struct DevCsr {
    int64_t rows = 0, np = 0, nnz = 0;
    int64_t* rowptr = nullptr;
    int32_t* colidx = nullptr;
    double* vals = nullptr;
};

int upload_csr(aces_ctx* ctx, aces_design* d, const HostCsr& h, int64_t rows, DevCsr* out) {
    out->rows = rows;
    out->np = dense::pad_up(std::max<int64_t>(rows, 1));
    out->nnz = (int64_t)h.colidx.size();
    ACES_TRY(dev_upload(ctx, d, &out->rowptr, h.rowptr));
    ACES_TRY(dev_upload(ctx, d, &out->colidx, h.colidx));
    ACES_TRY(dev_upload(ctx, d, &out->vals, h.vals));
    return ACES_OK;
}

void release_csr(aces_design* d, DevCsr* c) {
    void* ptrs[] = {c->rowptr, c->colidx, c->vals};
    for (void* p : ptrs) {
        if (!p) continue;
        auto it = std::find(d->owned.begin(), d->owned.end(), p);
        if (it != d->owned.end()) d->owned.erase(it);
        cudaFree(p);
    }
    *c = DevCsr{};
}

int transform_and_eigen(aces_ctx* ctx, const DevCsr& T, const double* S, int64_t lds, double* buf, double* w_dev) {
    if (T.rows == 0) return ACES_OK;
    dim3 grid((unsigned)((T.np + 127) / 128), (unsigned)T.np);
    transform_sandwich_kernel<<<grid, 128, 0, ctx->stream>>>(T.rows, T.np, T.rowptr, T.colidx, T.vals, S, lds, buf);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return eigen::syev_values(ctx, T.rows, T.np, buf, T.np, w_dev);
}

}  // namespace

extern "C" int32_t aces_design_merit(aces_ctx* ctx, aces_design* d, const double* gate_eigenvalues,
                                     const double* covariance_log_nzval, int64_t n_marginal,
                                     const int64_t* marg_colptr, const int64_t* marg_rowval, const double* marg_nzval,
                                     int64_t n_relative, const int64_t* rel_colptr, const int64_t* rel_rowval,
                                     const double* rel_nzval, int32_t index_base, double* cov_eigenvalues,
                                     double* gram_extremes) {
    ACES_TRY(check_design(ctx, d, "design_merit"));
    ACES_CHECK(ctx, !d->sharded, "design_merit: the design is a tuple shard");
    ACES_CHECK(ctx, gate_eigenvalues && covariance_log_nzval && cov_eigenvalues, "design_merit: NULL argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "design_merit: index_base must be 0 or 1");
    const int64_t N = d->N, Np = d->Np;
    ACES_CHECK(ctx, n_marginal >= 0 && n_marginal <= N && n_relative >= 0 && n_relative <= N,
               "design_merit: transforms must not have more rows than gate eigenvalues");
    ACES_CHECK(ctx, N <= eigen::max_order(ctx), "design_merit: %lld gate eigenvalues exceed the eigenvalue solver's %lld",
               (long long)N, (long long)eigen::max_order(ctx));
    HostCsr hm, hr;
    if (n_marginal) ACES_TRY(csc_to_csr(ctx, "marginal", n_marginal, N, marg_colptr, marg_rowval, marg_nzval, index_base, &hm));
    if (n_relative) ACES_TRY(csc_to_csr(ctx, "relative", n_relative, N, rel_colptr, rel_rowval, rel_nzval, index_base, &hr));
    DevCsr Tm, Tr;
    auto cleanup = [&](int rc) {
        release_csr(d, &Tm);
        release_csr(d, &Tr);
        return rc;
    };
    if (n_marginal) {
        const int rc = upload_csr(ctx, d, hm, n_marginal, &Tm);
        if (rc != ACES_OK) return cleanup(rc);
    }
    if (n_relative) {
        const int rc = upload_csr(ctx, d, hr, n_relative, &Tr);
        if (rc != ACES_OK) return cleanup(rc);
    }
    const int64_t per = N + n_marginal + n_relative;
    {
        int rc = aces_reserve_dev(ctx, ctx->d_eigA, sizeof(double) * (size_t)(3 * per + N));
        if (rc == ACES_OK) rc = ensure_workspaces(ctx, d, true, true, true, true);
        if (rc == ACES_OK && !d->W3) rc = dev_alloc(ctx, d, &d->W3, (size_t)Np * Np);
        if (rc != ACES_OK) return cleanup(rc);
    }
    double* w = (double*)ctx->d_eigA.ptr;
    // the covariance values are shared by the three estimators: upload once
    {
        cudaError_t e = cudaMemcpyAsync(d->covval, covariance_log_nzval, sizeof(double) * (size_t)d->nnz_cov,
                                        cudaMemcpyDefault, ctx->stream);
        if (e != cudaSuccess) return cleanup(aces_fail(ctx, ACES_E_CUDA, "design_merit: %s", cudaGetErrorString(e)));
    }
    const int order[3] = {2, 1, 0};  // GLS, WLS, OLS: the order of the Merit fields
    for (int k = 0; k < 3; ++k) {
        double* S = nullptr;
        int rc = ls_covariance_device(ctx, d, order[k], gate_eigenvalues, d->covval, &S);
        if (rc == ACES_OK) rc = check_flags(ctx, d, "design_merit");
        // marginal into Xn, relative into W3 (both free once Sigma exists), then Sigma itself in place
        double* wk = w + (size_t)k * per;
        if (rc == ACES_OK) rc = transform_and_eigen(ctx, Tm, S, Np, d->Xn, wk + N);
        if (rc == ACES_OK) rc = transform_and_eigen(ctx, Tr, S, Np, d->W3, wk + N + n_marginal);
        if (rc == ACES_OK) rc = eigen::syev_values(ctx, N, Np, S, Np, wk);
        if (rc != ACES_OK) return cleanup(rc);
    }
    if (gram_extremes) {
        // d.matrix' * d.matrix: the OLS Gram matrix (unit weights, no clipping)
        FitVectors fv = fit_vectors(d);
        diag_factor_kernel<<<grid_for(d->M, 256), 256, 0, ctx->stream>>>(d->M, d->cov_colptr, d->cov_rowval, d->covval,
                                                                        d->blk_of_row, d->d_row0, d->d_prow0, nullptr,
                                                                        d->d_tuple_on, 1, fv.f);
        ctx->launches += 1;
        int rc = build_gram(ctx, d, 0, fv, nullptr);
        if (rc == ACES_OK) rc = dense::symmetrize_from_lower(ctx, Np, d->G, Np);
        if (rc == ACES_OK) rc = eigen::syev_values(ctx, N, Np, d->G, Np, w + (size_t)3 * per);
        if (rc != ACES_OK) return cleanup(rc);
    }
    d->est_ls_type = -1;
    cudaError_t e = cudaMemcpyAsync(cov_eigenvalues, w, sizeof(double) * (size_t)(3 * per), cudaMemcpyDefault, ctx->stream);
    double ext[2] = {0.0, 0.0};
    if (e == cudaSuccess && gram_extremes) {
        e = cudaMemcpyAsync(&ext[0], w + (size_t)3 * per + N - 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&ext[1], w + (size_t)3 * per, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cleanup(aces_fail(ctx, ACES_E_CUDA, "design_merit: %s", cudaGetErrorString(e)));
    if (gram_extremes) {
        gram_extremes[0] = ext[0];
        gram_extremes[1] = ext[1];
    }
    return cleanup(ACES_OK);
}
