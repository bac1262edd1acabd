// project_full.cuh -- full_project_gate_eigenvalues (src/estimate.jl:532-573): the simultaneous Mahalanobis
// projection of ALL gates onto their probability simplices, the reference's default below 5000 gate
// eigenvalues (split = false).  SURVEY.md section 8f, row f1.
//
// Included at the end of design.cu (it uses that file's internal helpers).
//
// The reference forms the N x N precision matrix P = D^-1 A' Omega'^-1 A D^-1 (calc_precision_matrix,
// src/merit.jl:754-770), transforms it to the space of unpadded gate probabilities,
//     Q = T P T',   T = pad' * W * pad_probs   (block diagonal: one n_g x n_g block per gate),
// and hands   min (y - v)' Q (y - v),  y >= 0   with all N variables to SCS (scs_project_nonnegative,
// src/utils.jl:57-97; an ADMM solver run to eps = 1e-8).  Here:
//   * P comes from the Gram matrix of the fit that has just run on this design (same clipping, weights
//     and block factors), rebuilt on the device like aces_design_precision_blocks does;
//   * Q = T P T' is two passes of small dot products (at most 15 terms each) over the N x N matrix;
//   * the quadratic program is solved EXACTLY by block principal pivoting (Judice & Pires; Kim & Park):
//     with F the variables off their bound, solve Q_FF y_F = (Q v)_F by Cholesky (dense::potrf on the
//     gathered block, two persistent triangular solves), form the multipliers w = Q y - Q v on the
//     bound set, exchange the variables that violate y >= 0 / w >= 0 (all of them while that reduces
//     their number, Murty's single exchange as the backup rule) and repeat.  A handful of iterations of
//     one N^3 / 3 factorisation each; the index logic is the host's.
// Parity with the reference is to SCS's tolerance; against another exact solver (Lawson-Hanson NNLS in the
// tests) to rounding.
#pragma once

namespace {

// X[i, j] = sum_c T_g[pos_i, c] * P[idx_g[c], j]   (rows) for transpose == 0
// Q[i, j] = sum_c X[i, idx_h[c]] * T_h[pos_j, c]   (columns) for transpose == 1
__global__ void gate_transform_kernel(int64_t N, int64_t ld, const int32_t* __restrict__ var_gate,
                                      const int32_t* __restrict__ var_pos, const int64_t* __restrict__ gate_ptr,
                                      const int32_t* __restrict__ gate_idx, const int64_t* __restrict__ blk_off,
                                      const double* __restrict__ T, const double* __restrict__ in,
                                      double* __restrict__ out, int columns) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // fast index: rows (coalesced)
    const int64_t j = blockIdx.y;
    if (i >= N) return;
    const int64_t tv = columns ? j : i;                 // the transformed index
    const int g = var_gate[tv], pos = var_pos[tv];
    const int64_t a = gate_ptr[g], n = gate_ptr[g + 1] - a;
    const double* Trow = T + blk_off[g] + (int64_t)pos * n;
    double s = 0.0;
    for (int64_t c = 0; c < n; ++c) {
        const int64_t src = gate_idx[a + c];
        s += Trow[c] * (columns ? in[i + src * ld] : in[src + j * ld]);
    }
    out[i + j * ld] = s;
}

// out[i] = sum_j Q[i, j] x[j] - h[i]  (Q symmetric, N x N, one warp per row chunk; deterministic)
__global__ void sym_residual_kernel(int64_t N, const double* __restrict__ Q, int64_t ld, const double* __restrict__ x,
                                    const double* __restrict__ h, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    double s0 = 0.0, s1 = 0.0;
    int64_t j = 0;
    for (; j + 1 < N; j += 2) {  // column-major: consecutive threads read consecutive rows of a column
        s0 += Q[i + j * ld] * x[j];
        s1 += Q[i + (j + 1) * ld] * x[j + 1];
    }
    if (j < N) s0 += Q[i + j * ld] * x[j];
    out[i] = (s0 + s1) - (h ? h[i] : 0.0);
}

// dst (nfp x nfp) = Q[F, F] with an identity on the padding; rhs_F = h[F]
__global__ void gather_free_kernel(int64_t nf, int64_t nfp, const int32_t* __restrict__ F, const double* __restrict__ Q,
                                   int64_t ld, const double* __restrict__ h, double* __restrict__ dst,
                                   double* __restrict__ rhs) {
    const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (a >= nfp) return;
    double v = (a == b) ? 1.0 : 0.0;
    if (a < nf && b < nf) v = Q[F[a] + (int64_t)F[b] * ld];
    dst[a + b * nfp] = v;
    if (b == 0) rhs[a] = a < nf ? h[F[a]] : 0.0;
}

__global__ void scatter_free_kernel(int64_t N, int64_t nf, const int32_t* __restrict__ F, const double* __restrict__ yF,
                                    double* __restrict__ y) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nf) y[F[i]] = yF[i];
}

__global__ void reciprocal_kernel(int64_t n, const double* __restrict__ g, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0 / g[i];
}

}  // namespace

extern "C" int32_t aces_design_project_full(aces_ctx* ctx, aces_design* d, int32_t ls_type, int32_t G,
                                            const int64_t* group_ptr, const int32_t* group_idx, int32_t index_base,
                                            const double* gate_eigenvalues, const double* transform_blocks,
                                            const double* v, double threshold, double* y_out, int32_t* iterations) {
    ACES_TRY(check_design(ctx, d, "design_project_full"));
    ACES_CHECK(ctx, !d->sharded, "design_project_full: the design is a tuple shard");
    ACES_CHECK(ctx, G >= 1 && group_ptr && group_idx && gate_eigenvalues && transform_blocks && v && y_out,
               "design_project_full: bad argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "design_project_full: index_base must be 0 or 1");
    ACES_CHECK(ctx, d->est_ls_type == ls_type && d->fs_stage == 0,
               "design_project_full: call aces_design_fit with the same ls_type on this design first (the precision "
               "matrix is that of the Gram matrix of that fit: its clipping, weights and block factors)");
    const int64_t N = d->N, Np = d->Np;
    ACES_CHECK(ctx, group_ptr[0] == 0 && group_ptr[G] == N, "design_project_full: the gates must cover the N gate eigenvalues once");
    // ---- host tables: gate and position of every variable, offsets of the transform blocks ---------------
    std::vector<int32_t> var_gate((size_t)N, -1), var_pos((size_t)N), idx((size_t)N);
    std::vector<int64_t> blk_off((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) {
        const int64_t n = group_ptr[g + 1] - group_ptr[g];
        ACES_CHECK(ctx, n >= 1 && n <= 255, "design_project_full: bad gate size");
        blk_off[(size_t)g + 1] = blk_off[(size_t)g] + n * n;
        for (int64_t e = group_ptr[g]; e < group_ptr[g + 1]; ++e) {
            const int64_t i = (int64_t)group_idx[e] - index_base;
            ACES_CHECK(ctx, i >= 0 && i < N && var_gate[(size_t)i] < 0, "design_project_full: index %lld outside 0..N-1 or listed twice",
                       (long long)i);
            var_gate[(size_t)i] = g;
            var_pos[(size_t)i] = (int32_t)(e - group_ptr[g]);
            idx[(size_t)e] = (int32_t)i;
        }
    }
    FitVectors fv = fit_vectors(d);
    ACES_TRY(ensure_workspaces(ctx, d, ls_type == 2, ls_type == 2, true, true));
    if (!d->W3) ACES_TRY(dev_alloc(ctx, d, &d->W3, (size_t)Np * Np));
    // the Gram matrix of the last fit again (its Cholesky factor overwrote it): same clipping and factors
    ACES_TRY(build_gram(ctx, d, ls_type, fv, d->fs_clipped ? d->row_map : nullptr));
    d->est_ls_type = ls_type;
    ACES_TRY(dense::symmetrize_from_lower(ctx, Np, d->G, Np));
    const size_t nT = (size_t)blk_off[(size_t)G];
    const size_t bytes = sizeof(int32_t) * 4 * (size_t)N + sizeof(int64_t) * 2 * ((size_t)G + 1) + sizeof(double) * nT + 256;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[6], bytes));
    uint8_t* base = (uint8_t*)ctx->d_work[6].ptr;
    int64_t* d_gptr = (int64_t*)base;
    int64_t* d_boff = d_gptr + (G + 1);
    double* d_T = (double*)(d_boff + (G + 1));
    int32_t* d_vg = (int32_t*)(d_T + nT);
    int32_t* d_vp = d_vg + N;
    int32_t* d_idx = d_vp + N;
    int32_t* d_F = d_idx + N;
    cudaStream_t st = ctx->stream;
    ACES_CUDA(ctx, cudaMemcpyAsync(d_gptr, group_ptr, sizeof(int64_t) * ((size_t)G + 1), cudaMemcpyHostToDevice, st));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_boff, blk_off.data(), sizeof(int64_t) * ((size_t)G + 1), cudaMemcpyHostToDevice, st));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_T, transform_blocks, sizeof(double) * nT, cudaMemcpyHostToDevice, st));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_vg, var_gate.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_vp, var_pos.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
    // P = D^-1 G D^-1 in place, X = T P (rows) into Xn, Q = X T' (columns) into Cn, upper triangle mirrored
    ACES_CUDA(ctx, cudaMemcpyAsync(fv.t1, gate_eigenvalues, sizeof(double) * (size_t)N, cudaMemcpyDefault, st));
    reciprocal_kernel<<<grid_for(N, 256), 256, 0, st>>>(N, fv.t1, fv.t2);
    scale_sym_kernel<<<148 * 8, 256, 0, st>>>(N, d->G, Np, fv.t2);
    dim3 tg((unsigned)((N + 127) / 128), (unsigned)N);
    gate_transform_kernel<<<tg, 128, 0, st>>>(N, Np, d_vg, d_vp, d_gptr, d_idx, d_boff, d_T, d->G, d->Xn, 0);
    gate_transform_kernel<<<tg, 128, 0, st>>>(N, Np, d_vg, d_vp, d_gptr, d_idx, d_boff, d_T, d->Xn, d->Cn, 1);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 4;
    ACES_TRY(eigen::mirror_upper(ctx, N, d->Cn, Np));  // P = UpperTriangular(precision_matrix) in the reference
    double* Q = d->Cn;
    double* h = fv.u;     // Q v
    double* yfull = fv.x; // current iterate (zeros on the bound set)
    double* w = fv.w;     // multipliers Q y - h
    ACES_CUDA(ctx, cudaMemcpyAsync(fv.t1, v, sizeof(double) * (size_t)N, cudaMemcpyDefault, st));
    sym_residual_kernel<<<grid_for(N, 128), 128, 0, st>>>(N, Q, Np, fv.t1, nullptr, h);
    ctx->launches += 1;
    std::vector<double> hv((size_t)N), hy((size_t)N), hw((size_t)N);
    ACES_CUDA(ctx, cudaMemcpyAsync(hv.data(), fv.t1, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
    ACES_CUDA(ctx, cudaStreamSynchronize(st));
    // ---- block principal pivoting ---------------------------------------------------------------------------
    std::vector<char> free_set((size_t)N);
    double vmax = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        free_set[(size_t)i] = hv[(size_t)i] > 0.0;
        vmax = std::max(vmax, std::fabs(hv[(size_t)i]));
    }
    const double tol = 1e-13 * std::max(vmax, 1e-300);
    int64_t best = N + 1;
    int grace = 3, it = 0;
    std::vector<int32_t> F;
    const int max_it = (int)std::min<int64_t>(10 * N + 50, 100000);
    for (;; ++it) {
        ACES_CHECK(ctx, it < max_it, "design_project_full: no convergence after %d pivoting steps", it);
        F.clear();
        for (int64_t i = 0; i < N; ++i)
            if (free_set[(size_t)i]) F.push_back((int32_t)i);
        const int64_t nf = (int64_t)F.size(), nfp = std::max<int64_t>(pad_up(nf), DB);
        ACES_CUDA(ctx, cudaMemsetAsync(yfull, 0, sizeof(double) * (size_t)Np, st));
        if (nf > 0) {
            ACES_CUDA(ctx, cudaMemcpyAsync(d_F, F.data(), sizeof(int32_t) * (size_t)nf, cudaMemcpyHostToDevice, st));
            dim3 gg((unsigned)((nfp + 127) / 128), (unsigned)nfp);
            gather_free_kernel<<<gg, 128, 0, st>>>(nf, nfp, d_F, Q, Np, h, d->W3, fv.v);
            ACES_CUDA(ctx, cudaMemsetAsync(d->flags + d->T, 0, sizeof(int32_t), st));
            ctx->launches += 1;
            ACES_TRY(dense::potrf(ctx, nfp, d->W3, nfp, d->dinvG, d->flags + d->T));
            ACES_TRY(dense::trsv_lower(ctx, nfp, d->W3, nfp, d->dinvG, fv.v, fv.t2));
            ACES_TRY(dense::trsv_lower_t(ctx, nfp, d->W3, nfp, d->dinvG, fv.t2, fv.g));
            scatter_free_kernel<<<grid_for(nf, 256), 256, 0, st>>>(N, nf, d_F, fv.g, yfull);
            ctx->launches += 1;
        }
        sym_residual_kernel<<<grid_for(N, 128), 128, 0, st>>>(N, Q, Np, yfull, h, w);
        ctx->launches += 1;
        int32_t flag = 0;
        ACES_CUDA(ctx, cudaMemcpyAsync(hy.data(), yfull, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
        ACES_CUDA(ctx, cudaMemcpyAsync(hw.data(), w, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
        ACES_CUDA(ctx, cudaMemcpyAsync(&flag, d->flags + d->T, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ACES_CUDA(ctx, cudaStreamSynchronize(st));
        if (flag) {
            d->est_ls_type = -1;
            return aces_fail(ctx, ACES_E_NOT_PD, "design_project_full: the transformed precision matrix is not positive definite");
        }
        // multipliers are compared on the scale of Q v
        double hmax = 0.0;
        for (int64_t i = 0; i < N; ++i) hmax = std::max(hmax, std::fabs(hw[(size_t)i]));
        const double tol_w = 1e-11 * std::max(hmax, 1e-300);
        int64_t bad = 0, last_bad = -1;
        for (int64_t i = 0; i < N; ++i) {
            const bool viol = free_set[(size_t)i] ? hy[(size_t)i] < -tol : hw[(size_t)i] < -tol_w;
            if (viol) {
                ++bad;
                last_bad = i;
            }
        }
        if (bad == 0) break;
        bool all = true;
        if (bad < best) {
            best = bad;
            grace = 3;
        } else if (grace > 0) {
            --grace;
        } else {
            all = false;  // Murty's rule: only the violating variable with the largest index changes sides
        }
        for (int64_t i = 0; i < N; ++i) {
            const bool viol = free_set[(size_t)i] ? hy[(size_t)i] < -tol : hw[(size_t)i] < -tol_w;
            if (viol && (all || i == last_bad)) free_set[(size_t)i] = !free_set[(size_t)i];
        }
    }
    d->est_ls_type = -1;  // the Gram workspace no longer holds the fit's matrix
    for (int64_t i = 0; i < N; ++i) {
        double yi = free_set[(size_t)i] ? hy[(size_t)i] : 0.0;
        if (yi < threshold) yi = 0.0;  // projected_vector[projected_vector .< precision] .= 0.0
        y_out[i] = yi;
    }
    if (iterations) *iterations = it + 1;
    return ACES_OK;
}
