// project.cu -- projection of the fitted gate eigenvalues onto the probability simplex, batched over
// the gates.
//
// Reference functions replaced (paths relative to the QuantumACES.jl root):
//   project_simplex                     src/utils.jl:36-50      Euclidean projection, sort based
//   simple_project_gate_eigenvalues     src/estimate.jl:635-663 per gate: WHT^-1, project_simplex, WHT
//   scs_project_nonnegative             src/utils.jl:57-97      min (y - v)' Q (y - v), y >= 0 (SCS there)
//   split_project_gate_eigenvalues      src/estimate.jl:580-628 per gate: Mahalanobis projection with the
//                                                               gate's slice of the precision matrix
//
// A fit has thousands of gates of 1, 3 or 15 eigenvalues each (SPAM / one-qubit / two-qubit gates:
// padded distributions of 2, 4, 16 probabilities); the reference loops over them on the host and
// calls SCS once per gate.  Here one thread owns one gate: the transforms are the symplectically
// ordered Walsh-Hadamard matrices of src/noise.jl:285-312 generated from bit tricks, the Euclidean
// projection is the reference's sort-based formula evaluated in the same order of operations (the
// results are bit-identical), and the Mahalanobis projection is an exact active-set solve of the
// non-negative quadratic program (the reference's SCS solve is iterative to eps = 1e-8, so parity is
// to that tolerance).
#include <algorithm>
#include <vector>

#include "aces_common.h"

namespace {

constexpr int PMAX = 16;  // padded probabilities of the largest supported gate (two qubits)

// Sign of the symplectically ordered Walsh-Hadamard matrix (src/noise.jl:285-299): index a holds
// the X bits in its low n bits and the Z bits in the next n; entry = (-1)^(a_x . b_z + a_z . b_x).
// SPAM / mid-circuit-reset gates (one eigenvalue, two probabilities) use [1 1; 1 -1] (src/noise.jl:302-306).
__device__ __forceinline__ double wht_sign(int a, int b, int n) {
    if (n == 0) return (a & b & 1) ? -1.0 : 1.0;
    const int mask = (1 << n) - 1;
    const int par = (__popc((a & mask) & (b >> n)) + __popc((a >> n) & (b & mask))) & 1;
    return par ? -1.0 : 1.0;
}

// padded size (2, 4, 16) -> number of qubits n (0 = the two-outcome SPAM transform), -1 unsupported
__host__ __device__ __forceinline__ int qubits_of(int padded) {
    return padded == 2 ? 0 : padded == 4 ? 1 : padded == 16 ? 2 : -1;
}

// q[i] = scale * sum_j W[i][j] x[j], j ascending (the order of the reference's sparse mat-vec; the
// entries are +-scale with scale a power of two, so every product is exact and only the order of the
// additions matters)
__device__ __forceinline__ void wht_apply(const double* x, double* q, int P, int n, double scale) {
    for (int i = 0; i < P; ++i) {
        double s = 0.0;
        for (int j = 0; j < P; ++j) s = __dadd_rn(s, __dmul_rn(__dmul_rn(wht_sign(i, j, n), scale), x[j]));
        q[i] = s;
    }
}

// project_simplex (src/utils.jl:36-50), same operations in the same order
__device__ __forceinline__ void project_simplex_dev(const double* p, double* out, int P) {
    double s[PMAX];
    for (int i = 0; i < P; ++i) s[i] = p[i];
    for (int i = 1; i < P; ++i) {  // insertion sort, descending
        const double v = s[i];
        int j = i - 1;
        while (j >= 0 && s[j] < v) {
            s[j + 1] = s[j];
            --j;
        }
        s[j + 1] = v;
    }
    double csum = 0.0, shift = 0.0;
    for (int i = 0; i < P; ++i) {
        csum = i == 0 ? s[0] : __dadd_rn(csum, s[i]);            // cumsum(sorted)
        const double sp = __dmul_rn(__ddiv_rn(1.0, (double)(i + 1)), __dsub_rn(1.0, csum));
        if (__dadd_rn(s[i], sp) > 0.0) shift = sp;              // findlast(sorted + sum .> 0)
    }
    for (int i = 0; i < P; ++i) out[i] = fmax(__dadd_rn(p[i], shift), 0.0);
}

// One thread per gate: eigenvalues -> padded probabilities -> projected -> eigenvalues.
__global__ void simple_project_kernel(int G, const int64_t* __restrict__ gptr, const int32_t* __restrict__ gidx,
                                      const double* __restrict__ unproj, double* __restrict__ proj,
                                      double* __restrict__ unproj_probs, double* __restrict__ probs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t e0 = gptr[g];
    const int n_g = (int)(gptr[g + 1] - e0), P = n_g + 1, n = qubits_of(P);
    const int64_t pad0 = e0 + g;  // padded vector: one extra (identity) entry per preceding gate
    double lam[PMAX], q[PMAX], pr[PMAX];
    lam[0] = 1.0;                // pad_mask (src/noise.jl:793-800)
    for (int i = 0; i < n_g; ++i) lam[i + 1] = unproj[gidx[e0 + i]];
    wht_apply(lam, q, P, n, 1.0 / (double)P);  // wht_transform_inv = W ./ 4^n (or ./ 2)
    project_simplex_dev(q, pr, P);
    for (int i = 0; i < P; ++i) {
        unproj_probs[pad0 + i] = q[i];
        probs[pad0 + i] = pr[i];
    }
    wht_apply(pr, lam, P, n, 1.0);
    for (int i = 0; i < n_g; ++i) proj[gidx[e0 + i]] = lam[i + 1];  // pad_transform' drops the identity
}

// ---- non-negative quadratic program, one thread per problem ---------------------------------------
// min_y 1/2 (y - v)' Q (y - v)  subject to  y >= 0, Q symmetric positive definite (n <= 15).
// Active-set method of Lawson & Hanson in its quadratic-program form: the passive set grows by the
// coordinate with the largest negative gradient, the equality-constrained minimiser on the passive
// set comes from a Cholesky solve, and a step that would leave the orthant is cut at the boundary.
constexpr int QMAX = 15;

// solves Q[idx, idx] z = r[idx] for the m passive indices; false when a pivot is not positive
__device__ bool solve_passive(const double* Q, int n, const int* idx, int m, const double* r, double* z) {
    double L[QMAX * QMAX];
    for (int a = 0; a < m; ++a)
        for (int b = 0; b <= a; ++b) {
            double s = Q[idx[a] * n + idx[b]];
            for (int k = 0; k < b; ++k) s -= L[a * QMAX + k] * L[b * QMAX + k];
            if (a == b) {
                if (!(s > 0.0)) return false;
                L[a * QMAX + a] = sqrt(s);
            } else {
                L[a * QMAX + b] = s / L[b * QMAX + b];
            }
        }
    double t[QMAX];
    for (int a = 0; a < m; ++a) {
        double s = r[idx[a]];
        for (int k = 0; k < a; ++k) s -= L[a * QMAX + k] * t[k];
        t[a] = s / L[a * QMAX + a];
    }
    for (int a = m - 1; a >= 0; --a) {
        double s = t[a];
        for (int k = a + 1; k < m; ++k) s -= L[k * QMAX + a] * z[k];
        z[a] = s / L[a * QMAX + a];
    }
    return true;
}

__global__ void nnqp_kernel(int n_problems, const int64_t* __restrict__ ptr, const int64_t* __restrict__ qoff,
                            const double* __restrict__ Qall, const double* __restrict__ vall, double threshold,
                            double* __restrict__ yall, int32_t* __restrict__ status) {
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= n_problems) return;
    const int n = (int)(ptr[pidx + 1] - ptr[pidx]);
    const double* Q = Qall + qoff[pidx];  // n x n, symmetric (either triangle order reads the same)
    const double* v = vall + ptr[pidx];
    double* yout = yall + ptr[pidx];
    double y[QMAX], r[QMAX], z[QMAX];
    int passive[QMAX], in_set[QMAX];
    int m = 0;
    double scale = 0.0;
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += Q[i * n + j] * v[j];
        r[i] = s;            // r = Q v: the objective is 1/2 y'Qy - y'r + const
        y[i] = 0.0;
        in_set[i] = 0;
        scale = fmax(scale, fabs(s));
    }
    const double tol = 1e-13 * scale;
    // the whole block must be positive definite (the reference hands it to SCS as a convex objective)
    for (int i = 0; i < n; ++i) passive[i] = i;
    int ok = solve_passive(Q, n, passive, n, r, z) ? 1 : 0;
    for (int outer = 0; outer < 4 * QMAX && ok; ++outer) {
        // w = r - Q y: the negative gradient
        int best = -1;
        double wbest = tol;
        for (int i = 0; i < n; ++i) {
            double s = r[i];
            for (int j = 0; j < n; ++j) s -= Q[i * n + j] * y[j];
            if (!in_set[i] && s > wbest) {
                wbest = s;
                best = i;
            }
        }
        if (best < 0) break;  // Karush-Kuhn-Tucker conditions hold
        passive[m++] = best;
        in_set[best] = 1;
        for (int inner = 0; inner < 2 * QMAX; ++inner) {
            if (!solve_passive(Q, n, passive, m, r, z)) {
                ok = 0;
                break;
            }
            double alpha = 2.0;
            for (int a = 0; a < m; ++a)
                if (z[a] <= 0.0) {
                    const double yi = y[passive[a]];
                    alpha = fmin(alpha, yi / (yi - z[a]));
                }
            if (alpha >= 2.0) {  // the minimiser on the passive set is interior
                for (int a = 0; a < m; ++a) y[passive[a]] = z[a];
                break;
            }
            int m2 = 0;
            for (int a = 0; a < m; ++a) {
                const int i = passive[a];
                const double yn = y[i] + alpha * (z[a] - y[i]);
                if (z[a] <= 0.0 && yn <= 1e-300 + 4e-16 * fabs(y[i])) {  // hits the boundary: back to the active set
                    y[i] = 0.0;
                    in_set[i] = 0;
                } else {
                    y[i] = yn;
                    passive[m2++] = i;
                }
            }
            m = m2;
            if (m == 0) break;
        }
    }
    // "Remove numerical noise" (src/utils.jl:94-95): entries below the solver precision are set to 0
    for (int i = 0; i < n; ++i) yout[i] = (y[i] < threshold) ? 0.0 : y[i];
    if (status) status[pidx] = ok ? 0 : 1;
}

}  // namespace

extern "C" {

int32_t aces_project_simplex(aces_ctx* ctx, int64_t N, int32_t G, const int64_t* group_ptr, const int32_t* group_idx,
                             int32_t index_base, const double* unproj_gate_eigenvalues, double* gate_eigenvalues,
                             double* unproj_probabilities, double* probabilities) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, N >= 1 && G >= 1 && group_ptr && group_idx && unproj_gate_eigenvalues && gate_eigenvalues,
               "project_simplex: bad argument");
    ACES_CHECK(ctx, index_base == 0 || index_base == 1, "project_simplex: index_base must be 0 or 1");
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nidx = group_ptr[G] - group_ptr[0];
    ACES_CHECK(ctx, group_ptr[0] == 0 && nidx >= 0, "project_simplex: group_ptr must start at 0");
    std::vector<int32_t> idx((size_t)nidx);
    std::vector<uint8_t> seen((size_t)N, 0);
    for (int g = 0; g < G; ++g) {
        const int64_t sz = group_ptr[g + 1] - group_ptr[g];
        ACES_CHECK(ctx, sz >= 0 && qubits_of((int)sz + 1) >= 0,
                   "project_simplex: gate %d has %lld eigenvalues; supported are 1 (SPAM), 3 and 15", g, (long long)sz);
        for (int64_t e = group_ptr[g]; e < group_ptr[g + 1]; ++e) {
            const int64_t i = (int64_t)group_idx[e] - index_base;
            ACES_CHECK(ctx, i >= 0 && i < N, "project_simplex: index %lld outside 0..N-1", (long long)i);
            idx[(size_t)e] = (int32_t)i;
            seen[(size_t)i] = 1;
        }
    }
    const int64_t n_pad = nidx + G;
    // device layout: [gptr | idx | unproj | proj | unproj_probs | probs]
    const size_t b_ptr = sizeof(int64_t) * ((size_t)G + 1), b_idx = (sizeof(int32_t) * (size_t)nidx + 7) & ~(size_t)7;
    const size_t b_n = sizeof(double) * (size_t)N, b_pad = sizeof(double) * (size_t)n_pad;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[7], b_ptr + b_idx + 2 * b_n + 2 * b_pad + 64));
    uint8_t* base = (uint8_t*)ctx->d_work[7].ptr;
    int64_t* d_ptr = (int64_t*)base;
    int32_t* d_idx = (int32_t*)(base + b_ptr);
    double* d_un = (double*)(base + b_ptr + b_idx);
    double* d_pr = d_un + N;
    double* d_up = d_pr + N;
    double* d_pp = d_up + n_pad;
    ACES_CUDA(ctx, cudaMemcpyAsync(d_ptr, group_ptr, b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_idx, idx.data(), sizeof(int32_t) * (size_t)nidx, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_un, unproj_gate_eigenvalues, b_n, cudaMemcpyDefault, ctx->stream));
    // eigenvalues that belong to no gate pass through unchanged
    ACES_CUDA(ctx, cudaMemcpyAsync(d_pr, d_un, b_n, cudaMemcpyDeviceToDevice, ctx->stream));
    simple_project_kernel<<<(unsigned)((G + 127) / 128), 128, 0, ctx->stream>>>(G, d_ptr, d_idx, d_un, d_pr, d_up, d_pp);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(gate_eigenvalues, d_pr, b_n, cudaMemcpyDefault, ctx->stream));
    if (unproj_probabilities) ACES_CUDA(ctx, cudaMemcpyAsync(unproj_probabilities, d_up, b_pad, cudaMemcpyDefault, ctx->stream));
    if (probabilities) ACES_CUDA(ctx, cudaMemcpyAsync(probabilities, d_pp, b_pad, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `idx` dies at return
    return ACES_OK;
}

int32_t aces_project_nonnegative_batched(aces_ctx* ctx, int32_t n_problems, const int64_t* ptr, const double* Q,
                                         const double* v, double threshold, double* y, int32_t* n_failed) {
    if (!ctx) return ACES_E_INVALID;
    ACES_CHECK(ctx, n_problems >= 0 && ptr && Q && v && y, "project_nonnegative_batched: bad argument");
    if (n_failed) *n_failed = 0;
    if (n_problems == 0) return ACES_OK;
    ACES_CUDA(ctx, cudaSetDevice(ctx->device));
    ACES_CHECK(ctx, ptr[0] == 0, "project_nonnegative_batched: ptr must start at 0");
    std::vector<int64_t> qoff((size_t)n_problems + 1, 0);
    for (int p = 0; p < n_problems; ++p) {
        const int64_t n = ptr[p + 1] - ptr[p];
        ACES_CHECK(ctx, n >= 0 && n <= QMAX, "project_nonnegative_batched: problem %d has %lld variables (at most %d)", p,
                   (long long)n, QMAX);
        qoff[(size_t)p + 1] = qoff[(size_t)p] + n * n;
    }
    const int64_t nv = ptr[n_problems], nq = qoff[(size_t)n_problems];
    const size_t b_ptr = sizeof(int64_t) * ((size_t)n_problems + 1);
    const size_t b_st = (sizeof(int32_t) * (size_t)n_problems + 7) & ~(size_t)7;
    ACES_TRY(aces_reserve_dev(ctx, ctx->d_work[7], 2 * b_ptr + b_st + sizeof(double) * (size_t)(nq + 2 * nv) + 64));
    uint8_t* base = (uint8_t*)ctx->d_work[7].ptr;
    int64_t* d_ptr = (int64_t*)base;
    int64_t* d_qoff = (int64_t*)(base + b_ptr);
    int32_t* d_st = (int32_t*)(base + 2 * b_ptr);
    double* d_Q = (double*)(base + 2 * b_ptr + b_st);
    double* d_v = d_Q + nq;
    double* d_y = d_v + nv;
    ACES_CUDA(ctx, cudaMemcpyAsync(d_ptr, ptr, b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_qoff, qoff.data(), b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_Q, Q, sizeof(double) * (size_t)nq, cudaMemcpyDefault, ctx->stream));
    ACES_CUDA(ctx, cudaMemcpyAsync(d_v, v, sizeof(double) * (size_t)nv, cudaMemcpyDefault, ctx->stream));
    nnqp_kernel<<<(unsigned)((n_problems + 63) / 64), 64, 0, ctx->stream>>>(n_problems, d_ptr, d_qoff, d_Q, d_v, threshold,
                                                                          d_y, d_st);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ACES_CUDA(ctx, cudaMemcpyAsync(y, d_y, sizeof(double) * (size_t)nv, cudaMemcpyDefault, ctx->stream));
    std::vector<int32_t> st((size_t)n_problems);
    ACES_CUDA(ctx, cudaMemcpyAsync(st.data(), d_st, sizeof(int32_t) * (size_t)n_problems, cudaMemcpyDeviceToHost, ctx->stream));
    ACES_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int bad = 0;
    for (int32_t s : st) bad += s != 0;
    if (n_failed) *n_failed = bad;
    if (bad)
        return aces_fail(ctx, ACES_E_NOT_PD, "project_nonnegative_batched: %d of %d precision-matrix blocks are not positive "
                         "definite", bad, n_problems);
    return ACES_OK;
}

}  // extern "C"
