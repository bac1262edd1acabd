// dense.cu -- FP64 dense engine: DMMA (mma.sync m8n8k4 f64, SASS DMMA) GEMM and the blocked
// Cholesky / triangular inverse / triangular solve built on it.  See dense.cuh for the
// conventions (column-major, padded to multiples of 128, no edge handling).
//
// Replaces, for the fit path, what the reference gets from LAPACK / SuiteSparse:
// cholesky + inv of the Gram matrix (src/merit.jl:564,586), the per-block covariance
// factorisations (src/merit.jl:398-402) and the least-squares solve (src/estimate.jl:500-502).
#include "dense.cuh"

#include <algorithm>

namespace dense {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, GT = 256;
constexpr int LD_MC = BM + 4;  // row stride of a [BK][BM] tile (operand contiguous along m / n)
constexpr int LD_KC = BK + 4;  // row stride of a [BM][BK] tile (operand contiguous along k)
constexpr int OPER_DOUBLES = BM * LD_KC;  // 2560 >= BK * LD_MC = 2112
constexpr int STAGE_DOUBLES = 2 * OPER_DOUBLES;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);

struct GemmP {
    const double* A;
    const double* B;
    double* C;
    int64_t lda, ldb, ldc;
    int k;
    double alpha, beta;
    int lower_only;
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

// One CTA computes one 128x128 tile of C.  8 warps as 2 (m) x 4 (n), warp tile 64x32 =
// 8x4 DMMA tiles; operands staged by cp.async in a 3-stage ring.
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GT, 1) gemm_kernel(const GemmP p) {
    const int bm = blockIdx.x, bn = blockIdx.y;
    if (p.lower_only && bn > bm) return;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    // A_KC: op(A)(i,k) = A[k + i*lda]; else A[i + k*lda].  Same for B with (j,k).
    const double* Ag = A_KC ? p.A + (int64_t)bm * BM * p.lda : p.A + (int64_t)bm * BM;
    const double* Bg = B_KC ? p.B + (int64_t)bn * BN * p.ldb : p.B + (int64_t)bn * BN;
    const int nk = p.k / BK;

    auto load = [&](int stage, int kt) {
        double* sA = sm + stage * STAGE_DOUBLES;
        double* sB = sA + OPER_DOUBLES;
        const int k0 = kt * BK;
#pragma unroll
        for (int c = tid; c < 1024; c += GT) {
            if (!A_KC) {
                const int kk = c >> 6, ci = c & 63;
                cp_async16(sA + kk * LD_MC + 2 * ci, Ag + 2 * ci + (int64_t)(k0 + kk) * p.lda);
            } else {
                const int i = c >> 3, ck = c & 7;
                cp_async16(sA + i * LD_KC + 2 * ck, Ag + (k0 + 2 * ck) + (int64_t)i * p.lda);
            }
            if (!B_KC) {
                const int kk = c >> 6, ci = c & 63;
                cp_async16(sB + kk * LD_MC + 2 * ci, Bg + 2 * ci + (int64_t)(k0 + kk) * p.ldb);
            } else {
                const int j = c >> 3, ck = c & 7;
                cp_async16(sB + j * LD_KC + 2 * ck, Bg + (k0 + 2 * ck) + (int64_t)j * p.ldb);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load(nxt % STAGES, nxt);
        cp_async_commit();
        const double* sA = sm + (kt % STAGES) * STAGE_DOUBLES;
        const double* sB = sA + OPER_DOUBLES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = A_KC ? sA[(wm + i * 8 + g) * LD_KC + k4 * 4 + t]
                            : sA[(k4 * 4 + t) * LD_MC + wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = B_KC ? sB[(wn + j * 8 + g) * LD_KC + k4 * 4 + t]
                            : sB[(k4 * 4 + t) * LD_MC + wn + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[g][2t], C[g][2t+1] of every 8x8 tile
    double* Cg = p.C + (int64_t)bm * BM + (int64_t)bn * BN * p.ldc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wm + i * 8 + g;
            const int c = wn + j * 8 + 2 * t;
            double* c0 = Cg + r + (int64_t)c * p.ldc;
            double* c1 = c0 + p.ldc;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            if (p.beta != 0.0) {
                v0 += p.beta * *c0;
                v1 += p.beta * *c1;
            }
            *c0 = v0;
            *c1 = v1;
        }
}

// Unblocked Cholesky of one 128x128 diagonal block in shared memory; writes L (strict upper
// triangle zeroed) back in place.  1024 threads.
constexpr int PLD = DB + 1;
__global__ void __launch_bounds__(1024, 1) potf2_kernel(double* A, int64_t lda, int* flag, int blk,
                                                        const double* diag0, double rel_tol) {
    extern __shared__ double s[];  // s[j*PLD + i] = A(i, j)
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int e = tid; e < DB * DB; e += 1024) {
        const int i = e & (DB - 1), j = e >> 7;
        s[j * PLD + i] = A[i + (int64_t)j * lda];
    }
    for (int k = 0; k < DB; ++k) {
        __syncthreads();
        double d = s[k * PLD + k];
        // a pivot that is not positive, or negligible against the original diagonal entry,
        // means the matrix is numerically not positive definite
        const double floor_k = diag0 ? rel_tol * fabs(diag0[k]) : 0.0;
        if (!(d > floor_k)) {
            if (tid == 0) atomicCAS(flag, 0, blk + 1);
            d = 1.0;
        }
        d = sqrt(d);
        __syncthreads();
        if (tid < DB) {
            if (tid > k) s[k * PLD + tid] /= d;
            else if (tid == k) s[k * PLD + k] = d;
        }
        __syncthreads();
        for (int j = k + 1 + ty; j < DB; j += 32) {
            const double lkj = s[k * PLD + j];
            for (int i = k + 1 + tx; i < DB; i += 32)
                if (i >= j) s[j * PLD + i] -= s[k * PLD + i] * lkj;
        }
    }
    __syncthreads();
    for (int e = tid; e < DB * DB; e += 1024) {
        const int i = e & (DB - 1), j = e >> 7;
        A[i + (int64_t)j * lda] = (i >= j) ? s[j * PLD + i] : 0.0;
    }
}

// X = inv(L) for one 128x128 lower-triangular block.  512 threads: column j is solved by four
// threads that each keep 32 entries of it in registers (forward substitution, all threads in
// lock step; entries not yet computed are zero, so no masking is needed).
__global__ void __launch_bounds__(512, 1) trtri128_kernel(const double* L, int64_t ldl, double* X) {
    extern __shared__ double s[];  // s[k*PLD + i] = L(i, k)
    const int tid = threadIdx.x;
    for (int e = tid; e < DB * DB; e += 512) {
        const int i = e & (DB - 1), j = e >> 7;
        s[j * PLD + i] = (i >= j) ? L[i + (int64_t)j * ldl] : 0.0;
    }
    __syncthreads();
    const int j = tid >> 2, l = tid & 3;
    double x[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) x[q] = 0.0;
    for (int i = 0; i < DB; ++i) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) sum += s[(4 * q + l) * PLD + i] * x[q];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const double val = ((i == j ? 1.0 : 0.0) - sum) / s[i * PLD + i];
        const int qi = i >> 2;
        const bool mine = (i & 3) == l;
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if (mine && q == qi) x[q] = val;
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) X[(4 * q + l) + j * DB] = x[q];
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
__global__ void symmetrize_kernel(int64_t n, double* A, int64_t lda) {
    const int64_t total = n * n;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % n, j = e / n;
        if (i < j) A[i + j * lda] = A[j + i * lda];
    }
}

template <bool A_KC, bool B_KC>
int launch_gemm(aces_ctx* ctx, const GemmP& p, int64_t m, int64_t n) {
    auto kern = gemm_kernel<A_KC, B_KC>;
    static bool configured = false;
    if (!configured) {
        ACES_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        configured = true;
    }
    dim3 grid((unsigned)(m / BM), (unsigned)(n / BN));
    kern<<<grid, GT, GEMM_SMEM, ctx->stream>>>(p);
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
    GemmP p{A, B, C, lda, ldb, ldc, (int)k, alpha, beta, lower_only};
    switch (form) {
        case GEMM_NT: return launch_gemm<false, false>(ctx, p, m, n);
        case GEMM_NN: return launch_gemm<false, true>(ctx, p, m, n);
        case GEMM_TN: return launch_gemm<true, true>(ctx, p, m, n);
    }
    return aces_fail(ctx, ACES_E_INVALID, "dense::gemm: unknown form");
}

int potrf(aces_ctx* ctx, int64_t n, double* A, int64_t lda, double* dinv, int* flag, const double* diag0,
          double rel_tol) {
    ACES_CHECK(ctx, n % DB == 0, "dense::potrf: n=%lld is not padded", (long long)n);
    static bool configured = false;
    const size_t smem = (size_t)DB * PLD * sizeof(double);
    if (!configured) {
        ACES_CUDA(ctx, cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ACES_CUDA(ctx, cudaFuncSetAttribute(trtri128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t nb = n / DB;
    for (int64_t k = 0; k < nb; ++k) {
        double* Akk = A + k * DB * (lda + 1);
        double* dk = dinv + k * DB * DB;
        potf2_kernel<<<1, 1024, smem, ctx->stream>>>(Akk, lda, flag, (int)k, diag0 ? diag0 + k * DB : nullptr, rel_tol);
        trtri128_kernel<<<1, 512, smem, ctx->stream>>>(Akk, lda, dk);
        ACES_CUDA(ctx, cudaGetLastError());
        ctx->launches += 2;
        const int64_t rows = n - (k + 1) * DB;
        if (rows > 0) {
            double* panel = Akk + DB;
            // panel <- panel * inv(L_kk)'   (in place: one CTA per 128-row tile reads exactly what it writes)
            ACES_TRY(gemm(ctx, GEMM_NT, rows, DB, DB, 1.0, panel, lda, dk, DB, 0.0, panel, lda, 0));
            // trailing (lower tiles) -= panel * panel'
            ACES_TRY(gemm(ctx, GEMM_NT, rows, rows, DB, -1.0, panel, lda, panel, lda, 1.0,
                          Akk + DB * (lda + 1), lda, 1));
        }
    }
    return ACES_OK;
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
    symmetrize_kernel<<<(unsigned)std::min<int64_t>((n * n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(n, A, lda);
    ACES_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return ACES_OK;
}

}  // namespace dense
