// dense.cuh -- internal FP64 dense engine of the fit path (K3-K6): one DMMA GEMM kernel and the
// blocked factorisations built on it.  Everything is column-major, device-resident and padded:
// every matrix handed to these routines has dimensions that are multiples of DB = 128 and an
// even leading dimension, with the padding kept harmless by the caller (zeros, or an identity
// block on the diagonal of a matrix that is factorised).  That removes all edge handling from
// the kernels.
#pragma once
#include "aces_common.h"

namespace dense {

constexpr int DB = 128;  // block size of the blocked algorithms == GEMM tile edge

inline int64_t pad_up(int64_t x) { return (x + DB - 1) / DB * DB; }

// How op(A)(i,k) and op(B)(k,j) are laid out in memory.
enum GemmForm {
    GEMM_NT = 0,  // A: m x k col-major (lda), B: n x k col-major (ldb)      C = A * B'
    GEMM_NN = 1,  // A: m x k col-major,       B: k x n col-major            C = A * B
    GEMM_TN = 2,  // A: k x m col-major,       B: k x n col-major            C = A' * B
};

// One GEMM of a batched launch (device-resident table, one entry per blockIdx.z).
struct GemmDesc {
    const double* A;
    const double* B;
    double* C;
    int64_t lda, ldb, ldc;
    int32_t k;           // inner dimension (multiple of 16); 0 = inactive entry
    int32_t tiles_m;     // m / 128
    int32_t tiles_n;     // n / 128
    int32_t lower_only;  // compute only tiles with bn <= bm
    double alpha, beta;
};

// One 128 x 128 diagonal block of a (batched) Cholesky step.
struct DiagDesc {
    double* A;           // the block inside the matrix being factorised (NULL = inactive)
    int64_t lda;
    double* dinv;        // receives inv(L_kk), 128 x 128 col-major
    const double* diag0; // original diagonal of the block (pivot test), may be NULL
    int32_t* flag;       // set to code when a pivot fails (must be 0 before), may be NULL
    int32_t code;
    int32_t dead_mode;   // != 0: a failed pivot eliminates its column and *flag counts them (see potrf)
};

// C (m x n, ldc) = alpha * op(A) op(B) + beta * C.  m, n multiples of 128, k a multiple of 16.
// lower_only != 0: only tiles on or below the block diagonal are computed (symmetric updates).
int gemm(aces_ctx* ctx, GemmForm form, int64_t m, int64_t n, int64_t k, double alpha,
         const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
         int64_t ldc, int lower_only);

// `count` GEMMs of the same form in one launch; descs is a DEVICE array.  max_tiles_m/n bound
// the grid (entries with fewer tiles leave the surplus CTAs idle).
int gemm_batched(aces_ctx* ctx, GemmForm form, const GemmDesc* descs, int count, int max_tiles_m,
                 int max_tiles_n);

// Factorises `count` 128 x 128 diagonal blocks (descs: DEVICE array) in one launch: L_kk in place
// (strict upper triangle zeroed) and inv(L_kk) into dinv.  A pivot that is not positive or, with
// diag0, not above rel_tol * |diag0|, sets *flag = code (first failure wins) and is replaced by 1.
int diag_blocks(aces_ctx* ctx, const DiagDesc* descs, int count, double rel_tol);

// In-place lower Cholesky of the n x n matrix A (n a multiple of 128).  dinv receives the
// inverses of the 128 x 128 diagonal blocks of L ([n/128][128*128], col-major).  flag (device
// int, must be zero on entry) is set to 1 + the failing block when a pivot is not positive or,
// with diag0 (the diagonal of A before the call, device), not above rel_tol * |diag0|.
// descs: device scratch for n/128 DiagDesc entries (filled here through `stage`, a pinned or
// pageable host array of the same length that must stay alive until the stream has consumed it).
// dead_mode != 0 (the Gram matrix of a least-squares fit): a pivot that fails the test is not an
// error; its column is eliminated (pivot replaced by 1e300, so the column of L underflows to ~0 and
// the variable gets the value 0 in every solve with this factor -- the basic solution of the
// rank-deficient problem), and *flag receives the NUMBER of eliminated columns.
int potrf(aces_ctx* ctx, int64_t n, double* A, int64_t lda, double* dinv, int* flag,
          const double* diag0 = nullptr, double rel_tol = 0.0, int dead_mode = 0);
int get_diag(aces_ctx* ctx, int64_t n, const double* A, int64_t lda, double* d);

// X = inv(L) for the lower-triangular factor produced by potrf (needs its dinv).  X is n x n
// (ldx); only its lower triangle is written, the strict upper triangle is zeroed.
int trtri(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* X,
          int64_t ldx);

// Triangular solves with one right-hand side, one persistent cooperative launch each (the
// 128-blocks hand their solution to the dependent blocks through flags in global memory):
//   trsv_lower:   y = inv(L)  * b
//   trsv_lower_t: x = inv(L') * y
int trsv_lower(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* b,
               double* y);
int trsv_lower_t(aces_ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* dinv, double* y,
                 double* x);

// B (n x nrhs, ldb) <- inv(L) * B by blocked forward substitution (nrhs a multiple of 128).
int trsm_lower(aces_ctx* ctx, int64_t n, int64_t nrhs, const double* L, int64_t ldl,
               const double* dinv, double* B, int64_t ldb);

// small helpers (all asynchronous on ctx->stream)
int set_identity(aces_ctx* ctx, int64_t n, double* A, int64_t lda);           // A = I
int fill_diag_range(aces_ctx* ctx, double* A, int64_t lda, int64_t from, int64_t to, double v);
int transpose(aces_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb);
int symmetrize_from_lower(aces_ctx* ctx, int64_t n, double* A, int64_t lda);  // A = tril(A) + tril(A,-1)'

}  // namespace dense
