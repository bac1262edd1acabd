// eigen.cuh -- eigenvalues of a dense symmetric FP64 matrix (eigen.cu); internal interface.
#pragma once
#include "aces_common.h"

namespace eigen {

// Largest order syev_values accepts on this context's device (the panel kernel keeps the rows of V and W a
// CTA owns in shared memory: 6 slots of 32 rows per SM).
int64_t max_order(const aces_ctx* ctx);

// All eigenvalues (ascending) of the symmetric n x n matrix held IN FULL (both triangles) in the Np x Np
// device buffer A (leading dimension lda >= Np, Np a multiple of 128; whatever the padding rows / columns
// hold is cleared).  A is destroyed.  w_dev: n doubles on the device.  Asynchronous on ctx->stream.
int syev_values(aces_ctx* ctx, int64_t n, int64_t Np, double* A, int64_t lda, double* w_dev);

// lower triangle <- upper triangle of the n x n block
int mirror_upper(aces_ctx* ctx, int64_t n, double* A, int64_t lda);

}  // namespace eigen
