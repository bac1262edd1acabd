/*
 * aces_b200.h — C ABI of libaces_b200.so: the B200 (sm_100a) implementation of the
 * ACES estimation hot path of QuantumACES.jl.
 *
 * The reference has no FFI for this path (it is plain Julia); the entry points below are
 * what a Julia shim binds with `ccall` (see INTEGRATION.md and
 * quantumaces.jl_b200/julia/ACESB200.jl) and what the Python host mirror binds with
 * ctypes (quantumaces.jl_b200/_lib.py).  Every entry point cites the reference code it
 * replaces as file:line relative to the reference repository root.
 *
 * Conventions
 *  - All functions return 0 on success, a negative ACES_E_* code on failure; the message
 *    is available from aces_last_error().
 *  - The caller owns every buffer it passes in and keeps it alive across the call
 *    (Julia: GC.@preserve).  The library owns device memory and never keeps a host
 *    pointer after returning.
 *  - Pointers marked "host or device" are classified with cudaPointerGetAttributes.
 *  - Calls are blocking unless stated otherwise and serialised per context; use one
 *    context per GPU (one process per GPU for multi-GPU runs).
 *  - There is no CPU fallback anywhere in this library.
 */
#ifndef ACES_B200_H
#define ACES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACES_OK 0
#define ACES_E_INVALID (-1)    /* bad argument */
#define ACES_E_CUDA (-2)       /* CUDA runtime/driver error */
#define ACES_E_NOMEM (-3)      /* allocation failure */
#define ACES_E_NOT_PD (-4)     /* a covariance block / Gram matrix is not positive definite
                                  (reference fallback: src/merit.jl:403-424) */
#define ACES_E_UNSUPPORTED (-5)

typedef struct aces_ctx aces_ctx;
typedef struct aces_tables aces_tables;

/* ---- context ------------------------------------------------------------------------- */

/* Library ABI version (major*1000 + minor). */
int32_t aces_version(void);

/* Creates a context on CUDA device `device` (streams, workspaces).  Fails with
 * ACES_E_CUDA when no sm_100 device is present: there is no CPU path. */
int32_t aces_init(int32_t device, aces_ctx** out);
int32_t aces_destroy(aces_ctx* ctx);

/* Last error message of `ctx` (or of the last failed aces_init when ctx == NULL). */
const char* aces_last_error(const aces_ctx* ctx);

/* Run the library's kernels and copies on `cuda_stream` (a cudaStream_t) instead of the
 * context's own stream; NULL restores the context's stream.  Lets a host that owns a
 * stream (e.g. torch) time and order the work. */
int32_t aces_set_stream(aces_ctx* ctx, void* cuda_stream);
int32_t aces_synchronize(aces_ctx* ctx);

/* Pinned host memory for sample matrices that are uploaded every call. */
int32_t aces_host_alloc(aces_ctx* ctx, int64_t bytes, void** out);
int32_t aces_host_free(aces_ctx* ctx, void* ptr);

/* Number of kernels this context has launched since creation (bench bookkeeping). */
int64_t aces_kernel_launches(const aces_ctx* ctx);

/* Rank deficiency of the last fit on this context (aces_estimate_gate_eigenvalues, aces_design_fit,
 * aces_design_fit_end): the number of gate eigenvalues whose column of the whitened design matrix was
 * numerically dependent on the columns before it (pivot below 1e-13 of the Gram diagonal), e.g. after
 * clipping (src/estimate.jl:435-460) removed every row that determines them.  Such a fit does not
 * fail: like the reference's sparse-QR solve (src/estimate.jl:500-502) it returns a BASIC solution --
 * the dependent columns get x = 0 (gate eigenvalue 1) and the others solve the reduced problem.
 * Which columns a basic solution zeroes depends on the elimination order (SPQR's fill-reducing column
 * ordering in the reference, left to right here), so for a rank-deficient fit only the fitted circuit
 * eigenvalues exp(-A x) agree with the reference, not the individual gate eigenvalues. */
int32_t aces_last_rank_deficiency(const aces_ctx* ctx);

/* With timing on, the dominant kernel of each call (K1: the parity kernel) is bracketed by
 * CUDA events on the stream it is launched on; aces_last_kernel_ms waits for the last one
 * and returns its duration.  Used by bench.py for the roofline figure. */
int32_t aces_set_timing(aces_ctx* ctx, int32_t on);
int32_t aces_last_kernel_ms(aces_ctx* ctx, float* ms);

/* ---- K1: shot records -> odd-parity counts --------------------------------------------
 *
 * Replaces pauli_estimate + estimate_experiment (src/estimate.jl:239-301) and the
 * per-budget inner loop of simulate_stim_estimate (src/simulate.jl:195-234).
 *
 * A *Pauli table* lists, per experiment, the measured Paulis as lists of record bits.
 * Record bit b (0-based) is bit (b % 8) of byte column (b / 8) of a shot record, i.e.
 * qubit m = b + 1 of src/estimate.jl:249.  It is what get_experiment_data /
 * get_covariance_experiment_data (src/estimate.jl:308-428) produce from a Design, flattened
 * once per Design:
 *   exp_ptr   [n_experiments + 1]  Paulis of experiment e are exp_ptr[e] .. exp_ptr[e+1]-1
 *   pauli_ptr [n_paulis + 1]       support of Pauli p is pauli_bits[pauli_ptr[p] .. pauli_ptr[p+1]-1]
 *   pauli_bits                     0-based record bits, each < n_qubits
 * A bit listed twice in one support cancels (as in the reference, where the per-shot sum is
 * reduced mod 2).  An empty support is the identity: its odd count is 0.
 */
int32_t aces_tables_create(aces_ctx* ctx, int32_t n_qubits, int32_t n_experiments,
                           const int64_t* exp_ptr, const int64_t* pauli_ptr,
                           const int32_t* pauli_bits, aces_tables** out);
int32_t aces_tables_destroy(aces_ctx* ctx, aces_tables* tables);
/* Byte columns the kernel reads per shot of `experiment`: the sum over its column groups (an
 * experiment whose Paulis fit one kernel pass is one group over all ceil(n_qubits / 8) columns;
 * wide records and large experiments are split into groups that read only the columns their
 * Paulis touch, a column shared by two groups being read twice).  Bookkeeping for the roofline. */
int64_t aces_tables_group_columns(const aces_tables* tables, int32_t experiment);

/* Odd-parity counts of a batch of sample matrices in ONE kernel launch.
 *
 *   item i: experiment experiment_ids[i]; sample matrix counts[i] of rows[i] shots,
 *           column-major bytes, ld[i] bytes between byte columns (Julia Matrix{UInt8}:
 *           ld = size(counts, 1)), at least ceil(n_qubits / 8) columns; host or device.
 *   prefix_shots[i*K + k], k < K: nondecreasing shot counts <= rows[i]; count k is taken
 *           over rows 0 .. prefix_shots[i*K+k]-1 (src/simulate.jl:206-213).
 *   odd_counts (host or device): for item i a block of E_i * K int64 placed after the
 *           blocks of items 0..i-1; entry [k * E_i + p] = number of shots among the first
 *           prefix_shots[i*K+k] whose parity over Pauli p's record bits is odd.
 *           The reference's pauli_shots (src/estimate.jl:251) is shots - 2 * odd.
 *   accumulate != 0: add to the values already in odd_counts (batched sampling,
 *           src/simulate.jl:174-186, without concatenating the batches).
 *
 * Device matrices with 16-byte aligned base and ld % 16 == 0 are read in place through the
 * TMA engine; anything else is first copied into a pitched device buffer.
 */
int32_t aces_parity_counts(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                           const int32_t* experiment_ids, const uint8_t* const* counts,
                           const int64_t* rows, const int64_t* ld, int32_t K,
                           const int64_t* prefix_shots, int64_t* odd_counts,
                           int32_t accumulate);

/* Same launch as aces_parity_counts but asynchronous on the context's stream.  By contract every
 * matrix and odd_counts are DEVICE pointers (not checked: the per-pointer query would dominate the
 * host cost of a call; a host pointer here faults in the kernel).  Used to time the kernel with
 * CUDA events and to feed NCCL without a host round trip. */
int32_t aces_parity_counts_async(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                                 const int32_t* experiment_ids, const uint8_t* const* counts,
                                 const int64_t* rows, const int64_t* ld, int32_t K,
                                 const int64_t* prefix_shots, int64_t* odd_counts,
                                 int32_t accumulate);

/* Ingest: the same counts from ROW-major records, i.e. the buffer Stim's
 * sampler.sample(shots, bit_packed = true) returns before the reference's pyconvert(Matrix{UInt8}, ...)
 * turns it into a column-major Julia matrix with a strided host copy (src/simulate.jl:57-63):
 *   records[i]  uint8 [rows[i] x B], shot s at records[i] + s * row_pitch[i] (row_pitch >= B =
 *               ceil(n_qubits / 8); numpy C order: row_pitch = B), host or device.
 * One contiguous copy per item, a device-side byte transposition into the column-major arena, then
 * the kernel of aces_parity_counts; every other argument as there. */
int32_t aces_parity_counts_rowmajor(aces_ctx* ctx, const aces_tables* tables, int32_t n_items,
                                    const int32_t* experiment_ids, const uint8_t* const* records,
                                    const int64_t* rows, const int64_t* row_pitch, int32_t K,
                                    const int64_t* prefix_shots, int64_t* odd_counts, int32_t accumulate);

/* Drop-in for estimate_experiment(counts::Matrix{UInt8}, experiment_meas_indices,
 * experiment_pauli_signs, shots_value) (src/estimate.jl:284-301): meas_indices are the
 * reference's 1-based qubit indices in CSR form (meas_ptr [E+1]); est[p] =
 * ((-1)^sign[p] * (shots_value - 2*odd[p])) / shots_value, the division done in IEEE
 * double on the host exactly as the reference does. */
int32_t aces_estimate_experiment(aces_ctx* ctx, const uint8_t* counts, int64_t rows,
                                 int64_t ld, int32_t n_cols, int32_t E,
                                 const int64_t* meas_ptr, const int32_t* meas_indices,
                                 const uint8_t* signs, int64_t shots_value, double* est);

/* ---- K4/K5: least-squares fit of the gate eigenvalues -------------------------------------------
 *
 * Drop-in for estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64,Int},
 * est_covariance_log_inv_factor::SparseMatrixCSC{Float64,Int}, est_eigenvalues; constrain)
 * (src/estimate.jl:492-524) and its two-argument OLS form (:510-524, pass F_* = NULL):
 *     b = -log(est_eigenvalues);  x = (F*A) \ (F*b);  constrain: x[x < 0] = 0;  g = exp(-x).
 * A is M x N, F is M x M, both CSC exactly as Julia stores them (colptr [ncols+1], rowval,
 * nzval); index_base = 1 for Julia's 1-based colptr/rowval, 0 for C.  The whitened matrix F*A is
 * formed densely on the device, its Gram matrix accumulated on the FP64 tensor pipe, factorised
 * by a blocked Cholesky and the solution refined twice with FP64 residuals.  A numerically rank
 * deficient F*A gives a basic solution (see aces_last_rank_deficiency). */
int32_t aces_estimate_gate_eigenvalues(aces_ctx* ctx, int64_t M, int64_t N, const int64_t* A_colptr,
                                       const int64_t* A_rowval, const double* A_nzval,
                                       const int64_t* F_colptr, const int64_t* F_rowval,
                                       const double* F_nzval, int32_t index_base,
                                       const double* est_eigenvalues, int32_t constrain,
                                       double* gate_eigenvalues);

/* Dense N x N (column-major, host or device) matrices derived from the Gram matrix
 * G = (F*A)' (F*A) of the same arguments:
 *   want_gram == 0: out = S * inv(G) * S   -- calc_gls_covariance (src/merit.jl:556-568) with
 *                   S = diag(scale) = diag(gate eigenvalues) and F'F = inv(covariance_log);
 *   want_gram != 0: out = S * G * S        -- calc_precision_matrix (src/merit.jl:754-770) with
 *                   scale_inverse != 0 (S = diag(1 ./ scale)).
 * scale == NULL means S = I. */
int32_t aces_gram_inverse(aces_ctx* ctx, int64_t M, int64_t N, const int64_t* A_colptr,
                          const int64_t* A_rowval, const double* A_nzval, const int64_t* F_colptr,
                          const int64_t* F_rowval, const double* F_nzval, int32_t index_base,
                          const double* scale, int32_t scale_inverse, int32_t want_gram, double* out);

/* ---- design-resident fit path ---------------------------------------------------------------------
 *
 * Everything the numeric core of estimate_gate_noise (src/estimate.jl:703-743, 820-831) and the
 * estimator covariances of src/merit.jl:556-629 need, with the Design uploaded once instead of
 * per call.  An aces_design holds what those functions read from a `Design`:
 *   mapping_lengths [T]        length.(d.mapping_ensemble)                    (block sizes m_t)
 *   A_colptr/A_rowval/A_nzval  d.matrix exactly as stored: SparseMatrixCSC{Int32,Int32}, M x N
 *   experiment_numbers [T], shot_weights [T], tuple_times [T]
 *   diag_counts [M]            covariance_dict_ensemble[t][(p,p)][2]  (src/design.jl:803-817)
 *   cov_ptr [T+1], cov_p/cov_q/cov_counts [C]
 *                              the off-diagonal keys of covariance_dict_ensemble[t] in the sorted
 *                              order of get_covariance_mapping_data (src/merit.jl:144-163), as
 *                              block-local Pauli indices, with their experiment counts; the c-th
 *                              entry is the c-th covariance circuit eigenvalue.  cov_ptr == NULL:
 *                              diagonal covariance (full_covariance = false).
 * index_base = 1 when colptr / rowval / cov_p / cov_q are Julia's 1-based indices, 0 for C.
 *
 * The covariance matrices of this path all share one sparsity pattern (the diagonal plus both
 * orientations of every key); they are exchanged as the nzval array of a CSC matrix whose
 * colptr / rowval come from aces_design_covariance_pattern (0-based, rows sorted).
 */
typedef struct aces_design aces_design;

int32_t aces_design_create(aces_ctx* ctx, int32_t T, const int64_t* mapping_lengths, int64_t N,
                           const int32_t* A_colptr, const int32_t* A_rowval, const int32_t* A_nzval,
                           int32_t index_base, const int64_t* experiment_numbers,
                           const double* shot_weights, const double* tuple_times,
                           const int64_t* diag_counts, const int64_t* cov_ptr, const int64_t* cov_p,
                           const int64_t* cov_q, const int64_t* cov_counts, aces_design** out);
int32_t aces_design_destroy(aces_ctx* ctx, aces_design* design);
int64_t aces_design_covariance_nnz(const aces_design* design);
int32_t aces_design_covariance_pattern(const aces_design* design, int64_t* colptr /* [M+1] */,
                                       int64_t* rowval /* [nnz] */);

/* K2.  calc_covariance(d, eigenvalues_ensemble, covariance_ensemble; epsilon, weight_time)
 * (src/merit.jl:206-291): eigenvalues [M] (tuples concatenated), covariance_eigenvalues [C].
 * log_form != 0 additionally applies calc_covariance_log(covariance, eigenvalues)
 * (src/merit.jl:356-365) with the same eigenvalues.  nzval [nnz] host or device. */
int32_t aces_calc_covariance(aces_ctx* ctx, aces_design* design, const double* eigenvalues,
                             const double* covariance_eigenvalues, double epsilon,
                             int32_t weight_time, int32_t log_form, double* nzval);

/* K3.  sparse_covariance_inv_factor(covariance_log[kept, kept], clipped_mapping_lengths)
 * (src/merit.jl:378-427, called at src/estimate.jl:823): covariance_log_nzval in the design's
 * pattern; kept_rows (sorted, unique; NULL = all M rows) are the clipped indices of
 * src/estimate.jl:435-460.  factor_blocks receives, for every tuple, the dense lower-triangular
 * block inv(L_t) (k_t x k_t column-major, k_t = kept rows of the tuple, blocks concatenated) with
 * covariance_log_t = L_t L_t'; block_lengths [T] receives k_t.  The reference's CHOLMOD factor is
 * a row permutation of this one; F'F = inv(covariance_log) is the same.
 * A block whose factorisation fails gets the reference's fallback (src/merit.jl:403-424) when
 * fallback_epsilon > 0 (the reference's default epsilon is 1e-12): elementwise absolute value,
 * then the smallest eigenvalue raised to fallback_epsilon by a diagonal shift (found by bisection
 * on the shift instead of ARPACK).  fallback_epsilon <= 0: such a block returns ACES_E_NOT_PD.
 * aces_design_fallback_count reports how many blocks took the fallback in the last call. */
int32_t aces_design_inv_factor(aces_ctx* ctx, aces_design* design, const double* covariance_log_nzval,
                               const int64_t* kept_rows, int64_t n_kept, int32_t index_base,
                               double fallback_epsilon, double* factor_blocks, int64_t* block_lengths);
int32_t aces_design_fallback_count(const aces_design* design);
/* Dense FP64 flops the GLS Gram matrix of this design executes on the tensor pipe per fit (the
 * lower 128 x 128 tiles of sum_t At_t' At_t over the column window of every tuple block, or of the
 * full At' At when the windows do not pay).  Bookkeeping for the roofline figure. */
double aces_design_gram_flops(const aces_design* design);
/* Phases of the last aces_design_fit when timing was on (aces_set_timing): index 0, 1, ... until
 * ACES_E_INVALID; name (e.g. "gram", "potrf", "first solve") and CUDA-event milliseconds. */
int32_t aces_design_phase_ms(const aces_design* design, int32_t index, char* name, int32_t name_len, float* ms);

/* K2-K5 fused: the gate eigenvalues of one estimator from the circuit-eigenvalue estimates, as
 * estimate_gate_noise computes them (src/estimate.jl:711-743, 820-831): covariance of the
 * estimates (weight_time = false), log-covariance, clipping to kept_rows, inverse factor
 * (ls_type 2 = GLS: block Cholesky; 1 = WLS: diagonal; 0 = OLS: identity), whitened
 * least-squares solve, g = exp(-x).  est_eigenvalues [M] are the unclipped means
 * (mean.(est_eigenvalues_experiment_ensemble)), est_covariance_eigenvalues [C] likewise. */
int32_t aces_design_fit(aces_ctx* ctx, aces_design* design, int32_t ls_type,
                        const double* est_eigenvalues, const double* est_covariance_eigenvalues,
                        const int64_t* kept_rows, int64_t n_kept, int32_t index_base, double epsilon,
                        int32_t constrain, double* gate_eigenvalues);

/* Tuple-sharded fit ("one large Gram"): the covariance is block diagonal by
 * tuple (src/merit.jl:284), so the Gram matrix A' Omega'^-1 A and the right-hand side are sums
 * over the tuple blocks.  Every shard holds the whole design but only works on the tuples whose
 * tuple_on byte is non-zero (NULL = all: the design is no shard).  aces_design_fit is these stages
 * back to back; a sharded fit sums the two device buffers of aces_design_fit_buffers over the
 * shards after _begin and the right-hand side again after every _refine_begin (ncclAllReduce,
 * sum, double, in place).  The Cholesky factorisation and the triangular solves are replicated.
 * No other call on the same context may come between _begin and _end.
 *   begin        : K2 covariance, K3 block factors of the shard's tuples, partial Gram (lower
 *                  triangle of an Np x Np column-major matrix, Np = N padded to 128) and partial
 *                  rhs A' F' F b;
 *   solve        : Cholesky of the summed Gram matrix, x = inv(G) rhs;
 *   refine_begin : partial rhs of the correction, A' F' F (b - A x);
 *   refine_end   : x += inv(G) rhs; returns max|x| and max|delta| (stop when delta <= 1e-7 |x|,
 *                  at most two corrections, as aces_design_fit does);
 *   end          : g = exp(-x) (clamped to <= 1 when constrain), copied to gate_eigenvalues. */
int32_t aces_design_set_tuple_shard(aces_ctx* ctx, aces_design* design, const uint8_t* tuple_on /* [T] or NULL */);
int32_t aces_design_fit_begin(aces_ctx* ctx, aces_design* design, int32_t ls_type,
                              const double* est_eigenvalues, const double* est_covariance_eigenvalues,
                              const int64_t* kept_rows, int64_t n_kept, int32_t index_base, double epsilon);
int32_t aces_design_fit_buffers(aces_design* design, double** gram, int64_t* gram_elems,
                                double** rhs, int64_t* rhs_elems);
int32_t aces_design_fit_solve(aces_ctx* ctx, aces_design* design);
int32_t aces_design_fit_refine_begin(aces_ctx* ctx, aces_design* design);
int32_t aces_design_fit_refine_end(aces_ctx* ctx, aces_design* design, double* x_max, double* delta_max);
int32_t aces_design_fit_end(aces_ctx* ctx, aces_design* design, int32_t constrain, double* gate_eigenvalues);

/* The Cholesky factorisation inside aces_design_fit_solve, distributed over the shards (SURVEY.md section 8e,
 * "one large Gram": at rotated d = 17 the replicated N^3 / 3 factorisation is what limits the sharded fit).
 * After aces_design_fit_begin and the sum of the buffers every shard holds the whole Gram matrix.  Block
 * columns (a multiple of 128 columns each, 512 in the hosts here) are owned cyclically.  For block column p:
 *   the owner calls _dist_factor(col0, ncols): the block column becomes the block column of L, in place,
 *     and the inverses of its 128 x 128 diagonal blocks are stored;
 *   the host broadcasts from the owner the ncols FULL columns -- gram + col0 * Np, ncols * Np doubles, one
 *     contiguous range (Np = N padded to 128, gram from aces_design_fit_buffers) -- and the block inverses,
 *     block_inverses + col0 * 128, ncols * 128 doubles (ncclBroadcast over NVLink);
 *   every shard calls _dist_update(col0, ncols, targets): the panel is applied to the later block columns
 *     this shard owns; columns it does not own are left alone (they arrive final with their owner's broadcast).
 * _dist_begin comes before the first block column (it records the diagonal for the relative pivot test);
 * _dist_solve replaces aces_design_fit_solve (x = inv(G) rhs with the now complete factor).  The count of
 * eliminated (numerically dependent) columns accumulates in *dead_columns on the owner of each block column:
 * sum it over the shards (an int32 on the device) before _dist_solve if aces_last_rank_deficiency matters. */
int32_t aces_design_fit_dist_begin(aces_ctx* ctx, aces_design* design);
int32_t aces_design_fit_dist_factor(aces_ctx* ctx, aces_design* design, int64_t col0, int64_t ncols);
int32_t aces_design_fit_dist_update(aces_ctx* ctx, aces_design* design, int64_t panel_col0, int64_t panel_ncols,
                                    int32_t n_targets, const int64_t* target_col0, const int64_t* target_ncols);
int32_t aces_design_fit_dist_buffers(aces_design* design, double** block_inverses, int64_t* block_inverse_elems,
                                     int32_t** dead_columns);
int32_t aces_design_fit_dist_solve(aces_ctx* ctx, aces_design* design);

/* ---- count exchange of the sharded K1 path over peer memory -----------------------
 *
 * Repetitions, experiments and shot ranges shard over the GPUs of a node with one small exchange per step: every
 * rank needs the odd-parity counts of all ranks (side by side for repetitions, summed for experiments / shots;
 * int64, 0.4 - 1.4 MB).  Instead of a collective library call between launches, a rank PUSHES its counts with a
 * kernel that stores them into a slot of every rank's exchange buffer over NVLink and releases a flag there, and
 * PULLS with a kernel that waits for the flags of all ranks and sums (or copies) the slots.  The push goes on
 * the stream of the parity kernel, the pull on a second stream: the next repetition's kernel is not held up.
 *   aces_exchange_bytes   size of one rank's buffer for `world` ranks of n_counts counts each
 *   aces_peer_alloc       device buffer (zeroed) + its 64-byte CUDA IPC handle, to be handed to the other ranks
 *   aces_peer_open/_close map / unmap another rank's buffer from its handle (another process of this node)
 *   aces_exchange_create  peer_base[r] = rank r's buffer as mapped HERE (own buffer: the pointer of aces_peer_alloc)
 *   aces_exchange_push    local_counts: device, 16-byte aligned; cuda_stream NULL = the context's stream
 *   aces_exchange_pull    reduce != 0: out[n_counts] = sum over ranks; else out[world][n_counts]; device pointer
 *   aces_exchange_attach  fuses the push into K1: while an exchange is attached to the context (NULL detaches),
 *                         the prefix kernel of every aces_parity_counts_async call stores each count into this
 *                         rank's slot of every rank's buffer as it forms it and releases the flags itself -- no
 *                         separate push (the call must produce exactly n_counts counts, accumulate = 0)
 *   aces_exchange_status  ACES_E_CUDA if a pull gave up without a peer's flag (blocking: reads a flag); the wait is
 *                         bounded (about 10 s by default, aces_exchange_set_timeout) so that a dead peer cannot hang the GPU
 * Every rank must call push and pull once per step, in the same order, and must not push exchange e before its own
 * pull of exchange e - 2 has completed (four slot sets rotate; stream order or an event gives that for free).  Sums of integers: exact, identical on
 * every rank. */
typedef struct aces_exchange aces_exchange;
int64_t aces_exchange_bytes(int32_t world, int64_t n_counts);
int32_t aces_peer_alloc(aces_ctx* ctx, int64_t bytes, void** dev_ptr, uint8_t* handle_out /* [64] */);
int32_t aces_peer_open(aces_ctx* ctx, const uint8_t* handle /* [64] */, void** dev_ptr);
int32_t aces_peer_close(aces_ctx* ctx, void* dev_ptr);
int32_t aces_peer_free(aces_ctx* ctx, void* dev_ptr);
int32_t aces_exchange_create(aces_ctx* ctx, int32_t world, int32_t rank, int64_t n_counts, void* const* peer_base,
                             aces_exchange** out);
int32_t aces_exchange_destroy(aces_ctx* ctx, aces_exchange* exchange);
int32_t aces_exchange_push(aces_ctx* ctx, aces_exchange* exchange, const int64_t* local_counts, void* cuda_stream);
int32_t aces_exchange_pull(aces_ctx* ctx, aces_exchange* exchange, int32_t reduce, int64_t* out, void* cuda_stream);
int32_t aces_exchange_attach(aces_ctx* ctx, aces_exchange* exchange);
int32_t aces_exchange_set_timeout(aces_ctx* ctx, aces_exchange* exchange, int64_t milliseconds);
int32_t aces_exchange_status(aces_ctx* ctx, aces_exchange* exchange);

/* Per-gate diagonal blocks of calc_precision_matrix(design_matrix, gate_eigenvalues, covariance_log_inv)
 * (src/merit.jl:754-770) = D^-1 A' Omega'^-1 A D^-1 -- all that split_project_gate_eigenvalues reads of
 * it (src/estimate.jl:606-612: est_precision_matrix[indices, indices] per gate).  Must directly follow
 * aces_design_fit of the same ls_type on this design: the blocks are those of that fit's Gram matrix
 * (its clipping, weights and block factors), rebuilt on the device.  Group g owns
 * group_idx[group_ptr[g] .. group_ptr[g+1]-1]; blocks receives the n_g x n_g blocks concatenated
 * (symmetric).  gate_eigenvalues [N] = the unprojected estimate D is made of (NULL: D = I). */
int32_t aces_design_precision_blocks(aces_ctx* ctx, aces_design* design, int32_t ls_type, int32_t G,
                                     const int64_t* group_ptr, const int32_t* group_idx, int32_t index_base,
                                     const double* gate_eigenvalues, double* blocks);

/* K6.  calc_gls_covariance / calc_wls_covariance / calc_ols_covariance(d, covariance_log)
 * (src/merit.jl:556-629) for ls_type 2 / 1 / 0: gate_eigenvalues [N] = get_gate_eigenvalues(d),
 * covariance_log_nzval in the design's pattern.  sigma (N x N column-major, host or device) may be
 * NULL when only the moments are wanted: trace = tr(sigma), trace_sq = tr(sigma^2), the inputs
 * of nrmse_moments (src/merit.jl:672-682). */
int32_t aces_design_ls_covariance(aces_ctx* ctx, aces_design* design, int32_t ls_type,
                                  const double* gate_eigenvalues, const double* covariance_log_nzval,
                                  double* sigma, double* trace, double* trace_sq);

/* the figure of merit of a design at the shot weights `shot_weights` and its gradient
 * with respect to the LOGARITHMS of the shot weights -- calc_gls_merit_grad_log /
 * calc_wls_merit_grad_log / calc_ols_merit_grad_log (src/optimise_weights.jl:103-152, 327-391,
 * 563-606) for ls_type 2 / 1 / 0, the inner step of gls_ / wls_ / ols_optimise_weights
 * (src/optimise_weights.jl:160-321, 398-555, 613-790).
 *   covariance_log_unweighted_nzval  covariance_log * get_shot_weights_factor_inv(d.shot_weights, ...)
 *                                    (src/optimise_weights.jl:183-184) in the design's pattern (symmetric;
 *                                    the GLS method of the reference takes its block inverse, which is
 *                                    formed here)
 *   shot_weights, tuple_times [T]    the (Nesterov) shot weights of this step and d.tuple_times
 *   tt_colptr / tt_rowval / tt_nzval T'T for gate_transform_matrix T = get_transform(d, est_type) *
 *                                    Diagonal(get_gate_eigenvalues(d)) (n_transformed x N): N x N CSC,
 *                                    symmetric, both triangles stored
 *   n_transformed                    size(gate_transform_matrix, 1), the N of the merit's normalisation
 * Outputs (any may be NULL): merit, merit_grad_log [T], sigma_tr = tr(Sigma), sigma_sq_tr = tr(Sigma^2).
 * Nothing larger than N x N is formed: the reference's M x N / M x M intermediates only enter through
 * per-tuple sums that are inner products with N x N matrices (csrc/merit_grad.cuh). */
int32_t aces_design_merit_grad(aces_ctx* ctx, aces_design* design, int32_t ls_type,
                               const double* covariance_log_unweighted_nzval, const double* shot_weights,
                               const double* tuple_times, int64_t n_transformed, const int64_t* tt_colptr,
                               const int64_t* tt_rowval, const double* tt_nzval, int32_t index_base,
                               double* merit, double* merit_grad_log, double* sigma_tr, double* sigma_sq_tr);

/* ---- calc_merit tail ------------------------------------------------------
 *
 * All eigenvalues (ascending) of the symmetric n x n matrix a -- drop-in for
 * LinearAlgebra.eigvals(::Symmetric{Float64, Matrix{Float64}}) (LAPACK dsyevr, JOBZ = 'N') as calc_merit
 * calls it nine times per design (src/merit.jl:954-986).  a: column-major, leading dimension lda, host or
 * device; like Julia's Symmetric (uplo = :U) only the UPPER triangle is read.  Householder
 * tridiagonalisation (panels of 64 reflectors in one persistent kernel each, trailing update on the FP64
 * tensor pipe) and bisection on Sturm counts; absolute accuracy eps * ||a|| like LAPACK's.  Orders up to
 * 192 rows per SM (28 416 on a B200) are supported. */
int32_t aces_symmetric_eigenvalues(aces_ctx* ctx, int64_t n, const double* a, int64_t lda, double* eigenvalues);

/* The numeric core of calc_merit(d::Design) (src/merit.jl:925-1006) for a design with full covariance data:
 * for ls_type GLS, WLS, OLS (in this order) the estimator covariance Sigma of src/merit.jl:556-629 (as
 * aces_design_ls_covariance), its marginal and relative transforms T Sigma T' (get_marginal_gate_covariance /
 * get_relative_gate_covariance, src/merit.jl:477-509) and the eigenvalues of all three, ascending.
 *   gate_eigenvalues [N], covariance_log_nzval   as for aces_design_ls_covariance
 *   marg_* / rel_*   get_transform(d, :marginal) / (d, :relative) (src/noise.jl:689-741) as Julia stores them:
 *                    SparseMatrixCSC n_marginal x N / n_relative x N (colptr [N+1], rowval, nzval)
 *   cov_eigenvalues  [3][N + n_marginal + n_relative]: per estimator the ordinary, marginal and relative spectra
 *                    (the Merit fields *_cov_eigenvalues; nrmse_moments(::Vector) of src/merit.jl:684-692 is
 *                    two sums over them and stays with the host)
 *   gram_extremes    [2] (may be NULL): largest and smallest eigenvalue of d.matrix' * d.matrix, whose square
 *                    roots give design_matrix_cond_num and design_matrix_pinv_norm (src/merit.jl:993-1006; the
 *                    reference asks ARPACK for the two extremes, here the whole spectrum is computed). */
int32_t aces_design_merit(aces_ctx* ctx, aces_design* design, const double* gate_eigenvalues,
                          const double* covariance_log_nzval, int64_t n_marginal, const int64_t* marg_colptr,
                          const int64_t* marg_rowval, const double* marg_nzval, int64_t n_relative,
                          const int64_t* rel_colptr, const int64_t* rel_rowval, const double* rel_nzval,
                          int32_t index_base, double* cov_eigenvalues, double* gram_extremes);

/* Bookkeeping / test hook: the static task order of the persistent Cholesky kernel for an nb x nb tile
 * matrix (128 x 128 tiles) scheduled on `workers` CTAs.  out (may be NULL to query the count) receives
 * 6 ints per task: type (0 factor, 1 panel solve, 2 update, 3 / 4 their 16-row slabs), slab (+ 256 for a
 * task of the near list), i, j, k0, k1; order: far list, near list, the nb factor tasks of the chain CTA.
 * No GPU needed. */
int32_t aces_potrf_schedule(int32_t nb, int32_t workers, int32_t* out, int64_t capacity, int64_t* n_tasks);

/* Development hook: with ACES_POTRF_TRACE=1 in the environment the persistent Cholesky kernel records
 * per-task timestamps; this returns 9 int64 per task of the last factorisation on ctx (type, slab, i, j,
 * k0, k1, t_grab, t_ready, t_done in ns).  *n_tasks = 0 when tracing is off. */
int32_t aces_potrf_trace(aces_ctx* ctx, int64_t* out, int64_t capacity, int64_t* n_tasks);

/* ---- simplex projection of the fitted gate eigenvalues ----------------
 *
 * A gate is a group of 1 (SPAM / mid-circuit reset), 3 (one qubit) or 15 (two qubits) gate
 * eigenvalues: group g owns group_idx[group_ptr[g] .. group_ptr[g+1]-1] (GateIndex.indices of
 * src/noise.jl:16-24, in the Pauli order of the symplectic Walsh-Hadamard matrix, src/noise.jl:285-312).
 * The padded vectors (identity entry first: GateIndex.pad_indices) are laid out gate after gate, so
 * gate g starts at group_ptr[g] + g.
 *
 * aces_project_simplex: drop-in for simple_project_gate_eigenvalues (src/estimate.jl:635-663), i.e.
 * per gate p = W^-1 [1; lambda], project_simplex(p) (src/utils.jl:36-50, sort based), lambda' = W p'.
 * The arithmetic follows the reference's order of operations: the results are bit-identical.
 *   gate_eigenvalues [N]; unproj_probabilities / probabilities [N + G] (may be NULL). */
int32_t aces_project_simplex(aces_ctx* ctx, int64_t N, int32_t G, const int64_t* group_ptr, const int32_t* group_idx,
                             int32_t index_base, const double* unproj_gate_eigenvalues, double* gate_eigenvalues,
                             double* unproj_probabilities, double* probabilities);

/* Batched drop-in for scs_project_nonnegative(vector, precision_matrix; precision) (src/utils.jl:57-97)
 * as split_project_gate_eigenvalues calls it once per gate (src/estimate.jl:606-618):
 *     y_p = argmin (y - v_p)' Q_p (y - v_p)  subject to  y >= 0,   then y_p[y_p < threshold] = 0.
 * Problem p has n_p = ptr[p+1] - ptr[p] <= 15 variables; v and y are the concatenated vectors, Q the
 * concatenated dense symmetric n_p x n_p blocks.  Solved exactly (active-set method) where the
 * reference iterates SCS to eps = threshold = 1e-8: results agree to that tolerance.  A block that is
 * not positive definite returns ACES_E_NOT_PD (their number in *n_failed). */
int32_t aces_project_nonnegative_batched(aces_ctx* ctx, int32_t n_problems, const int64_t* ptr, const double* Q,
                                         const double* v, double threshold, double* y, int32_t* n_failed);

/* Drop-in for the numeric core of full_project_gate_eigenvalues(est_unproj_gate_eigenvalues, est_precision_matrix,
 * gate_data; precision) (src/estimate.jl:532-573), the simultaneous Mahalanobis projection of all gates -- the
 * reference's default below 5000 gate eigenvalues (split = false).  Must directly follow aces_design_fit of the
 * same ls_type on this design: the precision matrix D^-1 A' Omega'^-1 A D^-1 (src/merit.jl:754-770) is that of
 * the fit's Gram matrix, rebuilt on the device; it never leaves it.
 *   group_ptr / group_idx  the gates (as for aces_design_precision_blocks); they must cover the N gate eigenvalues
 *   gate_eigenvalues [N]   the unprojected estimate (the D of the precision matrix)
 *   transform_blocks       per gate the n_g x n_g block of probs_transform = pad' * W * pad_probs
 *                          (src/estimate.jl:553), row-major, gate after gate: T[a][b] = W[a+1][b+1] - W[a+1][0]
 *   v [N]                  the unprojected unpadded gate probabilities (src/estimate.jl:548-551)
 *   y [N]                  argmin (y - v)' (T P T') (y - v) over y >= 0, entries below `threshold` set to 0
 *                          (scs_project_nonnegative, src/utils.jl:57-97, which iterates SCS to eps = threshold = 1e-8)
 * The quadratic program is solved exactly by block principal pivoting: a few Cholesky factorisations of the free
 * block on the FP64 tensor pipe; *iterations receives their number.  The caller maps y back to padded
 * probabilities and gate eigenvalues (pad_probs * y + mask, W *, pad' *) as the reference does. */
int32_t aces_design_project_full(aces_ctx* ctx, aces_design* design, int32_t ls_type, int32_t G, const int64_t* group_ptr,
                                 const int32_t* group_idx, int32_t index_base, const double* gate_eigenvalues,
                                 const double* transform_blocks, const double* v, double threshold, double* y,
                                 int32_t* iterations);

#ifdef __cplusplus
}
#endif
#endif /* ACES_B200_H */
