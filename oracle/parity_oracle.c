/*
 * parity_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's shot-record -> circuit-eigenvalue reduction.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libaces_b200.so) never links or calls it.
 *
 * Each function follows the reference line by line (paths relative to the QuantumACES.jl
 * repository root):
 *   oracle_pauli_estimate      src/estimate.jl:239-254  pauli_estimate(::Matrix{UInt8}, ...)
 *   oracle_estimate_experiment src/estimate.jl:284-301  estimate_experiment (@threads :static
 *                                                       over Paulis == omp schedule(static))
 *   oracle_simulate_estimate   src/simulate.jl:206-213  one estimate_experiment pass per budget
 *                                                       prefix over the SAME sample matrix
 *   oracle_batch_shots         src/utils.jl:6-29        batch_shots
 *
 * Parity pin: the reference is Julia and cannot run in this image (no julia binary), and it
 * ships no golden files for this path.  The oracle is pinned against the reference's own
 * known-answer tests restated in tests/test_oracle.py: bit addressing round trip
 * (src/utils.jl:477-503, test/device_tests.jl:45-54), the noiseless "all eigenvalues == 1.0"
 * test (test/device_tests.jl:135-156) and the batch_shots assertions (src/utils.jl:26-27).
 * It has NOT been compared with outputs of the Julia package itself.
 */
#include <stdint.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Julia's fld for a positive divisor. */
static inline int64_t fld64(int64_t a, int64_t b) {
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}
static inline int64_t cld64(int64_t a, int64_t b) { return -fld64(-a, b); }

/* src/estimate.jl:239-254.  counts is column-major (shots x bytes), ld = size(counts, 1);
 * pauli_meas_indices are 1-based qubit indices.  Returns pauli_shots in [-S, S]. */
int64_t oracle_pauli_estimate(const uint8_t* counts, int64_t ld,
                              const int32_t* pauli_meas_indices, int32_t weight,
                              int64_t shots_value) {
    int64_t pauli_shots = 0;
    for (int64_t s = 1; s <= shots_value; ++s) {
        int64_t pauli_shot = 0;
        for (int32_t i = 0; i < weight; ++i) {
            int64_t m = pauli_meas_indices[i];
            /* stim_counts[s, 1 + fld(m - 1, 8)] >> ((m - 1) % 8) & 1 */
            uint8_t byte = counts[(s - 1) + fld64(m - 1, 8) * ld];
            pauli_shot += (byte >> ((m - 1) % 8)) & 1;
        }
        pauli_shots += 1 - 2 * (pauli_shot % 2);
    }
    return pauli_shots;
}

/* src/estimate.jl:284-301.  est[idx] = ((-1)^sign * pauli_shots) / shots_value. */
void oracle_estimate_experiment(const uint8_t* counts, int64_t ld, int32_t E,
                                const int64_t* meas_ptr, const int32_t* meas_indices,
                                const uint8_t* signs, int64_t shots_value, double* est,
                                int64_t* pauli_shots_out) {
#pragma omp parallel for schedule(static)
    for (int32_t idx = 0; idx < E; ++idx) {
        int64_t pauli_shots = oracle_pauli_estimate(
            counts, ld, meas_indices + meas_ptr[idx],
            (int32_t)(meas_ptr[idx + 1] - meas_ptr[idx]), shots_value);
        int64_t sgn = signs[idx] ? -1 : 1; /* (-1)^Bool */
        est[idx] = (double)(sgn * pauli_shots) / (double)shots_value;
        if (pauli_shots_out) pauli_shots_out[idx] = pauli_shots;
    }
}

/* src/simulate.jl:206-213: for each budget s, estimate_experiment on the first
 * shots_divided[i, s] rows.  est is [K][E] (k-major), odd is [K][E] odd-parity counts
 * (= (shots - pauli_shots) / 2), either may be NULL. */
void oracle_simulate_estimate(const uint8_t* counts, int64_t ld, int32_t E,
                              const int64_t* meas_ptr, const int32_t* meas_indices,
                              const uint8_t* signs, int32_t K, const int64_t* prefix_shots,
                              double* est, int64_t* odd) {
    double* est_k = (double*)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    int64_t* ps_k = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E > 0 ? E : 1));
    for (int32_t k = 0; k < K; ++k) {
        int64_t S = prefix_shots[k];
        oracle_estimate_experiment(counts, ld, E, meas_ptr, meas_indices, signs, S, est_k, ps_k);
        for (int32_t p = 0; p < E; ++p) {
            if (est) est[(int64_t)k * E + p] = est_k[p];
            if (odd) odd[(int64_t)k * E + p] = (S - ps_k[p]) / 2;
        }
    }
    free(est_k);
    free(ps_k);
}

/* src/utils.jl:6-29.  Writes the batch sizes into out (capacity cap) and returns their
 * number, or -1 if cap is too small / an assertion of the reference would fire. */
int64_t oracle_batch_shots(int64_t shots, int64_t measurements, int64_t max_samples,
                           int64_t* out, int64_t cap) {
    const int64_t base = 256;
    int64_t batch_num = cld64(measurements * shots, max_samples);
    int64_t base_mult = cld64(fld64(shots, base) + 1, batch_num);
    if (batch_num > cap) return -1;
    for (int64_t i = 0; i < batch_num; ++i) out[i] = base_mult * base;
    int64_t excess_shots = base_mult * base * batch_num - shots;
    if (excess_shots < 0) return -1;
    int64_t excess_batch_num = fld64(excess_shots, base);
    int64_t remainder = excess_shots % base;
    if (remainder > 0) {
        int64_t j = batch_num - 1 - excess_batch_num; /* shot_batches[end - excess_batch_num] */
        if (j < 0) return -1;
        out[j] -= remainder;
    }
    for (int64_t idx = 0; idx < excess_batch_num; ++idx) {
        int64_t j = batch_num - 1 - idx;
        if (j < 0) return -1;
        out[j] -= base;
    }
    int64_t n = 0, total = 0;
    for (int64_t i = 0; i < batch_num; ++i)
        if (out[i] > 0) { out[n++] = out[i]; total += out[n - 1]; }
    if (total != shots) return -1;
    return n;
}

int32_t oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Lets the harness use every host core even when the launcher exported OMP_NUM_THREADS=1
 * (torchrun does): the reference runs with JULIA_NUM_THREADS = nproc. */
void oracle_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
