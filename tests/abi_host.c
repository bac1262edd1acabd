/* abi_host.c -- drives libaces_b200.so from plain C exactly as the Julia shim (ACESB200.jl) does:
 * dlopen, Julia-style 1-based index arrays (index_base = 1), pageable malloc'ed host buffers, the
 * call sequences of simulate_stim_estimate / estimate_gate_noise / calc_covariance.  Inputs and the
 * expected results (from the CPU oracle) come from a fixture file written by tests/test_abi_host.py.
 *
 *   abi_host <libaces_b200.so> --symbols            resolve every entry point (no GPU needed)
 *   abi_host <libaces_b200.so> <fixture.bin>        run the sequences on cuda:0, compare, print PASS / FAIL
 *
 * Fixture: records of { char name[32]; int32 kind (0 u8, 1 i32, 2 i64, 3 f64); int64 count; data }.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aces_b200.h"

#define MAXARR 64
typedef struct { char name[32]; int32_t kind; int64_t count; void* data; } Arr;
static Arr arrs[MAXARR];
static int n_arrs = 0, n_fail = 0;

static const Arr* get(const char* name) {
    for (int i = 0; i < n_arrs; ++i)
        if (!strcmp(arrs[i].name, name)) return &arrs[i];
    fprintf(stderr, "fixture has no array %s\n", name);
    exit(2);
}
#define U8(n) ((const uint8_t*)get(n)->data)
#define I32(n) ((const int32_t*)get(n)->data)
#define I64(n) ((const int64_t*)get(n)->data)
#define F64(n) ((const double*)get(n)->data)
#define CNT(n) (get(n)->count)

static void load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    static const size_t sz[4] = {1, 4, 8, 8};
    while (n_arrs < MAXARR && fread(arrs[n_arrs].name, 1, 32, f) == 32) {
        Arr* a = &arrs[n_arrs];
        if (fread(&a->kind, 4, 1, f) != 1 || fread(&a->count, 8, 1, f) != 1) exit(2);
        a->data = malloc(sz[a->kind] * (size_t)(a->count > 0 ? a->count : 1));   /* pageable, like a Julia Array */
        if (a->count > 0 && fread(a->data, sz[a->kind], (size_t)a->count, f) != (size_t)a->count) exit(2);
        ++n_arrs;
    }
    fclose(f);
}

static void report(const char* what, int ok, double err) {
    printf("%s %-44s %.3e\n", ok ? "PASS" : "FAIL", what, err);
    if (!ok) ++n_fail;
}
static double rel_err(const double* a, const double* b, int64_t n) {
    double num = 0, den = 0;
    for (int64_t i = 0; i < n; ++i) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return sqrt(num) / (den > 0 ? sqrt(den) : 1.0);
}

/* entry points, resolved with dlsym (the Julia shim names them the same way: (:name, LIB[])) */
#define FN(name) static __typeof__(&name) p_##name
FN(aces_version); FN(aces_init); FN(aces_destroy); FN(aces_last_error); FN(aces_host_alloc); FN(aces_host_free);
FN(aces_tables_create); FN(aces_tables_destroy); FN(aces_parity_counts); FN(aces_parity_counts_rowmajor);
FN(aces_estimate_experiment); FN(aces_estimate_gate_eigenvalues); FN(aces_design_create); FN(aces_design_destroy);
FN(aces_design_covariance_nnz); FN(aces_design_covariance_pattern); FN(aces_calc_covariance); FN(aces_design_inv_factor);
FN(aces_design_fit); FN(aces_design_ls_covariance); FN(aces_design_precision_blocks); FN(aces_project_simplex);
FN(aces_project_nonnegative_batched); FN(aces_last_rank_deficiency); FN(aces_synchronize);
FN(aces_design_set_tuple_shard); FN(aces_design_fit_begin); FN(aces_design_fit_buffers); FN(aces_design_fit_solve);
FN(aces_design_fit_refine_begin); FN(aces_design_fit_refine_end); FN(aces_design_fit_end);
FN(aces_symmetric_eigenvalues); FN(aces_design_merit); FN(aces_design_merit_grad); FN(aces_design_project_full);
FN(aces_design_fit_dist_begin); FN(aces_design_fit_dist_factor); FN(aces_design_fit_dist_update);
FN(aces_design_fit_dist_buffers); FN(aces_design_fit_dist_solve);
FN(aces_exchange_bytes); FN(aces_peer_alloc); FN(aces_peer_open); FN(aces_peer_close); FN(aces_peer_free);
FN(aces_exchange_create); FN(aces_exchange_destroy); FN(aces_exchange_push); FN(aces_exchange_pull);
FN(aces_exchange_attach); FN(aces_exchange_set_timeout); FN(aces_exchange_status);
#define LOAD(name) do { *(void**)(&p_##name) = dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); ++missing; } } while (0)

#define CHECK(call) do { int32_t rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_aces_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: abi_host <lib> --symbols | <fixture>\n"); return 2; }
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int missing = 0;
    LOAD(aces_version); LOAD(aces_init); LOAD(aces_destroy); LOAD(aces_last_error); LOAD(aces_host_alloc); LOAD(aces_host_free);
    LOAD(aces_tables_create); LOAD(aces_tables_destroy); LOAD(aces_parity_counts); LOAD(aces_parity_counts_rowmajor);
    LOAD(aces_estimate_experiment); LOAD(aces_estimate_gate_eigenvalues); LOAD(aces_design_create); LOAD(aces_design_destroy);
    LOAD(aces_design_covariance_nnz); LOAD(aces_design_covariance_pattern); LOAD(aces_calc_covariance); LOAD(aces_design_inv_factor);
    LOAD(aces_design_fit); LOAD(aces_design_ls_covariance); LOAD(aces_design_precision_blocks); LOAD(aces_project_simplex);
    LOAD(aces_project_nonnegative_batched); LOAD(aces_last_rank_deficiency); LOAD(aces_synchronize);
    LOAD(aces_design_set_tuple_shard); LOAD(aces_design_fit_begin); LOAD(aces_design_fit_buffers); LOAD(aces_design_fit_solve);
    LOAD(aces_design_fit_refine_begin); LOAD(aces_design_fit_refine_end); LOAD(aces_design_fit_end);
    LOAD(aces_symmetric_eigenvalues); LOAD(aces_design_merit); LOAD(aces_design_merit_grad); LOAD(aces_design_project_full);
    LOAD(aces_design_fit_dist_begin); LOAD(aces_design_fit_dist_factor); LOAD(aces_design_fit_dist_update);
    LOAD(aces_design_fit_dist_buffers); LOAD(aces_design_fit_dist_solve);
    LOAD(aces_exchange_bytes); LOAD(aces_peer_alloc); LOAD(aces_peer_open); LOAD(aces_peer_close); LOAD(aces_peer_free);
    LOAD(aces_exchange_create); LOAD(aces_exchange_destroy); LOAD(aces_exchange_push); LOAD(aces_exchange_pull);
    LOAD(aces_exchange_attach); LOAD(aces_exchange_set_timeout); LOAD(aces_exchange_status);
    if (missing) return 1;
    if (!strcmp(argv[2], "--symbols")) {
        printf("ALL SYMBOLS RESOLVED (ABI version %d)\n", (int)p_aces_version());
        return 0;
    }
    load(argv[2]);
    aces_ctx* ctx = NULL;
    if (p_aces_init(0, &ctx) != 0) { fprintf(stderr, "aces_init: %s\n", p_aces_last_error(NULL)); return 1; }

    /* ---- K1: simulate_stim_estimate's estimation half ------------------------------------------------ */
    const int32_t n_qubits = I32("n_qubits")[0], n_exp = I32("n_exp")[0], K = I32("K")[0];
    const int B = (n_qubits + 7) / 8;
    aces_tables* tables = NULL;
    CHECK(p_aces_tables_create(ctx, n_qubits, n_exp, I64("exp_ptr"), I64("pauli_ptr"), I32("pauli_bits"), &tables));
    const int64_t* exp_ptr = I64("exp_ptr");
    const int64_t* rows = I64("rows");
    const int64_t* rec_off = I64("rec_off");
    const int64_t* prefix = I64("prefix");     /* [n_exp][K] */
    const int64_t* want_odd = I64("odd");
    {   /* (a) one call per experiment, row-major records staged through pinned memory, as the shim does */
        int64_t off = 0, bad = 0;
        for (int32_t e = 0; e < n_exp; ++e) {
            const int64_t E = exp_ptr[e + 1] - exp_ptr[e], nb = rows[e] * B;
            void* pinned = NULL;
            CHECK(p_aces_host_alloc(ctx, nb, &pinned));
            memcpy(pinned, U8("records") + rec_off[e], (size_t)nb);
            int64_t* odd = (int64_t*)calloc((size_t)(E * K), 8);
            const uint8_t* ptrs[1] = {(const uint8_t*)pinned};
            const int64_t pitch = B;
            CHECK(p_aces_parity_counts_rowmajor(ctx, tables, 1, &e, ptrs, &rows[e], &pitch, K, prefix + (int64_t)e * K, odd, 0));
            for (int64_t i = 0; i < E * K; ++i) bad += odd[i] != want_odd[off + i];
            off += E * K;
            free(odd);
            CHECK(p_aces_host_free(ctx, pinned));
        }
        report("parity counts, row-major, per experiment", bad == 0, (double)bad);
    }
    {   /* (b) the same records in two accumulated batches (src/simulate.jl:174-186 without the vcat) */
        const int32_t e = 0;
        const int64_t E = exp_ptr[1] - exp_ptr[0], cut = rows[0] / 2 / 256 * 256 + 3;
        int64_t* odd = (int64_t*)calloc((size_t)(E * K), 8);
        int64_t bad = 0;
        for (int b = 0; b < 2; ++b) {
            const int64_t r0 = b ? cut : 0, nr = b ? rows[0] - cut : cut, pitch = B;
            int64_t pre[16];
            for (int k = 0; k < K; ++k) { int64_t s = prefix[k] - r0; pre[k] = s < 0 ? 0 : (s > nr ? nr : s); }
            const uint8_t* ptrs[1] = {U8("records") + rec_off[0] + r0 * B};   /* pageable, unaligned start */
            CHECK(p_aces_parity_counts_rowmajor(ctx, tables, 1, &e, ptrs, &nr, &pitch, K, pre, odd, b));
        }
        for (int64_t i = 0; i < E * K; ++i) bad += odd[i] != want_odd[i];
        report("parity counts, two accumulated batches", bad == 0, (double)bad);
        free(odd);
    }
    {   /* (c) all experiments in one launch from column-major matrices (Matrix{UInt8}, pageable) */
        const uint8_t** ptrs = (const uint8_t**)malloc(sizeof(void*) * (size_t)n_exp);
        int32_t* ids = (int32_t*)malloc(4 * (size_t)n_exp);
        uint8_t** cm = (uint8_t**)malloc(sizeof(void*) * (size_t)n_exp);
        for (int32_t e = 0; e < n_exp; ++e) {
            ids[e] = e;
            cm[e] = (uint8_t*)malloc((size_t)(rows[e] * B) + 1);
            const uint8_t* r = U8("records") + rec_off[e];
            for (int64_t s = 0; s < rows[e]; ++s)
                for (int c = 0; c < B; ++c) cm[e][c * rows[e] + s] = r[s * B + c];
            ptrs[e] = cm[e];
        }
        const int64_t total = CNT("odd");
        int64_t* odd = (int64_t*)calloc((size_t)total, 8);
        CHECK(p_aces_parity_counts(ctx, tables, n_exp, ids, ptrs, rows, rows, K, prefix, odd, 0));
        int64_t bad = 0;
        for (int64_t i = 0; i < total; ++i) bad += odd[i] != want_odd[i];
        report("parity counts, column-major, one launch", bad == 0, (double)bad);
        /* S1: estimate_experiment on the first matrix with 1-based qubit indices */
        const int64_t E0 = CNT("s1_est");
        double* est = (double*)malloc(8 * (size_t)E0);
        CHECK(p_aces_estimate_experiment(ctx, cm[0], rows[0], rows[0], B, (int32_t)E0, I64("s1_ptr"), I32("s1_qubits"),
                                         U8("s1_signs"), prefix[K - 1], est));
        int64_t badf = 0;
        for (int64_t i = 0; i < E0; ++i) badf += est[i] != F64("s1_est")[i];
        report("estimate_experiment (S1), bit-identical", badf == 0, (double)badf);
        for (int32_t e = 0; e < n_exp; ++e) free(cm[e]);
        free(cm); free(ids); free(ptrs); free(odd); free(est);
    }

    /* ---- fit: estimate_gate_noise's numeric core on the device-resident design ------------------------ */
    const int32_t T = I32("T")[0];
    const int64_t N = I64("N")[0], M = I64("M")[0];
    aces_design* dd = NULL;
    CHECK(p_aces_design_create(ctx, T, I64("mapping_lengths"), N, I32("A_colptr"), I32("A_rowval"), I32("A_nzval"), 1,
                               I64("experiment_numbers"), F64("shot_weights"), F64("tuple_times"), I64("diag_counts"),
                               I64("cov_ptr"), I64("cov_p"), I64("cov_q"), I64("cov_counts"), &dd));
    {
        const int64_t nnz = p_aces_design_covariance_nnz(dd);
        double* nz = (double*)malloc(8 * (size_t)nnz);
        CHECK(p_aces_calc_covariance(ctx, dd, F64("lam"), F64("lam_cov"), 1e-10, 0, 1, nz));
        const double e = nnz == CNT("cov_log_nzval") ? rel_err(nz, F64("cov_log_nzval"), nnz) : 1.0;
        report("calc_covariance + calc_covariance_log", e <= 1e-14, e);
        free(nz);
    }
    double* g = (double*)malloc(8 * (size_t)N);
    const char* names[3] = {"ols", "wls", "gls"};
    for (int ls = 0; ls < 3; ++ls) {
        CHECK(p_aces_design_fit(ctx, dd, ls, F64("lam"), F64("lam_cov"), I64("kept_rows"), CNT("kept_rows"), 1, 1e-10, 0, g));
        char what[64];
        snprintf(what, sizeof what, "design_fit %s, clipped, 1-based kept rows", names[ls]);
        const double e = rel_err(g, F64(names[ls]), N);
        const int deficiency = p_aces_last_rank_deficiency(ctx);
        if (deficiency != 0) printf("     rank deficiency reported: %d\n", deficiency);
        report(what, e <= 1e-9 && deficiency == 0, e);
    }
    {   /* precision slices of the GLS fit that just ran, 1-based gate groups */
        const int32_t G = (int32_t)(CNT("gate_ptr") - 1);
        const int64_t nb = CNT("precision_blocks");
        double* blocks = (double*)malloc(8 * (size_t)nb);
        CHECK(p_aces_design_precision_blocks(ctx, dd, 2, G, I64("gate_ptr"), I32("gate_idx1"), 1, g, blocks));
        const double e = rel_err(blocks, F64("precision_blocks"), nb);
        report("design_precision_blocks (GLS)", e <= 1e-9, e);
        free(blocks);
        double* proj = (double*)malloc(8 * (size_t)N);
        double* up = (double*)malloc(8 * (size_t)(N + G));
        double* pp = (double*)malloc(8 * (size_t)(N + G));
        CHECK(p_aces_project_simplex(ctx, N, G, I64("gate_ptr"), I32("gate_idx1"), 1, F64("gls"), proj, up, pp));
        int64_t bad = 0;
        for (int64_t i = 0; i < N; ++i) bad += proj[i] != F64("gls_simple_proj")[i];
        report("project_simplex, bit-identical", bad == 0, (double)bad);
        free(proj); free(up); free(pp);
    }
    {   /* S2 with 1-based CSC arrays (SparseMatrixCSC{Float64,Int}) */
        double* g2 = (double*)malloc(8 * (size_t)N);
        CHECK(p_aces_estimate_gate_eigenvalues(ctx, CNT("kept_rows"), N, I64("Ac_colptr"), I64("Ac_rowval"), F64("Ac_nzval"), NULL, NULL,
                                               NULL, 1, F64("lam_kept"), 0, g2));
        const double e = rel_err(g2, F64("ols"), N);
        report("estimate_gate_eigenvalues (S2, OLS form)", e <= 1e-9, e);
        free(g2);
    }
    {   /* S4: trace moments of the GLS covariance */
        double tr = 0, trsq = 0;
        CHECK(p_aces_design_ls_covariance(ctx, dd, 2, F64("gate_eigenvalues"), F64("true_cov_log_nzval"), NULL, &tr, &trsq));
        const double e = fabs(tr - F64("tr_gls")[0]) / F64("tr_gls")[0];
        report("design_ls_covariance (GLS trace)", e <= 1e-9, e);
    }
    {   /* f2: eigvals(::Symmetric) -- the 1-D Laplacian has the spectrum 2 - 2 cos(k pi / (n + 1)); only the upper
           triangle is read (Julia's Symmetric default), the lower one holds garbage here */
        const int64_t n = 200;
        double* a = (double*)malloc(sizeof(double) * (size_t)(n * n));
        double* w = (double*)malloc(sizeof(double) * (size_t)n);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = 0; i < n; ++i) a[i + j * n] = i > j ? 1e30 : (i == j ? 2.0 : (i + 1 == j ? -1.0 : 0.0));
        CHECK(p_aces_symmetric_eigenvalues(ctx, n, a, n, w));
        double worst = 0;
        for (int64_t k = 0; k < n; ++k) {
            const double want = 2.0 - 2.0 * cos((double)(k + 1) * 3.14159265358979323846 / (double)(n + 1));
            if (fabs(w[k] - want) > worst) worst = fabs(w[k] - want);
        }
        report("symmetric_eigenvalues (1-D Laplacian, n = 200)", worst <= 4e-13, worst);
        free(a); free(w);
    }
    free(g);
    CHECK(p_aces_design_destroy(ctx, dd));
    CHECK(p_aces_tables_destroy(ctx, tables));
    CHECK(p_aces_destroy(ctx));
    printf(n_fail ? "%d FAILED\n" : "ALL PASS\n", n_fail);
    return n_fail ? 1 : 0;
}
