// aces_common.h — internal definitions shared by the translation units of libaces_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/aces_b200.h"

#define ACES_VERSION 1000

// Growable device / pinned-host buffers owned by the context.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
};

struct aces_exchange;

struct aces_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream in use (own_stream or caller's)
    cudaEvent_t ev = nullptr;
    cudaStream_t aux_stream = nullptr;            // low-priority side stream (look-ahead updates)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t aux2_stream = nullptr;           // medium priority: off-critical-path work of a Cholesky step
    cudaEvent_t ev_crit = nullptr, ev_aux2 = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;  // bracket the dominant kernel when timing is on
    int timing = 0;
    int timed = 0;
    int64_t launches = 0;
    aces_exchange* xch = nullptr;  // attached count exchange: aces_parity_counts_async pushes from its prefix kernel
    int rank_deficiency = 0;  // columns eliminated by the last fit (aces_last_rank_deficiency)
    std::string err;

    // K1 workspaces
    DevBuf d_arena;    // pitched copies of sample matrices that cannot be read in place
    DevBuf d_raw;      // row-major records as they arrive from the host (ingest path), before the transposition
    DevBuf d_tdesc;    // descriptor table of the batched transposition
    DevBuf d_streams;  // stream table + item prefix
    DevBuf d_seg;      // per-segment odd counts (uint64)
    DevBuf d_odd;      // cumulative odd counts when the caller's output is on the host
    DevBuf h_stage;    // pinned staging for the small per-call tables (two slots, used alternately so
    DevBuf h_stage2;   // that the host can prepare call i+1 while call i's copy is still queued)
    DevBuf h_pack;     // pinned staging of small host sample matrices (one upload for all of them)
    cudaEvent_t ev2 = nullptr;
    int stage_slot = 0;
    std::vector<uint8_t> tbl_scratch, tbl_uploaded;  // K1 tables: last upload (skipped when unchanged)
    const void* tbl_dev = nullptr;
    bool seg_clean = false;
    // fit workspaces are added by fit.cu through this generic pool
    DevBuf d_work[8];
    DevBuf d_flags;    // inter-CTA flags of the persistent triangular solves
    DevBuf d_dag_trace;
    DevBuf d_eig, d_eigA;  // workspace / matrix copy of the symmetric eigenvalue solver (eigen.cu)
    DevBuf d_dag_tasks, d_dag_state;  // task list (cached per matrix size) and counters of the persistent Cholesky
    int dag_nb = 0, dag_grid = 0, dag_ntasks = 0, dag_nlist = 0;
};

int aces_fail(aces_ctx* ctx, int code, const char* fmt, ...);
int aces_reserve_dev(aces_ctx* ctx, DevBuf& b, size_t bytes);
int aces_reserve_host(aces_ctx* ctx, DevBuf& b, size_t bytes);
bool aces_is_device_ptr(const void* p);

#define ACES_CUDA(ctx, call)                                                              \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess)                                                            \
            return aces_fail((ctx), ACES_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, \
                             __LINE__, cudaGetErrorString(_e));                           \
    } while (0)

#define ACES_CHECK(ctx, cond, ...)                                    \
    do {                                                              \
        if (!(cond)) return aces_fail((ctx), ACES_E_INVALID, __VA_ARGS__); \
    } while (0)

#define ACES_TRY(expr)               \
    do {                             \
        int _rc = (expr);            \
        if (_rc != ACES_OK) return _rc; \
    } while (0)
