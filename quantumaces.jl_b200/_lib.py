"""ctypes binding of libaces_b200.so (the C ABI declared in include/aces_b200.h).

This is the Python twin of the Julia `ccall` shim in julia/ACESB200.jl: the same entry
points, the same argument meaning.  There is no CPU fallback: if the shared library is
missing, or no sm_100 GPU is present when a context is created, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACES_B200_LIB") or os.path.join(_HERE, "libaces_b200.so")

ACES_OK = 0
ACES_E_INVALID = -1
ACES_E_CUDA = -2
ACES_E_NOMEM = -3
ACES_E_NOT_PD = -4
ACES_E_UNSUPPORTED = -5


class AcesError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libaces_b200 error {code}: {message}")
        self.code = code


class NotPositiveDefinite(AcesError):
    """A covariance block or Gram matrix is not positive definite (ACES_E_NOT_PD)."""


_lib: Optional[C.CDLL] = None

_i32, _i64, _vp = C.c_int32, C.c_int64, C.c_void_p
_p = C.POINTER

# name -> (restype, argtypes); kept in sync with include/aces_b200.h (tests check the symbols)
SIGNATURES = {
    "aces_version": (_i32, []),
    "aces_init": (_i32, [_i32, _p(_vp)]),
    "aces_destroy": (_i32, [_vp]),
    "aces_last_error": (C.c_char_p, [_vp]),
    "aces_set_stream": (_i32, [_vp, _vp]),
    "aces_synchronize": (_i32, [_vp]),
    "aces_host_alloc": (_i32, [_vp, _i64, _p(_vp)]),
    "aces_host_free": (_i32, [_vp, _vp]),
    "aces_kernel_launches": (_i64, [_vp]),
    "aces_last_rank_deficiency": (_i32, [_vp]),
    "aces_set_timing": (_i32, [_vp, _i32]),
    "aces_last_kernel_ms": (_i32, [_vp, _p(C.c_float)]),
    "aces_tables_create": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _p(_vp)]),
    "aces_tables_destroy": (_i32, [_vp, _vp]),
    "aces_tables_group_columns": (_i64, [_vp, _i32]),
    "aces_parity_counts": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "aces_parity_counts_async": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "aces_parity_counts_rowmajor": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "aces_estimate_experiment": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "aces_estimate_gate_eigenvalues": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp]),
    "aces_gram_inverse": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp]),
    "aces_design_create": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _p(_vp)]),
    "aces_design_destroy": (_i32, [_vp, _vp]),
    "aces_design_covariance_nnz": (_i64, [_vp]),
    "aces_design_covariance_pattern": (_i32, [_vp, _vp, _vp]),
    "aces_calc_covariance": (_i32, [_vp, _vp, _vp, _vp, C.c_double, _i32, _i32, _vp]),
    "aces_design_inv_factor": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, C.c_double, _vp, _vp]),
    "aces_design_fallback_count": (_i32, [_vp]),
    "aces_design_gram_flops": (C.c_double, [_vp]),
    "aces_design_phase_ms": (_i32, [_vp, _i32, C.c_char_p, _i32, _p(C.c_float)]),
    "aces_design_fit": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, C.c_double, _i32, _vp]),
    "aces_design_ls_covariance": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "aces_design_merit_grad": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "aces_symmetric_eigenvalues": (_i32, [_vp, _i64, _vp, _i64, _vp]),
    "aces_design_merit": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp]),
    "aces_design_set_tuple_shard": (_i32, [_vp, _vp, _vp]),
    "aces_design_fit_begin": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, C.c_double]),
    "aces_design_fit_buffers": (_i32, [_vp, _p(_vp), _p(_i64), _p(_vp), _p(_i64)]),
    "aces_design_fit_solve": (_i32, [_vp, _vp]),
    "aces_design_fit_refine_begin": (_i32, [_vp, _vp]),
    "aces_design_fit_refine_end": (_i32, [_vp, _vp, _p(C.c_double), _p(C.c_double)]),
    "aces_design_fit_end": (_i32, [_vp, _vp, _i32, _vp]),
    "aces_design_fit_dist_begin": (_i32, [_vp, _vp]),
    "aces_design_fit_dist_factor": (_i32, [_vp, _vp, _i64, _i64]),
    "aces_design_fit_dist_update": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "aces_design_fit_dist_buffers": (_i32, [_vp, _p(_vp), _p(_i64), _p(_vp)]),
    "aces_design_fit_dist_solve": (_i32, [_vp, _vp]),
    "aces_exchange_bytes": (_i64, [_i32, _i64]),
    "aces_peer_alloc": (_i32, [_vp, _i64, _p(_vp), _vp]),
    "aces_peer_open": (_i32, [_vp, _vp, _p(_vp)]),
    "aces_peer_close": (_i32, [_vp, _vp]),
    "aces_peer_free": (_i32, [_vp, _vp]),
    "aces_exchange_create": (_i32, [_vp, _i32, _i32, _i64, _vp, _p(_vp)]),
    "aces_exchange_destroy": (_i32, [_vp, _vp]),
    "aces_exchange_push": (_i32, [_vp, _vp, _vp, _vp]),
    "aces_exchange_pull": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "aces_exchange_attach": (_i32, [_vp, _vp]),
    "aces_exchange_set_timeout": (_i32, [_vp, _vp, _i64]),
    "aces_exchange_status": (_i32, [_vp, _vp]),
    "aces_design_project_full": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, C.c_double, _vp, _p(_i32)]),
    "aces_design_precision_blocks": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "aces_potrf_schedule": (_i32, [_i32, _i32, _vp, _i64, _p(_i64)]),
    "aces_potrf_trace": (_i32, [_vp, _vp, _i64, _p(_i64)]),
    "aces_project_simplex": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "aces_project_nonnegative_batched": (_i32, [_vp, _i32, _vp, _vp, _vp, C.c_double, _vp, _p(_i32)]),
}


def load() -> C.CDLL:
    """Loads libaces_b200.so; raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "this package has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a) -> Optional[int]:
    """Address of a numpy array / int address / None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


class Context:
    """One libaces_b200 context (= one GPU).  Mirrors aces_init / aces_destroy."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.aces_init(int(device), C.byref(h))
        if rc != ACES_OK:
            msg = self.lib.aces_last_error(None)
            raise AcesError(rc, msg.decode() if msg else "aces_init failed")
        self.handle = h
        self.device = int(device)

    def check(self, rc: int) -> None:
        if rc == ACES_OK:
            return
        msg = self.lib.aces_last_error(self.handle)
        text = msg.decode() if msg else ""
        if rc == ACES_E_NOT_PD:
            raise NotPositiveDefinite(rc, text)
        raise AcesError(rc, text)

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.aces_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self.check(self.lib.aces_set_stream(self.handle, cuda_stream))

    def synchronize(self) -> None:
        self.check(self.lib.aces_synchronize(self.handle))

    def kernel_launches(self) -> int:
        return int(self.lib.aces_kernel_launches(self.handle))

    def last_rank_deficiency(self) -> int:
        """Columns the last fit eliminated as numerically dependent (0 for a full-rank fit)."""
        return int(self.lib.aces_last_rank_deficiency(self.handle))

    def set_timing(self, on: bool) -> None:
        self.check(self.lib.aces_set_timing(self.handle, 1 if on else 0))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self.check(self.lib.aces_last_kernel_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a uint8 numpy array (freed with host_free)."""
        p = _vp()
        self.check(self.lib.aces_host_alloc(self.handle, int(nbytes), C.byref(p)))
        buf = (C.c_uint8 * max(int(nbytes), 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=int(nbytes))
        arr.flags.writeable = True
        return arr

    def host_free(self, arr: np.ndarray) -> None:
        self.check(self.lib.aces_host_free(self.handle, arr.ctypes.data))


class PauliTables:
    """Device-resident Pauli tables of a design (aces_tables_create)."""

    def __init__(self, ctx: Context, n_qubits: int, exp_ptr, pauli_ptr, pauli_bits):
        self.ctx = ctx
        self.n_qubits = int(n_qubits)
        self.exp_ptr = np.ascontiguousarray(exp_ptr, dtype=np.int64)
        self.pauli_ptr = np.ascontiguousarray(pauli_ptr, dtype=np.int64)
        self.pauli_bits = np.ascontiguousarray(pauli_bits, dtype=np.int32)
        self.n_experiments = len(self.exp_ptr) - 1
        if self.pauli_bits.size == 0:
            self.pauli_bits = np.zeros(1, dtype=np.int32)
        h = _vp()
        ctx.check(
            ctx.lib.aces_tables_create(
                ctx.handle, self.n_qubits, self.n_experiments, _ptr(self.exp_ptr),
                _ptr(self.pauli_ptr), _ptr(self.pauli_bits), C.byref(h),
            )
        )
        self.handle = h

    def experiment_size(self, e: int) -> int:
        return int(self.exp_ptr[e + 1] - self.exp_ptr[e])

    def columns_read(self, e: int) -> int:
        """Byte columns the kernel reads per shot of experiment e (aces_tables_group_columns)."""
        return int(self.ctx.lib.aces_tables_group_columns(self.handle, int(e)))

    def close(self) -> None:
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.aces_tables_destroy(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parity_counts(
    ctx: Context,
    tables: PauliTables,
    experiment_ids: Sequence[int],
    matrices: Sequence,
    rows: Sequence[int],
    ld: Sequence[int],
    prefix_shots,
    out=None,
    accumulate: bool = False,
    asynchronous: bool = False,
    rowmajor: bool = False,
):
    """aces_parity_counts / aces_parity_counts_async / aces_parity_counts_rowmajor (`rowmajor`: the
    matrices are C-ordered [rows x B] records and `ld` is the byte pitch between shots).

    `matrices` are numpy uint8 arrays (column-major, i.e. Fortran order [rows x cols]) or raw
    integer addresses (host or device).  `out` is None (a host int64 array is returned), a
    numpy int64 array, or an integer device address.
    """
    n = len(experiment_ids)
    eid = np.ascontiguousarray(experiment_ids, dtype=np.int32)
    rows_a = np.ascontiguousarray(rows, dtype=np.int64)
    ld_a = np.ascontiguousarray(ld, dtype=np.int64)
    pre = np.ascontiguousarray(prefix_shots, dtype=np.int64).reshape(n, -1)
    K = pre.shape[1]
    ptrs = (C.c_void_p * max(n, 1))()
    for i, m in enumerate(matrices):
        ptrs[i] = _ptr(m)
    total = int(sum(tables.experiment_size(int(e)) for e in eid) * K)
    ret = None
    if out is None:
        ret = np.zeros(total, dtype=np.int64)
        out_ptr = _ptr(ret)
    else:
        out_ptr = _ptr(out)
        ret = out
    fn = ctx.lib.aces_parity_counts_async if asynchronous else ctx.lib.aces_parity_counts
    if rowmajor:
        assert not asynchronous
        fn = ctx.lib.aces_parity_counts_rowmajor
    ctx.check(
        fn(
            ctx.handle, tables.handle, n, _ptr(eid), C.cast(ptrs, C.c_void_p), _ptr(rows_a),
            _ptr(ld_a), K, _ptr(pre), out_ptr, 1 if accumulate else 0,
        )
    )
    return ret


class PreparedBatch:
    """The argument arrays of one aces_parity_counts[_async] call, marshalled once: repeated launches
    over the same buffers (the repetition loop re-fills the same sample matrices) then cost one
    foreign call each."""

    def __init__(self, ctx: Context, tables: PauliTables, experiment_ids, matrices, rows, ld, prefix_shots, out,
                 accumulate: bool = False, asynchronous: bool = False, rowmajor: bool = False):
        n = len(experiment_ids)
        self.ctx, self.tables, self.n = ctx, tables, n
        self.eid = np.ascontiguousarray(experiment_ids, dtype=np.int32)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.ld = np.ascontiguousarray(ld, dtype=np.int64)
        self.pre = np.ascontiguousarray(prefix_shots, dtype=np.int64).reshape(n, -1)
        self.K = self.pre.shape[1]
        self.ptrs = (C.c_void_p * max(n, 1))()
        self.keep = list(matrices)  # keeps numpy-backed matrices alive
        for i, m in enumerate(matrices):
            self.ptrs[i] = _ptr(m)
        self.out = out
        self.fn = ctx.lib.aces_parity_counts_async if asynchronous else ctx.lib.aces_parity_counts
        if rowmajor:
            assert not asynchronous
            self.fn = ctx.lib.aces_parity_counts_rowmajor
        self.args = (ctx.handle, tables.handle, n, _ptr(self.eid), C.cast(self.ptrs, C.c_void_p), _ptr(self.rows),
                     _ptr(self.ld), self.K, _ptr(self.pre), _ptr(out), 1 if accumulate else 0)

    def launch(self):
        rc = self.fn(*self.args)
        if rc != ACES_OK:
            self.ctx.check(rc)
        return self.out


def estimate_experiment_raw(ctx: Context, counts: np.ndarray, meas_ptr, meas_indices, signs, shots_value: int):
    """aces_estimate_experiment on a Fortran-ordered uint8 matrix."""
    assert counts.dtype == np.uint8 and counts.ndim == 2
    if not counts.flags.f_contiguous:
        counts = np.asfortranarray(counts)
    rows, ncols = counts.shape
    meas_ptr = np.ascontiguousarray(meas_ptr, dtype=np.int64)
    meas_indices = np.ascontiguousarray(meas_indices, dtype=np.int32)
    signs = np.ascontiguousarray(signs, dtype=np.uint8)
    E = len(meas_ptr) - 1
    est = np.zeros(E, dtype=np.float64)
    if meas_indices.size == 0:
        meas_indices = np.zeros(1, dtype=np.int32)
    ctx.check(
        ctx.lib.aces_estimate_experiment(
            ctx.handle, _ptr(counts), rows, max(rows, 1), ncols, E, _ptr(meas_ptr),
            _ptr(meas_indices), _ptr(signs), int(shots_value), _ptr(est),
        )
    )
    return est
