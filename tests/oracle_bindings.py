"""ctypes binding of oracle/libaces_oracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libaces_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
        L = C.CDLL(path)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.oracle_pauli_estimate.restype = i64
        L.oracle_pauli_estimate.argtypes = [vp, i64, vp, i32, i64]
        L.oracle_estimate_experiment.restype = None
        L.oracle_estimate_experiment.argtypes = [vp, i64, i32, vp, vp, vp, i64, vp, vp]
        L.oracle_simulate_estimate.restype = None
        L.oracle_simulate_estimate.argtypes = [vp, i64, i32, vp, vp, vp, i32, vp, vp, vp]
        L.oracle_batch_shots.restype = i64
        L.oracle_batch_shots.argtypes = [i64, i64, i64, vp, i64]
        L.oracle_num_threads.restype = i32
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def _csr(lists):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    flat = np.zeros(max(int(ptr[-1]), 1), dtype=np.int32)
    for i, l in enumerate(lists):
        flat[ptr[i]:ptr[i + 1]] = l
    return ptr, flat


def pauli_estimate(counts, meas_indices, shots_value):
    counts = np.asfortranarray(counts)
    mi = np.ascontiguousarray(meas_indices, dtype=np.int32)
    if mi.size == 0:
        mi = np.zeros(1, dtype=np.int32)
    return int(lib().oracle_pauli_estimate(counts.ctypes.data, max(counts.shape[0], 1), mi.ctypes.data,
                                           len(meas_indices), int(shots_value)))


def estimate_experiment(counts, experiment_meas_indices, experiment_pauli_signs, shots_value):
    counts = np.asfortranarray(counts)
    E = len(experiment_meas_indices)
    ptr, flat = _csr(experiment_meas_indices)
    sg = np.ascontiguousarray(experiment_pauli_signs, dtype=np.uint8)
    est = np.zeros(max(E, 1), dtype=np.float64)
    ps = np.zeros(max(E, 1), dtype=np.int64)
    lib().oracle_estimate_experiment(counts.ctypes.data, max(counts.shape[0], 1), E, ptr.ctypes.data,
                                     flat.ctypes.data, sg.ctypes.data, int(shots_value),
                                     est.ctypes.data, ps.ctypes.data)
    return est[:E], ps[:E]


def simulate_estimate(counts, experiment_meas_indices, experiment_pauli_signs, prefix_shots):
    """Returns (est [K,E], odd [K,E]) for nested prefixes of one sample matrix."""
    counts = np.asfortranarray(counts)
    E = len(experiment_meas_indices)
    K = len(prefix_shots)
    ptr, flat = _csr(experiment_meas_indices)
    sg = np.ascontiguousarray(experiment_pauli_signs, dtype=np.uint8)
    pre = np.ascontiguousarray(prefix_shots, dtype=np.int64)
    est = np.zeros((K, max(E, 1)), dtype=np.float64)
    odd = np.zeros((K, max(E, 1)), dtype=np.int64)
    lib().oracle_simulate_estimate(counts.ctypes.data, max(counts.shape[0], 1), E, ptr.ctypes.data,
                                   flat.ctypes.data, sg.ctypes.data, K, pre.ctypes.data,
                                   est.ctypes.data, odd.ctypes.data)
    if E == 0:
        return est[:, :0], odd[:, :0]
    return est, odd


def batch_shots(shots, measurements, max_samples):
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.int64)
    n = lib().oracle_batch_shots(int(shots), int(measurements), int(max_samples), out.ctypes.data, cap)
    if n < 0:
        raise AssertionError("batch_shots assertion failed")
    return out[:n].copy()


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def table_supports(est, e):
    """1-based qubit lists of every Pauli of table experiment e of a DesignEstimator."""
    t = est.tables
    sup = []
    for p in range(int(t.exp_ptr[e]), int(t.exp_ptr[e + 1])):
        sup.append((t.pauli_bits[t.pauli_ptr[p]:t.pauli_ptr[p + 1]] + 1).tolist())
    return sup


def oracle_odd_counts(est, mats, prefix, experiment_ids=None):
    """Oracle twin of DesignEstimator.odd_counts: same flat [item][k][p] layout."""
    if experiment_ids is None:
        experiment_ids = range(len(mats))
    out = []
    for i, e in enumerate(experiment_ids):
        sup = table_supports(est, int(e))
        _, odd = simulate_estimate(mats[i], sup, np.zeros(len(sup), dtype=np.uint8), prefix[i])
        out.append(odd.reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)
