"""Shot-weight optimisation on the device-resident design.

Host-side mirror of src/optimise_weights.jl: the gradient of the figure of merit with respect to
the logarithms of the shot weights is one call of aces_design_merit_grad (csrc/merit_grad.cuh: Gram
matrix, Cholesky factor, inverse and two to six N^3 DMMA GEMMs per step, nothing larger than N x N),
the Nesterov gradient-descent loop around it is the reference's host logic restated line by line.
Names, argument meaning and defaults follow the reference (`OptimOptions`, src/kwargs.jl:59-117).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from ._lib import _ptr
from .merit import DeviceDesign, LS_TYPES

# get_orbit_indices_dict (src/noise.jl:66-124): 1-based Pauli indices of every orbit of a gate type
_ONE = [[1], [2], [3]]
_TWO = [[i] for i in range(1, 16)]
ORBIT_INDICES = {
    "I": _ONE, "IM": _ONE, "X": _ONE, "Y": _ONE, "Z": _ONE,
    "H": [[1, 2], [3]], "S": [[1, 3], [2]], "S_DAG": [[1, 3], [2]],
    "II": _TWO,
    "CX": [[1, 3], [2], [4], [5, 7], [6], [8, 12], [9, 15], [10, 14], [11, 13]],
    "CZ": [[1, 9], [2, 6], [3, 15], [4], [5, 13], [7, 11], [8], [10, 14], [12]],
    "SQRT_XX": [[1], [2], [3], [4, 7], [5, 6], [8, 11], [9, 10], [12], [13], [14], [15]],
    "SQRT_XY": [[1], [2, 9], [3, 8], [4, 15], [5, 14], [6], [7], [10], [11], [12], [13]],
    "SQRT_XZ": [[1], [2, 11], [3, 10], [4, 13], [5, 12], [6], [7], [8], [9], [14], [15]],
    "SQRT_YX": [[1, 6], [2], [3, 4], [5], [7], [8, 15], [9], [10, 13], [11], [12], [14]],
    "SQRT_YY": [[1, 14], [2, 13], [3], [4, 11], [5], [6], [7, 8], [9], [10], [12], [15]],
    "SQRT_YZ": [[1, 12], [2, 15], [3], [4, 9], [5], [6], [7, 10], [8], [11], [13], [14]],
    "SQRT_ZX": [[1, 7], [2], [3, 5], [4], [6], [8, 14], [9], [10, 12], [11], [13], [15]],
    "SQRT_ZY": [[1, 15], [2, 12], [3], [4], [5, 11], [6, 8], [7], [9], [10], [13], [14]],
    "SQRT_ZZ": [[1, 13], [2, 14], [3], [4], [5, 9], [6, 10], [7], [8], [11], [12], [15]],
}
ORBIT_INDICES["CNOT"] = ORBIT_INDICES["CX"]
for _p in ("XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"):
    ORBIT_INDICES[_p] = _TWO
    ORBIT_INDICES["SQRT_" + _p + "_DAG"] = ORBIT_INDICES["SQRT_" + _p]
_SPAM_OR_MID = ("PZ", "PX", "PY", "MZ", "MX", "MY", "M", "R")


def get_transform(fd, est_type: str) -> sp.csc_matrix:
    """get_transform(d, est_type) (src/merit.jl:455-470): the identity (:ordinary), the orbit averages
    of every gate (:marginal, src/noise.jl:689-709) or those of the gates estimable to relative
    precision (:relative, src/noise.jl:716-741; is_additive with strict = false, src/tableau.jl:736-743)."""
    N = int(fd.gate_ptr[-1]) if fd.gate_ptr is not None else int(fd.matrix.shape[1])
    if est_type == "ordinary":
        return sp.identity(N, format="csc", dtype=np.float64)
    if est_type not in ("marginal", "relative"):
        raise ValueError(f"The estimator type {est_type} must be either :ordinary, :marginal, or :relative.")
    if fd.gate_types is None or fd.gate_ptr is None:
        raise ValueError("the design carries no gate types: marginal / relative transforms need them")
    rows, cols, vals = [], [], []
    r = 0
    for g, ty in enumerate(fd.gate_types):
        idx = int(fd.gate_ptr[g])
        if est_type == "relative" and (ty in _SPAM_OR_MID or ty == "IM"):
            continue
        if ty in _SPAM_OR_MID:
            rows.append(r)
            cols.append(idx)
            vals.append(1.0)
            r += 1
            continue
        orbits = ORBIT_INDICES[ty]
        for k, orbit in enumerate(orbits):
            for o in orbit:
                rows.append(r + k)
                cols.append(idx + o - 1)
                vals.append(1.0 / len(orbit))
        r += len(orbits)
    return sp.csc_matrix((vals, (rows, cols)), shape=(r, N))


def project_simplex_host(probabilities: np.ndarray) -> np.ndarray:
    """project_simplex (src/utils.jl:36-50) for the T shot weights (host: T is the number of tuples)."""
    p = np.asarray(probabilities, dtype=np.float64)
    assert not np.any(np.isnan(p)), "The supplied vector contains NaN."
    srt = np.sort(p)[::-1]
    sums = (1.0 / np.arange(1, len(p) + 1)) * (1.0 - np.cumsum(srt))
    last = np.nonzero(srt + sums > 0)[0][-1]
    return np.maximum(p + sums[last], 0.0)


def get_shot_weights_factor(shot_weights, tuple_times, mapping_lengths) -> np.ndarray:
    """src/optimise_weights.jl:6-26 (the diagonal as a vector)."""
    times_factor = np.sum(shot_weights * tuple_times)
    return times_factor * np.repeat(1.0 / np.asarray(shot_weights), np.asarray(mapping_lengths, dtype=np.int64))


def get_shot_weights_factor_inv(shot_weights, tuple_times, mapping_lengths) -> np.ndarray:
    """src/optimise_weights.jl:33-49."""
    times_factor = np.sum(shot_weights * tuple_times)
    return (1.0 / times_factor) * np.repeat(np.asarray(shot_weights), np.asarray(mapping_lengths, dtype=np.int64))


def calc_merit_grad_log(dd: DeviceDesign, ls_type: str, shot_weights, covariance_log_unweighted, gate_transform_matrix
                        ) -> Tuple[np.ndarray, float]:
    """calc_gls_merit_grad_log / calc_wls_merit_grad_log / calc_ols_merit_grad_log
    (src/optimise_weights.jl:103-152, 327-391, 563-606) -> (merit_grad_log, merit).  The GLS method of
    the reference takes the block inverse of covariance_log_unweighted; here every estimator takes the
    matrix itself (the device factorises the blocks)."""
    fd = dd.fd
    w = np.ascontiguousarray(shot_weights, dtype=np.float64)
    tau = np.ascontiguousarray(fd.tuple_times, dtype=np.float64)
    val = dd.pattern_values(covariance_log_unweighted) if sp.issparse(covariance_log_unweighted) else \
        np.ascontiguousarray(covariance_log_unweighted, dtype=np.float64)
    Tm = sp.csc_matrix(gate_transform_matrix)
    tt = sp.csc_matrix(Tm.T @ Tm)
    tt.sort_indices()
    colptr = np.ascontiguousarray(tt.indptr, dtype=np.int64)
    rowval = np.ascontiguousarray(tt.indices, dtype=np.int64)
    nzval = np.ascontiguousarray(tt.data, dtype=np.float64)
    merit = C.c_double()
    grad = np.zeros(fd.tuple_number, dtype=np.float64)
    dd.ctx.check(dd.ctx.lib.aces_design_merit_grad(
        dd.ctx.handle, dd.handle, LS_TYPES[ls_type], _ptr(val), _ptr(w), _ptr(tau), int(Tm.shape[0]), _ptr(colptr),
        _ptr(rowval), _ptr(nzval), 0, C.byref(merit), _ptr(grad), None, None))
    return grad, float(merit.value)


def calc_gls_merit_grad_log(dd, shot_weights, covariance_log_unweighted, gate_transform_matrix):
    return calc_merit_grad_log(dd, "gls", shot_weights, covariance_log_unweighted, gate_transform_matrix)


def calc_wls_merit_grad_log(dd, shot_weights, covariance_log_unweighted, gate_transform_matrix):
    return calc_merit_grad_log(dd, "wls", shot_weights, covariance_log_unweighted, gate_transform_matrix)


def calc_ols_merit_grad_log(dd, shot_weights, covariance_log_unweighted, gate_transform_matrix):
    return calc_merit_grad_log(dd, "ols", shot_weights, covariance_log_unweighted, gate_transform_matrix)


@dataclass
class OptimOptions:
    """The gradient-descent fields of OptimOptions (src/kwargs.jl:59-117) with the reference's defaults."""
    ls_type: str = "gls"
    est_type: str = "prod"
    est_weight: float = 0.5
    learning_rate: Optional[float] = None          # (ls_type == :ols ? 1 : 10.0^(3/4))
    momentum: float = 0.99
    learning_rate_scale_factor: float = 10.0 ** (1 / 4)
    max_steps: int = 1000
    convergence_threshold: float = 1e-5
    convergence_steps: int = 5
    grad_diagnostics: bool = False

    def __post_init__(self):
        if self.learning_rate is None:
            self.learning_rate = 1.0 if self.ls_type == "ols" else 10.0 ** (3 / 4)


def optimise_weights(dd: DeviceDesign, covariance_log, options: Optional[OptimOptions] = None, grad_fn=None
                     ) -> Tuple[np.ndarray, sp.csc_matrix, List[float]]:
    """gls_ / wls_ / ols_optimise_weights (src/optimise_weights.jl:160-321, 398-555, 613-790): Nesterov
    gradient descent on the logarithms of the shot weights with the reference's restart / learning-rate
    rule and convergence test.  Returns (shot_weights, covariance_log at those weights, merit_descent);
    the reference returns the design with `shot_weights` replaced (src/optimise_weights.jl:313-314).
    grad_fn(ls_type, shot_weights, covariance_log_unweighted, gate_transform_matrix) -> (grad, merit)
    replaces the device gradient (a test hook: the loop itself is host logic and can be driven by any
    gradient; `dd` may then be a FlatDesign)."""
    opt = options or OptimOptions()
    fd = dd.fd if hasattr(dd, "fd") else dd
    ls_type, est_type, est_weight = opt.ls_type, opt.est_type, opt.est_weight
    T = fd.tuple_number
    ml = np.asarray(fd.mapping_lengths, dtype=np.int64)
    tau = np.asarray(fd.tuple_times, dtype=np.float64)
    shot_weights = project_simplex_host(np.asarray(fd.shot_weights, dtype=np.float64))
    sfi = get_shot_weights_factor_inv(np.asarray(fd.shot_weights, dtype=np.float64), tau, ml)
    covariance_log = sp.csc_matrix(covariance_log)
    covariance_log_unweighted = sp.csc_matrix(covariance_log @ sp.diags(sfi))
    cu_val = dd.pattern_values(covariance_log_unweighted) if grad_fn is None else None
    gdiag = sp.diags(np.asarray(fd.gate_eigenvalues, dtype=np.float64))
    if est_type in ("sum", "prod"):
        transforms = [get_transform(fd, "ordinary") @ gdiag, get_transform(fd, "relative") @ gdiag]
    else:
        transforms = [get_transform(fd, est_type) @ gdiag]

    def grad_and_merit(w):
        if grad_fn is None:
            res = [calc_merit_grad_log(dd, ls_type, w, cu_val, Tm) for Tm in transforms]
        else:
            res = [grad_fn(ls_type, w, covariance_log_unweighted, Tm) for Tm in transforms]
        if est_type == "sum":
            (g0, m0), (g1, m1) = res
            return est_weight * g0 + (1 - est_weight) * g1, est_weight * m0 + (1 - est_weight) * m1
        if est_type == "prod":
            (g0, m0), (g1, m1) = res
            merit = m0 ** est_weight * m1 ** (1 - est_weight)
            return merit * (est_weight * g0 / m0 + (1 - est_weight) * g1 / m1), merit
        return res[0]

    stepping, step = True, 1
    recently_pruned = recently_zeroed = 0
    scaled_learning_rate = opt.learning_rate
    old_shot_weights = shot_weights.copy()
    velocity = np.zeros(T)
    merit_descent: List[float] = []
    while stepping:
        nesterov_log = -np.log(shot_weights) + opt.momentum * velocity
        nesterov_shot_weights = np.exp(-nesterov_log) / np.sum(np.exp(-nesterov_log))
        merit_grad_log, merit = grad_and_merit(nesterov_shot_weights)
        merit_descent.append(float(merit))
        velocity = opt.momentum * velocity - scaled_learning_rate * merit_grad_log
        log_update = -np.log(shot_weights) + velocity
        shot_weights_update = np.exp(-log_update) / np.sum(np.exp(-log_update))
        if len(merit_descent) >= 2 and merit_descent[-1] > merit_descent[-2]:
            shot_weights = old_shot_weights
            velocity = np.zeros(T)
            if recently_zeroed > 0:
                scaled_learning_rate /= opt.learning_rate_scale_factor
            if opt.grad_diagnostics:
                print(f"The updated shot weights worsened the figure of merit in step {step}; the velocity and shot "
                      "weights have been reset" + (" and the learning rate has been reduced." if recently_zeroed > 0 else "."))
            recently_pruned += 1
            recently_zeroed = opt.convergence_steps
        else:
            old_shot_weights = nesterov_shot_weights.copy()
            shot_weights = shot_weights_update
        if step >= opt.convergence_steps:
            last = merit_descent[step - opt.convergence_steps:step]
            merit_converged = all(abs(1 - b / a) < opt.convergence_steps * opt.convergence_threshold for a in last for b in last) \
                and abs(1 - merit_descent[-2] / merit_descent[-1]) < opt.convergence_threshold
        else:
            merit_converged = False
        if (merit_converged or step >= opt.max_steps) and recently_pruned == 0:
            stepping = False
        else:
            step += 1
            if recently_pruned > 0:
                recently_pruned -= 1
            if recently_zeroed > 0:
                recently_zeroed -= 1
    sf = get_shot_weights_factor(shot_weights, tau, ml)
    return project_simplex_host(shot_weights), sp.csc_matrix(covariance_log_unweighted @ sp.diags(sf)), merit_descent
