"""optimise_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy, dense) restatement of the figure-of-merit gradients of the reference's shot-weight
optimiser (src/optimise_weights.jl) and of the gate-eigenvalue transforms they use
(src/noise.jl:66-124,679-741, src/merit.jl:455-470).  Only tests/ and bench.py's cpu legs may
import this module; the product never does.

The formulas are transcribed line by line with DENSE M x M / M x N intermediates, exactly as the
reference forms them (the product uses an algebraically equal trace form that never builds an M x M
matrix, so this is an independent check of that algebra).  Sizes: rotated d = 3 (M = 1248) runs in
seconds; d = 5 in about a minute.

Parity pin: Julia is not available here; the pin is the reference's own test of these gradients,
which compares them with automatic differentiation of the figure of merit
(test/merit_tests.jl:185,319,432).  tests/test_optimise_oracle.py restates that check with central
finite differences of `merit` below.  NOT compared with outputs of the Julia package.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

# src/noise.jl:66-124 (1-based Pauli indices inside a gate's block)
_ONE_TRIVIAL = [[1], [2], [3]]
_TWO_TRIVIAL = [[i] for i in range(1, 16)]
ORBITS = {
    "I": _ONE_TRIVIAL, "IM": _ONE_TRIVIAL, "X": _ONE_TRIVIAL, "Y": _ONE_TRIVIAL, "Z": _ONE_TRIVIAL,
    "H": [[1, 2], [3]], "S": [[1, 3], [2]], "S_DAG": [[1, 3], [2]],
    "II": _TWO_TRIVIAL,
    "CX": [[1, 3], [2], [4], [5, 7], [6], [8, 12], [9, 15], [10, 14], [11, 13]],
    "CNOT": [[1, 3], [2], [4], [5, 7], [6], [8, 12], [9, 15], [10, 14], [11, 13]],
    "CZ": [[1, 9], [2, 6], [3, 15], [4], [5, 13], [7, 11], [8], [10, 14], [12]],
}
for _p in ("XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"):
    ORBITS[_p] = _TWO_TRIVIAL
_SPAM = ("PZ", "PX", "PY", "MZ", "MX", "MY")
_MID = ("M", "R")


def get_transform(gate_types, gate_ptr, est_type):
    """get_transform (src/merit.jl:455-470) -> get_ordinary / marginal / relative_transform
    (src/noise.jl:679-741); gate g of type gate_types[g] owns columns gate_ptr[g]..gate_ptr[g+1]-1."""
    N = int(gate_ptr[-1])
    if est_type == "ordinary":
        return sp.identity(N, format="csc", dtype=np.float64)
    rows, cols, vals = [], [], []
    r = 0
    for g, ty in enumerate(gate_types):
        idx = int(gate_ptr[g])
        additive = ty in _SPAM or ty in _MID or ty == "IM"        # is_additive, strict = false (src/tableau.jl:736-743)
        if est_type == "relative" and additive:
            continue
        if ty in _SPAM or ty in _MID:
            rows.append(r); cols.append(idx); vals.append(1.0)
            r += 1
        else:
            for k, orbit in enumerate(ORBITS[ty]):
                for o in orbit:
                    rows.append(r + k); cols.append(idx + o - 1); vals.append(1.0 / len(orbit))
            r += len(ORBITS[ty])
    assert est_type in ("marginal", "relative")
    return sp.csc_matrix((vals, (rows, cols)), shape=(r, N))


def shot_weights_factor(shot_weights, tuple_times, mapping_lengths):
    """get_shot_weights_factor (src/optimise_weights.jl:6-26), as the diagonal vector."""
    times_factor = np.sum(shot_weights * tuple_times)
    return times_factor * np.concatenate([(1 / shot_weights[i]) * np.ones(int(m)) for i, m in enumerate(mapping_lengths)])


def shot_weights_factor_inv(shot_weights, tuple_times, mapping_lengths):
    """get_shot_weights_factor_inv (src/optimise_weights.jl:33-49)."""
    times_factor = np.sum(shot_weights * tuple_times)
    return (1 / times_factor) * np.concatenate([shot_weights[i] * np.ones(int(m)) for i, m in enumerate(mapping_lengths)])


def shot_weights_local_grad(shot_weights, tuple_times):
    """src/optimise_weights.jl:56-64."""
    return -np.sum(shot_weights * tuple_times) / shot_weights ** 2


def get_merit_grad(sigma_tr, sigma_tr_grad, sigma_sq_tr, sigma_sq_tr_grad, N):
    """src/optimise_weights.jl:71-86."""
    return (sigma_tr_grad * (((sigma_tr ** 2 / 2) + (3 * sigma_sq_tr / 8)) / (sigma_tr ** (5 / 2)))
            - sigma_sq_tr_grad * (1 / (4 * sigma_tr ** (3 / 2)))) / np.sqrt(N)


def shot_weights_log_matrix(shot_weights):
    """src/optimise_weights.jl:93-96."""
    return np.outer(shot_weights, shot_weights) - np.diag(shot_weights)


def _tuple_sums(v, mapping_lengths):
    off = np.concatenate([[0], np.cumsum(mapping_lengths)]).astype(np.int64)
    return np.array([np.sum(v[off[t]:off[t + 1]]) for t in range(len(mapping_lengths))])


def _finish(shot_weights, tuple_times, sigma_tr, sigma_sq_tr, tr_partial, sq_partial, N):
    local = shot_weights_local_grad(shot_weights, tuple_times)
    sigma_tr_grad = np.sum(tr_partial / shot_weights) * tuple_times + tr_partial * local
    sigma_sq_tr_grad = np.sum(sq_partial / shot_weights) * tuple_times + sq_partial * local
    merit = np.sqrt(sigma_tr) * (1 - sigma_sq_tr / (4 * sigma_tr ** 2)) / np.sqrt(N)
    merit_grad = get_merit_grad(sigma_tr, sigma_tr_grad, sigma_sq_tr, sigma_sq_tr_grad, N)
    return shot_weights_log_matrix(shot_weights) @ merit_grad, float(merit)


def _inv_chol(S):
    c = sla.cho_factor(S, lower=True)
    return sla.cho_solve(c, np.eye(S.shape[0]))


def calc_gls_merit_grad_log(fd, shot_weights, covariance_log_unweighted_inv, gate_transform_matrix):
    """calc_gls_merit_grad_log (src/optimise_weights.jl:103-152), dense."""
    A = fd.matrix.astype(np.float64).toarray()
    ml = fd.mapping_lengths
    sfi = shot_weights_factor_inv(shot_weights, fd.tuple_times, ml)
    Tm = gate_transform_matrix.toarray()
    covariance_log_inv = covariance_log_unweighted_inv.toarray() * sfi[None, :]
    sigma_prime = _inv_chol(A.T @ covariance_log_inv @ A)
    design_weights = A.T * sfi[None, :]
    covariance_sigma_prime = covariance_log_inv @ A @ sigma_prime
    N = Tm.shape[0]
    sigma = Tm @ sigma_prime @ Tm.T
    gate_sigma_prime = Tm.T @ Tm @ sigma_prime
    gls_matrix = covariance_sigma_prime @ gate_sigma_prime
    sigma_tr = np.trace(sigma)
    sigma_sq_tr = np.trace(sigma @ sigma)
    tr_diag = np.einsum("ij,ji->i", gls_matrix, design_weights)
    sq_diag = 2 * np.einsum("ij,ji->i", gls_matrix @ gate_sigma_prime, design_weights)
    return _finish(shot_weights, fd.tuple_times, sigma_tr, sigma_sq_tr, _tuple_sums(tr_diag, ml), _tuple_sums(sq_diag, ml), N)


def calc_wls_merit_grad_log(fd, shot_weights, covariance_log_unweighted, gate_transform_matrix):
    """calc_wls_merit_grad_log (src/optimise_weights.jl:327-391), dense."""
    A = fd.matrix.astype(np.float64).toarray()
    ml = fd.mapping_lengths
    sf = shot_weights_factor(shot_weights, fd.tuple_times, ml)
    Tm = gate_transform_matrix.toarray()
    cu = covariance_log_unweighted.toarray()
    cu = np.triu(cu) + np.triu(cu, 1).T                        # Symmetric(Array(.)) reads the upper triangle
    cov = covariance_log_unweighted.toarray() * sf[None, :]
    cov_m = np.triu(cov) + np.triu(cov, 1).T
    dinv = 1.0 / np.diag(cov)
    covariance_combined = cov * dinv[None, :]
    conjugated_inv = _inv_chol((A.T * dinv[None, :]) @ A)
    wls_estimator_prime = conjugated_inv @ (A.T * dinv[None, :])
    commutant = covariance_combined @ A @ wls_estimator_prime - covariance_combined
    N = Tm.shape[0]
    wls_estimator = Tm @ wls_estimator_prime
    sigma = wls_estimator @ cov_m @ wls_estimator.T
    wls_gram = wls_estimator.T @ wls_estimator
    wls_sq_gram = wls_gram @ cov_m @ wls_gram
    sigma_tr = np.trace(sigma)
    sigma_sq_tr = np.trace(sigma @ sigma)

    def sym(X):
        return np.triu(X) + np.triu(X, 1).T

    tr_diag = np.einsum("ij,ji->i", sym(wls_gram + 2 * np.diag(np.diag(wls_gram @ commutant))), cu)
    sq_diag = 2 * np.einsum("ij,ji->i", sym(wls_sq_gram + 2 * np.diag(np.diag(wls_sq_gram @ commutant))), cu)
    return _finish(shot_weights, fd.tuple_times, sigma_tr, sigma_sq_tr, _tuple_sums(tr_diag, ml), _tuple_sums(sq_diag, ml), N)


def ols_matrices(fd, covariance_log_unweighted, gate_transform_matrix):
    """The weight-independent matrices ols_optimise_weights builds once (src/optimise_weights.jl:651-658)."""
    A = fd.matrix.astype(np.float64).toarray()
    Tm = gate_transform_matrix.toarray()
    ols_estimator = Tm @ np.linalg.inv(A.T @ A) @ A.T          # inv(bunchkaufman(Symmetric(.)))
    ols_estimator_covariance = ols_estimator @ covariance_log_unweighted.toarray()
    ols_gram_covariance = ols_estimator.T @ ols_estimator_covariance
    return ols_estimator, ols_estimator_covariance, ols_gram_covariance


def calc_ols_merit_grad_log(fd, shot_weights, ols_estimator, ols_estimator_covariance, ols_gram_covariance):
    """calc_ols_merit_grad_log (src/optimise_weights.jl:563-606), dense."""
    ml = fd.mapping_lengths
    sf = shot_weights_factor(shot_weights, fd.tuple_times, ml)
    N = ols_estimator.shape[0]
    sigma = (ols_estimator_covariance * sf[None, :]) @ ols_estimator.T
    sigma_tr = np.trace(sigma)
    sigma_sq_tr = np.trace(sigma @ sigma)
    tr_diag = np.diag(ols_gram_covariance)
    sq_diag = 2 * np.einsum("ij,ji->i", ols_gram_covariance * sf[None, :], ols_gram_covariance)
    return _finish(shot_weights, fd.tuple_times, sigma_tr, sigma_sq_tr, _tuple_sums(tr_diag, ml), _tuple_sums(sq_diag, ml), N)


def merit_of_weights(fd, ls_type, shot_weights, covariance_log_unweighted, gate_transform_matrix):
    """The figure of merit itself as a function of the shot weights (what the gradients differentiate):
    sqrt(tr S) (1 - tr S^2 / (4 tr^2 S)) / sqrt(N) of the estimator covariance S at those weights."""
    A = fd.matrix.astype(np.float64).toarray()
    ml = fd.mapping_lengths
    sf = shot_weights_factor(shot_weights, fd.tuple_times, ml)
    Tm = gate_transform_matrix.toarray()
    cu = covariance_log_unweighted.toarray()
    cov = cu * sf[None, :]
    if ls_type == "gls":
        Sp = _inv_chol(A.T @ np.linalg.inv(cov) @ A)
        S = Tm @ Sp @ Tm.T
    else:
        w = 1.0 / np.diag(cov) if ls_type == "wls" else np.ones(cov.shape[0])
        est = Tm @ _inv_chol((A.T * w[None, :]) @ A) @ (A.T * w[None, :])
        S = est @ cov @ est.T
    tr, sq = np.trace(S), np.trace(S @ S)
    return float(np.sqrt(tr) * (1 - sq / (4 * tr ** 2)) / np.sqrt(Tm.shape[0]))
