"""fit_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy / scipy) restatement of the reference's covariance assembly, least-squares fit and
estimator-covariance formulas.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.

Each function cites the reference code it follows (paths relative to the QuantumACES.jl root).
The design is consumed through the flat arrays of aces_b200.FlatDesign.

Third-party arithmetic that lives outside /root/reference (no Manifest.toml; compat ranges in
Project.toml:23-39) and how it is restated:
  * SuiteSparse SPQR behind sparse `\\` (src/estimate.jl:500-502): the published algorithm is a
    Householder QR least-squares solve; restated with LAPACK's QR-based dense solver
    (scipy.linalg.lstsq, driver gelsy) on the same whitened system (F*A) x = F*b.
  * SuiteSparse CHOLMOD behind cholesky(::SparseMatrixCSC) (src/merit.jl:398): restated with
    LAPACK dpotrf per tuple block.  CHOLMOD applies a fill-reducing permutation, so only
    F'F = inv(covariance) is comparable, never F itself (test/merit_tests.jl:531-543 tests
    exactly that invariant).
  * LAPACK via LinearAlgebra (cholesky / inv / bunchkaufman, src/merit.jl:564,586,614): LAPACK.

Parity pin: Julia is not available in this image and the reference ships no golden files for
this path, so this oracle is pinned against the reference's own identities, restated in
tests/test_fit_oracle.py: F'F*cov == I and sparse-factor solve == dense-factor solve
(test/merit_tests.jl:531-549), the same after clipping (:569-591), precision ==
inv(cholesky(gls covariance)) (:561-567), GLS == WLS for a diagonal covariance
(test/design_tests.jl:243-246) and the exact "all eigenvalues == 1.0" known answer
(test/device_tests.jl:144-155).  It has NOT been compared with outputs of the Julia package.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


def mean_ensemble(experiment_ensemble):
    """mean.(ensemble[i]) (src/estimate.jl:711-712, src/merit.jl:301-303)."""
    return [np.array([np.sum(np.asarray(v, dtype=np.float64)) / len(v) for v in tup], dtype=np.float64)
            for tup in experiment_ensemble]


def get_eigenvalues(fd, gate_eigenvalues):
    """src/design.jl:255-258."""
    return np.exp(fd.matrix.astype(np.float64) @ np.log(gate_eigenvalues))


def get_eigenvalues_ensemble(fd, gate_eigenvalues):
    lam = get_eigenvalues(fd, gate_eigenvalues)
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    return [lam[off[t]:off[t + 1]] for t in range(fd.tuple_number)]


def calc_covariance_eigenvalues(fd, gate_eigenvalues):
    """src/merit.jl:171-190: exp(design_row' * log g) for every off-diagonal covariance key."""
    lg = np.log(gate_eigenvalues)
    out = []
    for t in range(fd.tuple_number):
        if fd.cov_keys and len(fd.cov_keys[t]):
            out.append(np.exp(fd.cov_rows[t].astype(np.float64) @ lg))
        else:
            out.append(np.zeros(0))
    return out


def calc_covariance(fd, eigenvalues_ensemble, covariance_ensemble, epsilon=1e-10, weight_time=True):
    """src/merit.jl:206-291.  Returns the block-diagonal M x M covariance as scipy CSC."""
    blocks = []
    for t in range(fd.tuple_number):
        X = float(fd.experiment_numbers[t])
        w = float(fd.shot_weights[t])
        m = int(fd.mapping_lengths[t])
        lam = np.asarray(eigenvalues_ensemble[t], dtype=np.float64)
        cd = np.asarray(fd.diag_counts[t], dtype=np.float64)
        # diagonal (src/merit.jl:245-254)
        factor = (1 / w) * (X / cd)
        diag = factor * np.maximum(1 - lam ** 2, epsilon)
        rows = list(range(m))
        cols = list(range(m))
        vals = list(diag)
        if fd.full_covariance and fd.cov_keys and len(fd.cov_keys[t]):
            keys = fd.cov_keys[t]
            lam12 = np.asarray(covariance_ensemble[t], dtype=np.float64)
            c12 = np.asarray(fd.cov_counts[t], dtype=np.float64)
            p, q = keys[:, 0], keys[:, 1]
            # src/merit.jl:269-276
            f = (1 / w) * ((X * c12) / (cd[p] * cd[q]))
            cov = f * (lam12 - lam[p] * lam[q])
            rows += list(p) + list(q)
            cols += list(q) + list(p)
            vals += list(cov) + list(cov)
        blocks.append(sp.csc_matrix((vals, (rows, cols)), shape=(m, m)))
    covariance = sp.block_diag(blocks, format="csc")
    if weight_time:
        times_factor = float(np.sum(np.asarray(fd.shot_weights) * np.asarray(fd.tuple_times)))
        covariance = times_factor * covariance
    return sp.csc_matrix(covariance)


def calc_covariance_from_experiments(fd, eig_exp_ens, cov_exp_ens, epsilon=1e-10, weight_time=True):
    """src/merit.jl:292-313."""
    return calc_covariance(fd, mean_ensemble(eig_exp_ens),
                           mean_ensemble(cov_exp_ens) if cov_exp_ens else [[] for _ in eig_exp_ens],
                           epsilon=epsilon, weight_time=weight_time)


def calc_covariance_log(covariance, eigenvalues):
    """src/merit.jl:356-365: sparse(Symmetric(D^-1 * cov * D^-1)) -- Symmetric takes the upper triangle."""
    dinv = sp.diags(np.asarray(eigenvalues, dtype=np.float64) ** (-1.0))
    c = sp.csc_matrix(dinv @ covariance @ dinv)
    upper = sp.triu(c, k=0, format="csc")
    strict = sp.triu(c, k=1, format="csc")
    return sp.csc_matrix(upper + strict.T)


def get_clipped_indices(est_eigenvalues, min_eigenvalue=0.1):
    """src/estimate.jl:435-460 (0-based indices)."""
    est = np.asarray(est_eigenvalues)
    assert min_eigenvalue > 0.0
    if np.any(est > 1.0):
        raise ValueError(f"{int(np.sum(est > 1.0))} of {len(est)} estimated eigenvalues are larger than 1")
    return np.nonzero(~(est < min_eigenvalue))[0]


def get_clipped_mapping_lengths(mapping_lengths, clipped_indices):
    """src/estimate.jl:467-484."""
    off = np.concatenate([[0], np.cumsum(mapping_lengths)])
    ci = np.asarray(clipped_indices)
    return np.array([int(np.sum((ci >= off[t]) & (ci < off[t + 1]))) for t in range(len(mapping_lengths))],
                    dtype=np.int64)


def sparse_covariance_inv_factor(covariance_log, mapping_lengths, epsilon=1e-12):
    """src/merit.jl:378-427.  F with F'F = inv(covariance_log); per block F_t = inv(L_t) with
    covariance_log_t = L_t L_t' (no fill-reducing permutation: see the module docstring).  On a
    failed factorisation: elementwise abs, then shift the smallest eigenvalue up to epsilon."""
    cl = sp.csc_matrix(covariance_log)
    off = np.concatenate([[0], np.cumsum(mapping_lengths)])
    blocks = []
    for t in range(len(mapping_lengths)):
        blk = cl[off[t]:off[t + 1], off[t]:off[t + 1]].toarray()
        if blk.shape[0] == 0:
            blocks.append(np.zeros((0, 0)))
            continue
        try:
            L = np.linalg.cholesky(blk)
        except np.linalg.LinAlgError:
            blk = np.abs(blk)
            blk = np.triu(blk) + np.triu(blk, 1).T  # sparse(Symmetric(abs.(block)))
            eig_min = float(sla.eigvalsh(blk, subset_by_index=[0, 0])[0])
            if eig_min < epsilon:
                blk = blk + (epsilon - eig_min) * np.eye(blk.shape[0])
            L = np.linalg.cholesky(blk)
        blocks.append(sla.solve_triangular(L, np.eye(L.shape[0]), lower=True))
    return sp.csc_matrix(sp.block_diag([sp.csc_matrix(b) for b in blocks], format="csc"))


def sparse_covariance_inv(covariance_log, mapping_lengths, epsilon=1e-12):
    """src/merit.jl:434-448."""
    F = sparse_covariance_inv_factor(covariance_log, mapping_lengths, epsilon=epsilon)
    return sp.csc_matrix(F.T @ F)


def estimate_gate_eigenvalues(design_matrix, est_eigenvalues, factor=None, constrain=False):
    """src/estimate.jl:492-524: x = (F*A) \\ (F*(-log y)); g = exp(-x).  `\\` on a sparse
    rectangular matrix is a QR least-squares solve; restated with LAPACK (gelsy)."""
    A = sp.csc_matrix(design_matrix).astype(np.float64)
    y = np.asarray(est_eigenvalues, dtype=np.float64)
    b = -np.log(y)
    if factor is None:
        FA, Fb = A.toarray(), b
    else:
        F = sp.csc_matrix(factor)
        FA, Fb = (F @ A).toarray(), F @ b
    x, *_ = sla.lstsq(FA, Fb, lapack_driver="gelsy", cond=1e-15)
    if constrain:
        x[x < 0.0] = 0.0
    return np.exp(-x)


def calc_precision_matrix(design_matrix, gate_eigenvalues, covariance_log_inv):
    """src/merit.jl:754-770."""
    A = sp.csc_matrix(design_matrix).astype(np.float64)
    dinv = sp.diags(np.asarray(gate_eigenvalues, dtype=np.float64) ** (-1.0))
    P = sp.csc_matrix(dinv @ A.T @ sp.csc_matrix(covariance_log_inv) @ A @ dinv)
    return sp.csc_matrix(sp.triu(P, 0) + sp.triu(P, 1).T)  # sparse(Symmetric(.))


def _inv_chol(S):
    c = sla.cho_factor(S, lower=True)
    return sla.cho_solve(c, np.eye(S.shape[0]))


def _symmetric(S):
    return np.triu(S) + np.triu(S, 1).T  # Julia Symmetric(.) uses the upper triangle


def calc_gls_covariance(fd, covariance_log):
    """src/merit.jl:556-568."""
    A = sp.csc_matrix(fd.matrix).astype(np.float64)
    g = np.asarray(fd.gate_eigenvalues, dtype=np.float64)
    inv = sparse_covariance_inv(covariance_log, fd.mapping_lengths)
    G = _symmetric((A.T @ inv @ A).toarray())
    return _symmetric(g[:, None] * _inv_chol(G) * g[None, :])


def calc_wls_covariance(fd, covariance_log):
    """src/merit.jl:581-598."""
    A = sp.csc_matrix(fd.matrix).astype(np.float64)
    g = np.asarray(fd.gate_eigenvalues, dtype=np.float64)
    cl = sp.csc_matrix(covariance_log)
    Winv = sp.diags(1.0 / cl.diagonal())
    est = _inv_chol(_symmetric((A.T @ Winv @ A).toarray())) @ (A.T @ Winv).toarray()
    return _symmetric(g[:, None] * (est @ cl.toarray() @ est.T) * g[None, :])


def calc_ols_covariance(fd, covariance_log):
    """src/merit.jl:611-624 (bunchkaufman inverse of the symmetric A'A; LAPACK here too)."""
    A = sp.csc_matrix(fd.matrix).astype(np.float64)
    g = np.asarray(fd.gate_eigenvalues, dtype=np.float64)
    cl = sp.csc_matrix(covariance_log)
    est = np.linalg.inv(_symmetric((A.T @ A).toarray())) @ A.T.toarray()
    return _symmetric(g[:, None] * (est @ cl.toarray() @ est.T) * g[None, :])


def nrmse_moments(cov):
    """src/merit.jl:672-682."""
    tr = float(np.trace(cov))
    tr_sq = float(np.sum(cov * cov.T))
    N = cov.shape[0]
    expectation = np.sqrt(tr) * (1 - tr_sq / (4 * tr ** 2)) / np.sqrt(N)
    variance = (tr_sq / (2 * tr)) * (1 - tr_sq / (8 * tr ** 2)) / N
    return expectation, variance


def estimate_gate_noise_core(fd, eig_exp_ens, cov_exp_ens, min_eigenvalue=0.1):
    """The numeric core of estimate_gate_noise (src/estimate.jl:703-743, 820-831): means,
    covariance, log-covariance, clipping and the unprojected OLS / WLS / GLS gate eigenvalues."""
    est_eigenvalues = np.concatenate(mean_ensemble(eig_exp_ens))
    est_cov = calc_covariance_from_experiments(fd, eig_exp_ens, cov_exp_ens, weight_time=False)
    est_cov_log = calc_covariance_log(est_cov, est_eigenvalues)
    clipped = get_clipped_indices(est_eigenvalues, min_eigenvalue=min_eigenvalue)
    A = sp.csr_matrix(fd.matrix)[clipped, :].astype(np.float64)
    y = est_eigenvalues[clipped]
    cml = get_clipped_mapping_lengths(fd.mapping_lengths, clipped)
    cl = sp.csc_matrix(est_cov_log)[clipped, :][:, clipped]
    dfac = sp.diags(cl.diagonal() ** (-0.5))
    out = {
        "est_eigenvalues": est_eigenvalues, "clipped_indices": clipped, "covariance_log": cl,
        "ols": estimate_gate_eigenvalues(A, y),
        "wls": estimate_gate_eigenvalues(A, y, dfac),
    }
    if fd.full_covariance:
        F = sparse_covariance_inv_factor(cl, cml)
        out["gls"] = estimate_gate_eigenvalues(A, y, F)
    else:
        out["gls"] = out["wls"]
    return out
