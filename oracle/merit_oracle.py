"""merit_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy / scipy, dense) restatement of calc_merit(d::Design) (src/merit.jl:925-1006) for a design
with full covariance data: the three estimator covariances (fit_oracle.calc_*_covariance,
src/merit.jl:556-629), their marginal / relative transforms (src/merit.jl:477-509 with the transform
matrices of src/noise.jl:689-741, restated in optimise_oracle.get_transform), LAPACK `eigvalsh` for
LinearAlgebra.eigvals(::Symmetric) (the same dsyevr / dsyevd family Julia calls), nrmse_moments of the
spectra (src/merit.jl:684-692) and ARPACK (scipy.sparse.linalg.eigsh -- the library behind Arpack.jl)
for the two extreme singular values of the design matrix (src/merit.jl:993-1006).  Only tests/ and
bench.py's cpu legs may import this module; the product never does.

Parity pin: Julia is not available here.  The pin is the reference's own test of calc_merit
(test/merit_tests.jl:90-111, 223-244, 342-363): the Merit spectra equal eigvals(T * calc_ls_covariance
* T') and the Merit moments equal nrmse_moments of them -- restated in tests/test_merit_oracle.py
together with trace identities (sum of a spectrum = trace, sum of squares = trace of the square).
NOT compared with outputs of the Julia package: "parity unpinned" in that sense.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import fit_oracle as fo
import optimise_oracle as oo

LS = ("gls", "wls", "ols")
EST = ("", "_marginal", "_relative")


def calc_covariance_log(fd, gate_eigenvalues=None):
    """calc_covariance_log(d) (src/merit.jl:366-377): weight_time = true, epsilon = 1e-10."""
    g = np.asarray(fd.gate_eigenvalues if gate_eigenvalues is None else gate_eigenvalues, dtype=np.float64)
    eig = fo.get_eigenvalues_ensemble(fd, g)
    cov = fo.calc_covariance_eigenvalues(fd, g)
    covariance = fo.calc_covariance(fd, eig, cov, epsilon=1e-10, weight_time=True)
    return fo.calc_covariance_log(covariance, np.concatenate(eig))


def eigvals(symmetric):
    """eigvals(::Symmetric{Float64, Matrix{Float64}}): LAPACK, upper triangle referenced."""
    return sla.eigvalsh(np.asarray(symmetric), lower=False)


def nrmse_moments_eigenvalues(ev):
    """src/merit.jl:684-692."""
    ev = np.asarray(ev, dtype=np.float64)
    tr, tr_sq, N = float(np.sum(ev)), float(np.sum(ev ** 2)), len(ev)
    expectation = np.sqrt(tr) * (1 - tr_sq / (4 * tr ** 2)) / np.sqrt(N)
    variance = (tr_sq / (2 * tr)) * (1 - tr_sq / (8 * tr ** 2)) / N
    return expectation, variance


def design_matrix_extremes(fd):
    """src/merit.jl:993-1006: largest and smallest eigenvalue of d.matrix' * d.matrix (ARPACK :LM, and
    :LM with sigma = 0, i.e. shift-invert)."""
    A = sp.csc_matrix(fd.matrix).astype(np.float64)
    G = sp.csc_matrix(A.T @ A)
    largest = float(spla.eigsh(G, k=1, which="LM", return_eigenvectors=False)[0])
    smallest = float(spla.eigsh(G, k=1, which="LM", sigma=0.0, return_eigenvectors=False)[0])
    return largest, smallest


def calc_merit(fd, gate_eigenvalues=None, covariances=False):
    """Numeric fields of Merit (src/merit.jl:30-73) keyed by the reference's field names."""
    g = np.asarray(fd.gate_eigenvalues if gate_eigenvalues is None else gate_eigenvalues, dtype=np.float64)
    saved = fd.gate_eigenvalues
    fd.gate_eigenvalues = g  # fit_oracle.calc_*_covariance read get_gate_eigenvalues(d) from the design
    try:
        cov_log = calc_covariance_log(fd, g)
        sig = {"gls": fo.calc_gls_covariance(fd, cov_log), "wls": fo.calc_wls_covariance(fd, cov_log),
               "ols": fo.calc_ols_covariance(fd, cov_log)}
    finally:
        fd.gate_eigenvalues = saved
    T = {"": None, "_marginal": oo.get_transform(fd.gate_types, fd.gate_ptr, "marginal"),
         "_relative": oo.get_transform(fd.gate_types, fd.gate_ptr, "relative")}
    out = {"N": len(g), "N_marginal": T["_marginal"].shape[0], "N_relative": T["_relative"].shape[0]}
    for ls in LS:
        for est in EST:
            S = sig[ls]
            if T[est] is not None:
                t = sp.csr_matrix(T[est])
                S = fo._symmetric(np.asarray(t @ (t @ S).T))  # T S T' (S symmetric)
            ev = eigvals(S)
            out[f"{ls}{est}_cov_eigenvalues"] = ev
            out[f"{ls}{est}_expectation"], out[f"{ls}{est}_variance"] = nrmse_moments_eigenvalues(ev)
            if covariances:
                out[f"{ls}{est}_cov"] = S
    largest, smallest = design_matrix_extremes(fd)
    out["cond_num"] = float(np.sqrt(largest) / np.sqrt(smallest))
    out["pinv_norm"] = float(1.0 / np.sqrt(smallest))
    return out
