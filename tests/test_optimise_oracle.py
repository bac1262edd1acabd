"""CPU tests: the dense restatement of the shot-weight gradients (oracle/optimise_oracle.py) against
central finite differences of the figure of merit -- the reference's own check of these functions
compares them with automatic differentiation (test/merit_tests.jl:185,319,432)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import aces_b200
from fit_util import fo

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import optimise_oracle as oo  # noqa: E402


@pytest.fixture(scope="module")
def setup():
    fd = aces_b200.example_design(full_covariance=True)
    lam = fo.get_eigenvalues_ensemble(fd, fd.gate_eigenvalues)
    lam12 = fo.calc_covariance_eigenvalues(fd, fd.gate_eigenvalues)
    cov = fo.calc_covariance(fd, lam, lam12)                       # weight_time = true (src/merit.jl:292-303)
    cov_log = fo.calc_covariance_log(cov, np.concatenate(lam))
    sfi = oo.shot_weights_factor_inv(fd.shot_weights, fd.tuple_times, fd.mapping_lengths)
    cu = sp.csc_matrix(cov_log.toarray() * sfi[None, :])         # covariance_log_unweighted (src/optimise_weights.jl:183-184)
    return fd, cu


def weights_of(ell):
    w = np.exp(-ell)
    return w / w.sum()


@pytest.mark.parametrize("est_type", ["ordinary", "marginal", "relative"])
@pytest.mark.parametrize("ls_type", ["gls", "wls", "ols"])
def test_gradient_matches_finite_differences(setup, ls_type, est_type):
    fd, cu = setup
    Tm = oo.get_transform(fd.gate_types, fd.gate_ptr, est_type) @ sp.diags(fd.gate_eigenvalues)
    rng = np.random.default_rng(3)
    w = fd.shot_weights * np.exp(0.3 * rng.standard_normal(fd.tuple_number))
    w /= w.sum()
    if ls_type == "gls":
        cu_inv = fo.sparse_covariance_inv(cu, fd.mapping_lengths)
        grad, merit = oo.calc_gls_merit_grad_log(fd, w, cu_inv, Tm)
    elif ls_type == "wls":
        grad, merit = oo.calc_wls_merit_grad_log(fd, w, cu, Tm)
    else:
        grad, merit = oo.calc_ols_merit_grad_log(fd, w, *oo.ols_matrices(fd, cu, Tm))
    assert abs(merit - oo.merit_of_weights(fd, ls_type, w, cu, Tm)) <= 1e-10 * merit
    ell = -np.log(w)
    h = 1e-5
    fdiff = np.zeros_like(ell)
    for j in range(len(ell)):
        e = np.zeros_like(ell)
        e[j] = h
        fdiff[j] = (oo.merit_of_weights(fd, ls_type, weights_of(ell + e), cu, Tm)
                    - oo.merit_of_weights(fd, ls_type, weights_of(ell - e), cu, Tm)) / (2 * h)
    assert np.linalg.norm(grad - fdiff) <= 1e-5 * np.linalg.norm(fdiff) + 1e-9


def test_transform_shapes(setup):
    fd, _ = setup
    N = int(fd.gate_ptr[-1])
    To = oo.get_transform(fd.gate_types, fd.gate_ptr, "ordinary")
    Tm = oo.get_transform(fd.gate_types, fd.gate_ptr, "marginal")
    Tr = oo.get_transform(fd.gate_types, fd.gate_ptr, "relative")
    assert To.shape == (N, N) and Tm.shape[1] == N and Tr.shape[1] == N
    assert Tr.shape[0] < Tm.shape[0] <= N
    # every row of an orbit average sums to one (src/noise.jl:706, 736)
    assert np.allclose(np.asarray(Tm.sum(axis=1)).ravel(), 1.0)
    assert np.allclose(np.asarray(Tr.sum(axis=1)).ravel(), 1.0)
