"""CPU tests: the restatement of calc_merit (oracle/merit_oracle.py) against the reference's own checks of
it (test/merit_tests.jl:90-111, 223-244, 342-363): the Merit spectra are eigvals(T * calc_ls_covariance * T'),
the Merit moments are nrmse_moments of the spectra, which equal nrmse_moments of the matrices."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import aces_b200
from fit_util import fo

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import merit_oracle as mo  # noqa: E402
import optimise_oracle as oo  # noqa: E402


@pytest.fixture(scope="module")
def merit():
    fd = aces_b200.example_design(full_covariance=True)
    return fd, mo.calc_merit(fd, covariances=True)


@pytest.mark.parametrize("ls", mo.LS)
@pytest.mark.parametrize("est", mo.EST)
def test_spectra_and_moments_match_the_matrices(merit, ls, est):
    fd, m = merit
    S, ev = m[f"{ls}{est}_cov"], m[f"{ls}{est}_cov_eigenvalues"]
    assert np.all(np.diff(ev) >= 0) and ev[0] > -1e-12 * ev[-1]
    assert np.isclose(ev.sum(), np.trace(S), rtol=1e-12)
    assert np.isclose((ev ** 2).sum(), np.sum(S * S), rtol=1e-12)
    e1, v1 = fo.nrmse_moments(S)                       # nrmse_moments(::Symmetric), src/merit.jl:672-682
    assert np.isclose(m[f"{ls}{est}_expectation"], e1, rtol=1e-12)
    assert np.isclose(m[f"{ls}{est}_variance"], v1, rtol=1e-12)


def test_transformed_covariances_are_the_reference_products(merit):
    """get_marginal_gate_covariance = T * Sigma * T' (src/merit.jl:477-491), dense."""
    fd, m = merit
    for est, name in (("_marginal", "marginal"), ("_relative", "relative")):
        T = oo.get_transform(fd.gate_types, fd.gate_ptr, name).toarray()
        for ls in mo.LS:
            want = T @ m[f"{ls}_cov"] @ T.T
            assert np.allclose(m[f"{ls}{est}_cov"], want, rtol=1e-12, atol=1e-14 * np.abs(want).max())
        assert m["N" + est] == T.shape[0]


def test_design_matrix_extremes_match_the_singular_values(merit):
    fd, m = merit
    sv = np.linalg.svd(sp.csc_matrix(fd.matrix).toarray().astype(np.float64), compute_uv=False)
    assert np.isclose(m["cond_num"], sv[0] / sv[-1], rtol=1e-8)
    assert np.isclose(m["pinv_norm"], 1.0 / sv[-1], rtol=1e-8)


def test_gls_is_at_least_as_good_as_wls_and_ols(merit):
    """GLS is the minimum-variance linear estimator: tr(Sigma_gls) <= tr(Sigma_wls), tr(Sigma_ols)."""
    _, m = merit
    assert m["gls_cov_eigenvalues"].sum() <= m["wls_cov_eigenvalues"].sum() * (1 + 1e-12)
    assert m["gls_cov_eigenvalues"].sum() <= m["ols_cov_eigenvalues"].sum() * (1 + 1e-12)
