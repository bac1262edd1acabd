"""GPU parity tests for the fit path: CUDA (through the C ABI) against the numpy/scipy oracle.
Floating point: relative tolerance 1e-9 in fp64 (BASELINE.json north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

from fit_util import fo, noisy_ensembles

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel(x, y):
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / max(np.linalg.norm(np.asarray(y)), 1e-300)


@pytest.fixture(scope="module")
def design():
    import aces_b200

    return aces_b200.synthetic_design("rotated", 3, full_covariance=True, with_matrix=True)


@pytest.fixture(scope="module")
def inputs(design):
    fd = design
    eig, cov = noisy_ensembles(fd, seed=11)
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
    F = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    return dict(eig=eig, cov=cov, lam=lam, cl=cl, F=F)


def test_fit_known_answers(ctx, design):
    import aces_b200

    A = design.matrix
    ones = np.ones(A.shape[0])
    # noiseless records => gate eigenvalues exactly 1.0 (test/device_tests.jl:144-155)
    assert np.all(aces_b200.estimate_gate_eigenvalues(ctx, A, ones) == 1.0)
    assert np.all(aces_b200.estimate_gate_eigenvalues(ctx, A, ones, sp.identity(A.shape[0]) * 3.0) == 1.0)
    # exact recovery from noiseless eigenvalues
    y = fo.get_eigenvalues(design, design.gate_eigenvalues)
    g = aces_b200.estimate_gate_eigenvalues(ctx, A, y)
    assert np.max(np.abs(g - design.gate_eigenvalues)) < 1e-12


def test_ols_wls_gls_vs_oracle(ctx, design, inputs):
    import aces_b200

    A, lam, cl, F = design.matrix, inputs["lam"], inputs["cl"], inputs["F"]
    ols = aces_b200.estimate_gate_eigenvalues(ctx, A, lam)
    assert rel(ols, fo.estimate_gate_eigenvalues(A, lam)) < RTOL
    dfac = sp.diags(cl.diagonal() ** -0.5)
    wls = aces_b200.estimate_gate_eigenvalues(ctx, A, lam, dfac)
    assert rel(wls, fo.estimate_gate_eigenvalues(A, lam, dfac)) < RTOL
    gls = aces_b200.estimate_gate_eigenvalues(ctx, A, lam, F)
    assert rel(gls, fo.estimate_gate_eigenvalues(A, lam, F)) < RTOL
    # the log-eigenvalues themselves (x = -log g) also agree to 1e-9 relative
    assert rel(np.log(gls), np.log(fo.estimate_gate_eigenvalues(A, lam, F))) < 1e-8
    # GLS with the diagonal factor == WLS (test/design_tests.jl:243-246)
    cld = sp.diags(cl.diagonal()).tocsc()
    gls_d = aces_b200.estimate_gate_eigenvalues(ctx, A, lam, fo.sparse_covariance_inv_factor(cld, design.mapping_lengths))
    assert rel(gls_d, wls) < RTOL
    c = aces_b200.estimate_gate_eigenvalues(ctx, A, lam, dfac, constrain=True)
    assert np.all(c <= 1.0) and rel(c, fo.estimate_gate_eigenvalues(A, lam, dfac, constrain=True)) < RTOL


def test_fit_after_clipping(ctx, design, inputs):
    import aces_b200

    A, lam, cl = design.matrix, inputs["lam"].copy(), inputs["cl"]
    rng = np.random.default_rng(5)
    # clip rows with several entries only: the single-entry rows carry the column rank
    multi = np.nonzero(np.diff(sp.csr_matrix(A).indptr) >= 2)[0]
    lam[rng.choice(multi, size=40, replace=False)] = 0.05  # below the 0.1 threshold
    ci = aces_b200.get_clipped_indices(lam, warning=False)
    assert np.array_equal(ci, fo.get_clipped_indices(lam))
    cml = aces_b200.get_clipped_mapping_lengths(design.mapping_lengths, ci)
    assert np.array_equal(cml, fo.get_clipped_mapping_lengths(design.mapping_lengths, ci))
    Ac = sp.csr_matrix(A)[ci, :]
    ccl = sp.csc_matrix(cl)[ci, :][:, ci]
    F = fo.sparse_covariance_inv_factor(ccl, cml)
    got = aces_b200.estimate_gate_eigenvalues(ctx, Ac, lam[ci], F)
    assert rel(got, fo.estimate_gate_eigenvalues(Ac, lam[ci], F)) < RTOL
    with pytest.raises(ValueError):
        aces_b200.get_clipped_indices(np.array([0.5, 1.5]))


def test_rank_deficient_gives_basic_solution(ctx):
    """(F*A) \\ (F*b) with a rank-deficient matrix: the reference's sparse QR returns a basic solution
    (src/estimate.jl:500-502); so does the CSC path (dependent column -> x = 0), reporting the rank
    deficiency.  The inverse of the singular Gram matrix stays an error."""
    import aces_b200

    A = sp.csc_matrix(np.array([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0]]))
    y = np.array([0.9, 0.8, 0.7])
    g = aces_b200.estimate_gate_eigenvalues(ctx, A, y)
    assert ctx.last_rank_deficiency() == 1
    assert g[1] == 1.0                                   # the second column is eliminated
    x0 = np.linalg.lstsq(np.array([[1.0], [2.0], [3.0]]), -np.log(y), rcond=None)[0][0]
    assert abs(-np.log(g[0]) - x0) < 1e-12
    with pytest.raises(aces_b200.NotPositiveDefinite):
        aces_b200.calc_gls_covariance_from_factor(ctx, A, np.ones(2), None)


def test_gls_covariance_and_precision(ctx, design):
    import aces_b200

    fd = design
    g = fd.gate_eigenvalues
    lam = fo.get_eigenvalues(fd, g)
    cov = fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, g), fo.calc_covariance_eigenvalues(fd, g))
    cl = fo.calc_covariance_log(cov, lam)
    F = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    gls = aces_b200.calc_gls_covariance_from_factor(ctx, fd.matrix, g, F)
    want = fo.calc_gls_covariance(fd, cl)
    assert rel(gls, want) < RTOL
    assert np.array_equal(gls, gls.T)
    prec = aces_b200.calc_precision_matrix_from_factor(ctx, fd.matrix, g, F)
    want_p = fo.calc_precision_matrix(fd.matrix, g, fo.sparse_covariance_inv(cl, fd.mapping_lengths)).toarray()
    assert rel(prec, want_p) < RTOL
    # test/merit_tests.jl:566: precision == inv(cholesky(gls covariance))
    assert rel(prec @ gls, np.eye(len(g))) < 1e-7
