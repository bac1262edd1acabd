"""GPU tests of SURVEY 8f row f2: the symmetric eigenvalue solver (aces_symmetric_eigenvalues: panel
Householder tridiagonalisation + Sturm bisection) against LAPACK, and calc_merit's numeric core
(aces_design_merit) against the CPU restatement (oracle/merit_oracle.py)."""
import os
import sys

import numpy as np
import pytest

import aces_b200

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import merit_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu
EIG_TOL = 2e-13   # |lambda_gpu - lambda_lapack| <= EIG_TOL * ||A||_2: both are backward stable to eps * ||A||
MERIT_TOL = 1e-9  # relative, fp64 (north star)


@pytest.fixture(scope="module")
def ctx():
    c = aces_b200.Context(0)
    yield c
    c.close()


def random_symmetric(n, seed, spectrum=None):
    rng = np.random.default_rng(seed)
    if spectrum is None:
        a = rng.standard_normal((n, n))
        return (a + a.T) / 2
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return (q * spectrum) @ q.T


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 64, 65, 127, 128, 129, 200, 777, 1500])
def test_eigenvalues_vs_lapack(ctx, n):
    a = random_symmetric(n, 100 + n)
    want = np.linalg.eigvalsh(a)
    got = aces_b200.eigvals(ctx, a)
    assert np.all(np.diff(got) >= 0)
    assert np.max(np.abs(got - want)) <= EIG_TOL * max(np.abs(want).max(), 1e-300) * max(1.0, np.sqrt(n) / 8)


def test_only_the_upper_triangle_is_read(ctx):
    """Julia's Symmetric(A) (uplo = :U) ignores the strict lower triangle; so does the drop-in."""
    a = random_symmetric(300, 5)
    junk = a.copy()
    junk[np.tril_indices(300, -1)] = 1e30
    assert np.array_equal(aces_b200.eigvals(ctx, junk), aces_b200.eigvals(ctx, a))


def test_graded_spectrum_and_scaling(ctx):
    """A covariance-like spectrum over ten decades, and the same matrix scaled by 1e-150 / 1e+150 (the
    solver rescales by a power of two, so the results scale exactly)."""
    n = 400
    spec = np.logspace(-10, 0, n)
    a = random_symmetric(n, 9, spec)
    a = (a + a.T) / 2
    want = np.linalg.eigvalsh(a)
    got = aces_b200.eigvals(ctx, a)
    assert np.max(np.abs(got - want)) <= EIG_TOL * 4
    for p in (-498, 498):
        s = 2.0 ** p
        assert np.array_equal(aces_b200.eigvals(ctx, a * s), got * s)


def test_special_matrices(ctx):
    n = 260
    assert np.max(np.abs(aces_b200.eigvals(ctx, np.zeros((n, n))))) <= 1e-300   # within the pivmin guard of zero
    d = np.arange(n, dtype=np.float64)[::-1].copy()
    assert np.allclose(aces_b200.eigvals(ctx, np.diag(d)), np.sort(d), rtol=0, atol=1e-12 * n)
    one = np.ones((n, n))                        # rank one: eigenvalues 0 (n - 1 times) and n
    got = aces_b200.eigvals(ctx, one)
    assert abs(got[-1] - n) <= 1e-11 * n and np.max(np.abs(got[:-1])) <= 1e-11 * n
    # tridiagonal input (Householder vectors with zero tails): the 1-D Laplacian, known spectrum
    lap = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    want = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    assert np.max(np.abs(aces_b200.eigvals(ctx, lap) - want)) <= 1e-13 * 4


def test_repeated_calls_are_bit_identical(ctx):
    a = random_symmetric(513, 77)
    first = aces_b200.eigvals(ctx, a)
    for _ in range(3):
        assert np.array_equal(aces_b200.eigvals(ctx, a), first)


@pytest.fixture(scope="module", params=["example", "real3"])
def design(request):
    if request.param == "example":
        return aces_b200.example_design(full_covariance=True)
    return aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)


def test_calc_merit_vs_oracle(ctx, design):
    fd = design
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        got = aces_b200.calc_merit(dd)
        want = mo.calc_merit(fd)
        for key in ("N", "N_marginal", "N_relative"):
            assert got[key] == want[key]
        for ls in mo.LS:
            for est in mo.EST:
                k = f"{ls}{est}_cov_eigenvalues"
                scale = np.abs(want[k]).max()
                # the reference's own check is `≈` (rtol = sqrt(eps) on the norm); here per eigenvalue
                assert np.max(np.abs(got[k] - want[k])) <= MERIT_TOL * scale, k
                assert np.isclose(got[f"{ls}{est}_expectation"], want[f"{ls}{est}_expectation"], rtol=MERIT_TOL)
                assert np.isclose(got[f"{ls}{est}_variance"], want[f"{ls}{est}_variance"], rtol=MERIT_TOL)
        assert np.isclose(got["cond_num"], want["cond_num"], rtol=1e-8)
        assert np.isclose(got["pinv_norm"], want["pinv_norm"], rtol=1e-8)
        # the moments of the ordinary spectra are the trace moments of K6 (aces_design_ls_covariance)
        cov_log = dd.calc_covariance_log_of_gates(fd.gate_eigenvalues)
        for ls in mo.LS:
            _, tr, tr_sq = dd.calc_ls_covariance(ls, cov_log, want_matrix=False)
            e, v = aces_b200.nrmse_moments(tr, tr_sq, dd.N)
            assert np.isclose(got[f"{ls}_expectation"], e, rtol=1e-10)
            assert np.isclose(got[f"{ls}_variance"], v, rtol=1e-10)
    finally:
        dd.close()


def test_calc_merit_of_another_noise_instance(ctx, design):
    """update_noise + calc_merit as calc_ensemble_scaling does per repetition (src/scaling.jl:510-536):
    the same design at other gate eigenvalues."""
    fd = design
    rng = np.random.default_rng(12)
    g = 1.0 - (1.0 - np.asarray(fd.gate_eigenvalues)) * rng.uniform(0.5, 1.5, size=len(fd.gate_eigenvalues))
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        got = aces_b200.calc_merit(dd, gate_eigenvalues=g)
        want = mo.calc_merit(fd, gate_eigenvalues=g)
        for ls in mo.LS:
            assert np.isclose(got[f"{ls}_expectation"], want[f"{ls}_expectation"], rtol=MERIT_TOL)
            assert np.isclose(got[f"{ls}_relative_variance"], want[f"{ls}_relative_variance"], rtol=MERIT_TOL)
    finally:
        dd.close()
