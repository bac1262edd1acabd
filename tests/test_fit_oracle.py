"""CPU tests: the numpy/scipy fit oracle against the reference's own identities."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import aces_b200
from fit_util import fo, noisy_ensembles


def approx(x, y, rtol=np.sqrt(np.finfo(float).eps)):
    """Julia's `x ≈ y` for arrays: norm(x - y) <= rtol * max(norm(x), norm(y))."""
    x, y = np.asarray(x), np.asarray(y)
    return np.linalg.norm(x - y) <= rtol * max(np.linalg.norm(x), np.linalg.norm(y))


@pytest.fixture(scope="module", params=["synthetic", "real"])
def design(request):
    """The synthetic look-alike and the real rotated d=3 design (circuit_design.py: the array
    restatement of generate_design); both have M = 1248, N = 659."""
    if request.param == "real":
        return aces_b200.rotated_design(3, full_covariance=True)
    return aces_b200.synthetic_design("rotated", 3, full_covariance=True, with_matrix=True)


def test_shapes(design):
    assert design.matrix.shape == (1248, 659)
    assert np.linalg.matrix_rank(design.matrix.toarray().astype(float)) == 659


def test_covariance_closed_forms(design):
    # src/merit.jl:245-279 including the epsilon floor at lambda = +-1
    fd = design
    lam = [np.full(int(m), 0.9) for m in fd.mapping_lengths]
    lam[0][0] = 1.0
    lam[0][1] = -1.0
    lam12 = [np.full(len(k), 0.83) for k in fd.cov_keys]
    cov = fo.calc_covariance(fd, lam, lam12, weight_time=False).toarray()
    X, w, cd = fd.experiment_numbers[0], fd.shot_weights[0], fd.diag_counts[0]
    assert cov[0, 0] == (1 / w) * (X / cd[0]) * 1e-10
    assert cov[1, 1] == (1 / w) * (X / cd[1]) * 1e-10
    assert cov[2, 2] == (1 / w) * (X / cd[2]) * max(1 - 0.9 ** 2, 1e-10)
    t = next(t for t in range(fd.tuple_number) if len(fd.cov_keys[t]))   # first tuple with off-diagonal keys
    off = int(np.sum(fd.mapping_lengths[:t]))
    X, w, cd = fd.experiment_numbers[t], fd.shot_weights[t], fd.diag_counts[t]
    p, q = fd.cov_keys[t][-1]
    c = fd.cov_counts[t][-1]
    want = (1 / w) * ((X * c) / (cd[p] * cd[q])) * (0.83 - 0.9 * 0.9)
    assert cov[off + p, off + q] == want and cov[off + q, off + p] == want
    tf = float(np.sum(fd.shot_weights * fd.tuple_times))
    assert np.allclose(fo.calc_covariance(fd, lam, lam12, weight_time=True).toarray(), tf * cov, rtol=1e-15)


def test_inverse_factor_identities(design):
    # test/merit_tests.jl:531-549: F'F == inv(chol(dense)); whitened solves agree
    fd = design
    eig, cov = noisy_ensembles(fd, seed=1)
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov), lam)
    F = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    inv = fo.sparse_covariance_inv(cl, fd.mapping_lengths)
    dense_inv = np.linalg.inv(cl.toarray())
    assert np.allclose(inv.toarray(), (F.T @ F).toarray(), rtol=1e-12, atol=0)
    assert approx(inv.toarray(), dense_inv)
    assert np.allclose((inv @ cl).toarray(), np.eye(cl.shape[0]), atol=1e-9)
    rng = np.random.default_rng(2)
    M = cl.shape[0]
    data, mat = rng.random(M), rng.random((M, 40))
    Ld = np.linalg.inv(np.linalg.cholesky(cl.toarray()))
    x1 = np.linalg.lstsq(F @ mat, F @ data, rcond=None)[0]
    x2 = np.linalg.lstsq(Ld @ mat, Ld @ data, rcond=None)[0]
    assert approx(x1, x2)


def test_inverse_factor_after_clipping(design):
    # test/merit_tests.jl:569-591
    fd = design
    eig, cov = noisy_ensembles(fd, seed=3)
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov), lam)
    rng = np.random.default_rng(4)
    test_data = rng.random(cl.shape[0])
    ci = fo.get_clipped_indices(test_data)
    cml = fo.get_clipped_mapping_lengths(fd.mapping_lengths, ci)
    assert cml.sum() == len(ci) and len(ci) < cl.shape[0]
    ccl = cl[ci, :][:, ci]
    inv = fo.sparse_covariance_inv(ccl, cml)
    assert np.allclose((inv @ ccl).toarray(), np.eye(len(ci)), atol=1e-9)
    with pytest.raises(ValueError):
        fo.get_clipped_indices(np.array([0.5, 1.2]))


def test_precision_is_inverse_of_gls_covariance(design):
    # test/merit_tests.jl:561-567
    fd = design
    g = fd.gate_eigenvalues
    lam = fo.get_eigenvalues(fd, g)
    cov = fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, g), fo.calc_covariance_eigenvalues(fd, g))
    cl = fo.calc_covariance_log(cov, lam)
    prec = fo.calc_precision_matrix(fd.matrix, g, fo.sparse_covariance_inv(cl, fd.mapping_lengths)).toarray()
    gls = fo.calc_gls_covariance(fd, cl)
    assert approx(prec, np.linalg.inv(gls))
    cld = sp.diags(cl.diagonal()).tocsc()
    wls_diag = fo.calc_wls_covariance(fd, cld)
    prec_d = fo.calc_precision_matrix(fd.matrix, g, sp.diags(1 / cl.diagonal())).toarray()
    assert approx(prec_d, np.linalg.inv(wls_diag))
    # merit ordering GLS <= WLS <= OLS (test/design_tests.jl:175-188) on the NRMSE expectation
    e = [fo.nrmse_moments(f(fd, cl))[0] for f in (fo.calc_gls_covariance, fo.calc_wls_covariance, fo.calc_ols_covariance)]
    assert e[0] <= e[1] * (1 + 1e-12) and e[1] <= e[2] * (1 + 1e-12)


def test_fit_known_answers(design):
    fd = design
    A = fd.matrix
    # noiseless records => y == 1 => b == 0 => x == 0 => g == 1.0 exactly (test/device_tests.jl:144-155)
    ones = np.ones(A.shape[0])
    F = sp.identity(A.shape[0], format="csc") * 3.0
    assert np.all(fo.estimate_gate_eigenvalues(A, ones) == 1.0)
    assert np.all(fo.estimate_gate_eigenvalues(A, ones, F) == 1.0)
    # noiseless y = exp(A log g) recovers g (full column rank)
    y = fo.get_eigenvalues(fd, fd.gate_eigenvalues)
    assert np.allclose(fo.estimate_gate_eigenvalues(A, y), fd.gate_eigenvalues, rtol=1e-11)
    # GLS with a diagonal covariance == WLS; WLS with unit weights == OLS (test/design_tests.jl:243-246)
    eig, cov = noisy_ensembles(fd, seed=5)
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov), lam)
    cld = sp.diags(cl.diagonal()).tocsc()
    gls_d = fo.estimate_gate_eigenvalues(A, lam, fo.sparse_covariance_inv_factor(cld, fd.mapping_lengths))
    wls = fo.estimate_gate_eigenvalues(A, lam, sp.diags(cl.diagonal() ** -0.5))
    assert np.allclose(gls_d, wls, rtol=1e-11)
    assert np.allclose(fo.estimate_gate_eigenvalues(A, lam, sp.identity(A.shape[0])), fo.estimate_gate_eigenvalues(A, lam), rtol=1e-12)
    # constrain clamps negative log-eigenvalues (gate eigenvalues > 1) to exactly 1
    gc = fo.estimate_gate_eigenvalues(A, lam, constrain=True)
    assert np.all(gc <= 1.0)


def test_gate_noise_core_statistics(design):
    fd = design
    eig, cov = noisy_ensembles(fd, shots=10**9, seed=6)
    out = fo.estimate_gate_noise_core(fd, eig, cov)
    for k in ("ols", "wls", "gls"):
        err = np.linalg.norm(out[k] - fd.gate_eigenvalues) / np.sqrt(len(fd.gate_eigenvalues))
        assert err < 5e-4, (k, err)


# ---- committed golden vectors (tests/golden/fit_golden.npz, made by a literal dense transcription) ----

def _golden_designs():
    import hashlib

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_golden.npz"))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_fit_golden import design_hash

    for name, fd in (("example", aces_b200.example_design(full_covariance=True)),
                     ("rotated_d3", aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=3))):
        assert design_hash(fd) == str(z[f"{name}_hash"]), "the design generator changed: regenerate the fixture"
        yield name, fd, {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "_")}
    del hashlib


def _split(fd, lam, lam_cov):
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    coff = np.concatenate([[0], np.cumsum([len(k) for k in fd.cov_keys])])
    return ([lam[off[t]:off[t + 1]] for t in range(fd.tuple_number)],
            [lam_cov[coff[t]:coff[t + 1]] for t in range(fd.tuple_number)])


def test_oracle_matches_golden_fixture():
    for name, fd, g in _golden_designs():
        lam_ens, cov_ens = _split(fd, g["lam"], g["lam_cov"])
        cl = fo.calc_covariance_log(fo.calc_covariance(fd, lam_ens, cov_ens, weight_time=False), g["lam"])
        A = fd.matrix.astype(np.float64)
        got = {"ols": fo.estimate_gate_eigenvalues(A, g["lam"]),
               "wls": fo.estimate_gate_eigenvalues(A, g["lam"], sp.diags(cl.diagonal() ** -0.5)),
               "gls": fo.estimate_gate_eigenvalues(A, g["lam"], fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths))}
        for k in ("ols", "wls", "gls"):
            assert np.linalg.norm(got[k] - g[k]) <= 1e-9 * np.linalg.norm(g[k]), (name, k)
        ge = fd.gate_eigenvalues
        ct = fo.calc_covariance_log(fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, ge),
                                                       fo.calc_covariance_eigenvalues(fd, ge)), fo.get_eigenvalues(fd, ge))
        assert np.allclose(ct.diagonal(), g["true_covariance_log_diag"], rtol=1e-13)
        for k, f in (("gls", fo.calc_gls_covariance), ("wls", fo.calc_wls_covariance), ("ols", fo.calc_ols_covariance)):
            assert abs(np.trace(f(fd, ct)) - float(g[f"tr_{k}"])) <= 1e-9 * float(g[f"tr_{k}"]), (name, k)
