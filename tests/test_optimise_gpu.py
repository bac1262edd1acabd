"""GPU tests of SURVEY 8f row f3: aces_design_merit_grad (trace form, nothing larger than N x N) against
the dense transcription of the reference's gradients (oracle/optimise_oracle.py), and the Nesterov loop
of gls_ / wls_ / ols_optimise_weights driven by it."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import aces_b200
from fit_util import fo

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import optimise_oracle as oo  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-9   # relative, fp64 (north star)


def unweighted_log_covariance(fd):
    lam = fo.get_eigenvalues_ensemble(fd, fd.gate_eigenvalues)
    lam12 = fo.calc_covariance_eigenvalues(fd, fd.gate_eigenvalues)
    cov_log = fo.calc_covariance_log(fo.calc_covariance(fd, lam, lam12), np.concatenate(lam))
    sfi = oo.shot_weights_factor_inv(fd.shot_weights, fd.tuple_times, fd.mapping_lengths)
    cu = cov_log.toarray() * sfi[None, :]
    return sp.csc_matrix(cov_log), sp.csc_matrix((cu + cu.T) / 2)


@pytest.fixture(scope="module", params=["example", "real3", "synthetic3"])
def setup(request):
    if request.param == "example":
        fd = aces_b200.example_design(full_covariance=True)
    elif request.param == "real3":
        fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    else:   # no column windows: the design forces the windowed Gram path for the GLS blocks
        fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True, with_matrix=True)
    cov_log, cu = unweighted_log_covariance(fd)
    ctx = aces_b200.Context(0)
    dd = aces_b200.DeviceDesign(ctx, fd)
    # tr(Sigma^2) = <P2, Bm> is a sum with cancellation; for the synthetic look-alike's OLS estimator the
    # terms are 2.6e5 times larger than the sum (1.1e4 for the real design), which multiplies the 1e-11
    # relative error of the explicit inverse: the merit then agrees to 1e-8 instead of 1e-12.  The
    # reference's own formulas (diag of products of the same matrices) have the same conditioning.
    fd.merit_tol = 1e-7 if request.param == "synthetic3" else TOL
    yield fd, cov_log, cu, dd
    dd.close()
    ctx.close()


def oracle_grad(fd, ls_type, w, cu, Tm):
    if ls_type == "gls":
        return oo.calc_gls_merit_grad_log(fd, w, fo.sparse_covariance_inv(cu, fd.mapping_lengths), Tm)
    if ls_type == "wls":
        return oo.calc_wls_merit_grad_log(fd, w, cu, Tm)
    return oo.calc_ols_merit_grad_log(fd, w, *oo.ols_matrices(fd, cu, Tm))


def transforms_of(fd):
    g = sp.diags(np.asarray(fd.gate_eigenvalues, dtype=np.float64))
    out = {"ordinary": aces_b200.get_transform(fd, "ordinary") @ g}
    if fd.gate_types is not None:
        out["marginal"] = aces_b200.get_transform(fd, "marginal") @ g
        out["relative"] = aces_b200.get_transform(fd, "relative") @ g
    return out


@pytest.mark.parametrize("ls_type", ["gls", "wls", "ols"])
def test_merit_grad_vs_oracle(setup, ls_type):
    fd, _, cu, dd = setup
    rng = np.random.default_rng(7)
    w = fd.shot_weights * np.exp(0.4 * rng.standard_normal(fd.tuple_number))
    w /= w.sum()
    for name, Tm in transforms_of(fd).items():
        grad, merit = aces_b200.optimise.calc_merit_grad_log(dd, ls_type, w, cu, Tm)
        want_grad, want_merit = oracle_grad(fd, ls_type, w, cu, sp.csc_matrix(Tm))
        assert abs(merit - want_merit) <= fd.merit_tol * abs(want_merit), (name, merit, want_merit)
        assert np.linalg.norm(grad - want_grad) <= 1e-6 * np.linalg.norm(want_grad), (name, grad, want_grad)


def test_transforms_match_the_oracle(setup):
    fd = setup[0]
    if fd.gate_types is None:
        pytest.skip("synthetic design: no gate types")
    for est in ("ordinary", "marginal", "relative"):
        a = aces_b200.get_transform(fd, est)
        b = oo.get_transform(fd.gate_types, fd.gate_ptr, est)
        assert a.shape == b.shape and abs(a - b).max() == 0


def test_repeated_calls_are_bit_identical(setup):
    fd, _, cu, dd = setup
    Tm = transforms_of(fd)["ordinary"]
    w = np.asarray(fd.shot_weights, dtype=np.float64)
    for ls in ("gls", "wls", "ols"):
        g1, m1 = aces_b200.optimise.calc_merit_grad_log(dd, ls, w, cu, Tm)
        g2, m2 = aces_b200.optimise.calc_merit_grad_log(dd, ls, w, cu, Tm)
        assert m1 == m2 and np.array_equal(g1, g2)


@pytest.mark.parametrize("ls_type", ["gls", "wls", "ols"])
def test_optimise_weights_loop(setup, ls_type):
    """The reference's descent loop with the device gradient against the same loop with the oracle's
    gradient: the figure of merit goes down and both runs agree step by step."""
    fd, cov_log, _, dd = setup
    if fd.M > 200:
        pytest.skip("the dense oracle loop is kept to the small design")
    opt = aces_b200.OptimOptions(ls_type=ls_type, est_type="prod" if fd.gate_types is not None else "ordinary", max_steps=12)
    w_gpu, cov_gpu, descent_gpu = aces_b200.optimise_weights(dd, cov_log, opt)
    w_cpu, cov_cpu, descent_cpu = aces_b200.optimise_weights(
        fd, cov_log, opt, grad_fn=lambda ls, w, cu, Tm: oracle_grad(fd, ls, w, cu, sp.csc_matrix(Tm)))
    assert len(descent_gpu) == len(descent_cpu)
    assert np.allclose(descent_gpu, descent_cpu, rtol=1e-8, atol=0)
    assert np.allclose(w_gpu, w_cpu, rtol=1e-6, atol=1e-12)
    assert min(descent_gpu) < descent_gpu[0]
    assert abs(w_gpu.sum() - 1) < 1e-12 and np.all(w_gpu >= 0)
    assert abs(cov_gpu - cov_cpu).max() <= 1e-6 * abs(cov_cpu).max()
