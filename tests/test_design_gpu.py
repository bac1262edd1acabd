"""GPU parity tests of the design-resident fit path (aces_design_*): covariance assembly (K2), block
inverse factor (K3), fused OLS / WLS / GLS fit (K4/K5) and estimator covariances (K6) against the
numpy/scipy oracle.  Floating point: relative tolerance 1e-9 in fp64 (BASELINE.json north_star);
the covariance values are element-wise products and are compared to 1e-15."""
import numpy as np
import pytest
import scipy.sparse as sp

from fit_util import fo, noisy_ensembles

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def rel(x, y):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    return np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300)


def dense(m):
    return m.toarray() if sp.issparse(m) else np.asarray(m)


@pytest.fixture(scope="module", params=[(3, True), (5, True), (3, False), ("real3", True), ("real5", True), ("unrot3", True)],
                ids=["d3_gls", "d5_gls", "d3_diag", "real_d3_gls", "real_d5_gls", "real_unrotated_d3_gls"])
def setup(request, ctx):
    import aces_b200

    d, full = request.param
    if isinstance(d, str):   # real designs (circuit_design.py): the reference's own A, keys and counts
        kind, d = d.rstrip("0123456789"), int(d[-1])
        fd = (aces_b200.rotated_design if kind == "real" else aces_b200.unrotated_design)(d, full_covariance=full)
    else:
        fd = aces_b200.synthetic_design("rotated", d, full_covariance=full, with_matrix=True)
    dd = aces_b200.DeviceDesign(ctx, fd)
    eig, cov = noisy_ensembles(fd, seed=7 + d)
    yield fd, dd, eig, cov
    dd.close()


def test_pattern(setup):
    fd, dd, _, _ = setup
    n_keys = sum(len(k) for k in fd.cov_keys) if fd.full_covariance else 0
    assert len(dd.cov_rowval) == fd.M + 2 * n_keys
    pat = dd._sparse(np.ones(len(dd.cov_rowval)))
    assert (pat - pat.T).count_nonzero() == 0
    assert np.all(pat.diagonal() == 1.0)


@pytest.mark.parametrize("weight_time", [True, False])
def test_calc_covariance_vs_oracle(setup, weight_time):
    import aces_b200

    fd, dd, eig, cov = setup
    lam_ens = aces_b200.mean_ensemble(eig)
    cov_ens = aces_b200.mean_ensemble(cov) if fd.full_covariance else None
    got = dd.calc_covariance(lam_ens, cov_ens, weight_time=weight_time)
    want = fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=weight_time)
    assert np.max(np.abs(dense(got) - dense(want))) <= 1e-15 * np.max(np.abs(dense(want)))
    lam = np.concatenate(lam_ens)
    got_log = dd.calc_covariance(lam_ens, cov_ens, weight_time=weight_time, log_form=True)
    want_log = fo.calc_covariance_log(want, lam)
    assert rel(dense(got_log), dense(want_log)) < 1e-15
    assert (got_log - got_log.T).count_nonzero() == 0
    # the epsilon floor at eigenvalues of exactly +-1 (src/merit.jl:253)
    ones = [np.ones(int(m)) for m in fd.mapping_lengths]
    ones[0][::2] = -1.0
    c1 = [np.ones(len(k)) for k in fd.cov_keys] if fd.full_covariance else None
    got1 = dd.calc_covariance(ones, c1, epsilon=1e-10, weight_time=False)
    want1 = fo.calc_covariance(fd, ones, c1 if c1 is not None else [[] for _ in ones], weight_time=False)
    assert np.array_equal(got1.diagonal(), want1.diagonal())
    assert np.all(got1.diagonal() > 0)


def test_inverse_factor_identities(setup):
    fd, dd, eig, cov = setup
    if not fd.full_covariance:
        pytest.skip("diagonal design")
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
    F = dd.sparse_covariance_inv_factor(cl)
    Fo = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    # test/merit_tests.jl:531-543: F'F == inv(cholesky(covariance_log))
    assert rel(dense(F.T @ F), dense(Fo.T @ Fo)) < RTOL
    assert rel(dense(F @ cl @ F.T), np.eye(fd.M)) < RTOL
    assert sp.tril(F, -1).count_nonzero() > 0 and sp.triu(F, 1).count_nonzero() == 0
    # after clipping (test/merit_tests.jl:569-591)
    rng = np.random.default_rng(3)
    clipped = np.sort(rng.choice(fd.M, size=fd.M - 37, replace=False))
    Fc = dd.sparse_covariance_inv_factor(cl, clipped)
    ccl = sp.csc_matrix(cl)[clipped, :][:, clipped]
    assert Fc.shape == (len(clipped), len(clipped))
    assert rel(dense(Fc @ ccl @ Fc.T), np.eye(len(clipped))) < RTOL


def test_fused_fit_vs_oracle(setup):
    import aces_b200

    fd, dd, eig, cov = setup
    want = fo.estimate_gate_noise_core(fd, eig, cov)
    got = aces_b200.estimate_gate_noise_core(dd, eig, cov, clip_warning=False)
    assert np.array_equal(got["est_eigenvalues"], want["est_eigenvalues"])
    assert np.array_equal(got["clipped_indices"], want["clipped_indices"])
    for k in ("ols", "wls", "gls"):
        assert rel(got[k], want[k]) < RTOL, k
        assert rel(np.log(got[k]), np.log(want[k])) < 1e-7, k


def test_fused_fit_after_clipping(setup):
    import aces_b200

    fd, dd, eig, cov = setup
    eig = [[list(v) for v in t] for t in eig]
    rng = np.random.default_rng(5)
    # clip rows with several entries only: the single-entry rows carry the column rank
    multi = np.nonzero(np.diff(sp.csr_matrix(fd.matrix).indptr) >= 2)[0]
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    for r in rng.choice(multi, size=25, replace=False):
        t = int(np.searchsorted(off, r, side="right") - 1)
        p = int(r - off[t])
        eig[t][p] = [0.05 for _ in eig[t][p]]  # below the 0.1 threshold
    want = fo.estimate_gate_noise_core(fd, eig, cov)
    got = aces_b200.estimate_gate_noise_core(dd, eig, cov, clip_warning=False)
    assert len(got["clipped_indices"]) == fd.M - 25
    assert np.array_equal(got["clipped_indices"], want["clipped_indices"])
    for k in ("ols", "wls", "gls"):
        assert rel(got[k], want[k]) < RTOL, k


def test_fit_known_answers(setup):
    fd, dd, _, _ = setup
    ones = np.ones(fd.M)
    c1 = np.ones(dd.C) if dd.C else None
    # noiseless records => every gate eigenvalue exactly 1.0 (test/device_tests.jl:144-155)
    for k in ("ols", "wls", "gls"):
        assert np.all(dd.fit(k, ones, c1) == 1.0), k
    # exact recovery from noiseless circuit eigenvalues
    lam = fo.get_eigenvalues(fd, fd.gate_eigenvalues)
    lc = np.concatenate(fo.calc_covariance_eigenvalues(fd, fd.gate_eigenvalues)) if dd.C else None
    for k in ("ols", "wls", "gls"):
        assert np.max(np.abs(dd.fit(k, lam, lc) - fd.gate_eigenvalues)) < 1e-12, k


def test_ls_covariances_vs_oracle(setup):
    import aces_b200

    fd, dd, _, _ = setup
    g = fd.gate_eigenvalues
    lam = fo.get_eigenvalues(fd, g)
    cov_eig = fo.calc_covariance_eigenvalues(fd, g) if fd.full_covariance else [[] for _ in range(fd.tuple_number)]
    cl = fo.calc_covariance_log(fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, g), cov_eig), lam)
    for k, ref in (("gls", fo.calc_gls_covariance), ("wls", fo.calc_wls_covariance), ("ols", fo.calc_ols_covariance)):
        want = ref(fd, cl)
        sigma, tr, trsq = dd.calc_ls_covariance(k, cl)
        assert rel(sigma, want) < RTOL, k
        assert np.array_equal(sigma, sigma.T), k
        assert abs(tr - np.trace(want)) < RTOL * np.trace(want), k
        assert abs(trsq - np.sum(want * want.T)) < RTOL * np.sum(want * want.T), k
        e_got = aces_b200.nrmse_moments(tr, trsq, fd.matrix.shape[1])
        e_want = fo.nrmse_moments(want)
        assert np.allclose(e_got, e_want, rtol=1e-9), k
        _, tr2, trsq2 = dd.calc_ls_covariance(k, cl, want_matrix=False)
        assert tr2 == tr and trsq2 == trsq


def test_not_positive_definite_block_is_reported(setup):
    import aces_b200

    fd, dd, eig, cov = setup
    if not fd.full_covariance:
        pytest.skip("diagonal design")
    lam = np.concatenate(fo.mean_ensemble(eig))
    cl = sp.lil_matrix(fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam))
    t0 = next(t for t in range(fd.tuple_number) if len(fd.cov_keys[t]))   # first tuple with off-diagonal keys
    r0 = int(np.sum(fd.mapping_lengths[:t0]))
    p, q = (r0 + int(x) for x in fd.cov_keys[t0][0])
    big = 10.0 * max(cl[p, p], cl[q, q])
    cl[p, q] = big
    cl[q, p] = big
    cl = sp.csc_matrix(cl)
    with pytest.raises(aces_b200.NotPositiveDefinite):
        dd.sparse_covariance_inv_factor(cl, fallback=False)
    # the reference's fallback (src/merit.jl:403-424): |block|, smallest eigenvalue raised to epsilon
    eps = 1e-6
    F = dd.sparse_covariance_inv_factor(cl, epsilon=eps)
    assert dd.fallback_count() == 1
    Fo = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths, epsilon=eps)
    r1 = r0 + int(fd.mapping_lengths[t0])
    inv_got, inv_want = dense(F.T @ F), dense(Fo.T @ Fo)
    # the shifted block has condition number ~ 1e7: compare the inverses loosely, the other blocks tightly
    rest = np.r_[0:r0, r1:inv_got.shape[0]]
    assert rel(inv_got[np.ix_(rest, rest)], inv_want[np.ix_(rest, rest)]) < RTOL
    assert rel(inv_got[r0:r1, r0:r1], inv_want[r0:r1, r0:r1]) < 1e-4
    shifted = np.abs(dense(cl)[r0:r1, r0:r1])
    w = np.linalg.eigvalsh(np.linalg.inv(inv_got[r0:r1, r0:r1]))
    assert abs(w[0] - eps) < 0.05 * eps, (w[0], np.linalg.norm(shifted, np.inf))


def test_design_create_rejects_bad_input(ctx):
    import aces_b200

    fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True, with_matrix=True)
    fd.cov_keys[0] = np.vstack([fd.cov_keys[0], fd.cov_keys[0][:1]])  # duplicate key
    fd.cov_counts[0] = np.concatenate([fd.cov_counts[0], fd.cov_counts[0][:1]])
    with pytest.raises(aces_b200.AcesError):
        aces_b200.DeviceDesign(ctx, fd)


@pytest.mark.parametrize("mode", ["zero_column", "random_70pct"])
def test_rank_deficient_fit_gives_basic_solution(ctx, mode):
    """Clipping (src/estimate.jl:435-460) can remove every row that determines a gate eigenvalue.  The
    reference's sparse-QR solve (src/estimate.jl:500-502) then returns a basic solution; so does the
    fit here (dependent columns -> x = 0, i.e. gate eigenvalue 1) instead of failing.  Which columns a
    basic solution zeroes depends on the elimination order, so the comparable quantities are the rank
    and the fitted circuit eigenvalues A x -- checked against LAPACK's rank-revealing gelsy."""
    import scipy.linalg as sla

    import aces_b200

    fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    A = sp.csr_matrix(fd.matrix).astype(np.float64)
    M, N = A.shape
    eig, cov = noisy_ensembles(fd, seed=3)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    if mode == "zero_column":
        rows = np.unique(sp.csc_matrix(A)[:, 100].nonzero()[0])
        kept = np.setdiff1d(np.arange(M), rows)
    else:
        kept = np.sort(np.random.default_rng(5).choice(M, size=int(0.7 * M), replace=False))
    Ak = A[kept].toarray()
    rank = int(np.linalg.matrix_rank(Ak))
    assert rank < N
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
        clk = sp.csc_matrix(cl)[kept, :][:, kept]
        factors = {"ols": None, "wls": sp.diags(clk.diagonal() ** (-0.5)),
                   "gls": fo.sparse_covariance_inv_factor(clk, fo.get_clipped_mapping_lengths(fd.mapping_lengths, kept))}
        b = -np.log(lam[kept])
        for k in ("ols", "wls", "gls"):
            g = dd.fit(k, lam, lc, clipped_indices=kept)
            assert ctx.last_rank_deficiency() == N - rank, (k, ctx.last_rank_deficiency(), N - rank)
            x = -np.log(g)
            assert int(np.sum(x == 0.0)) >= N - rank          # the eliminated columns sit at g = 1 exactly
            F = factors[k]
            FA, Fb = (Ak, b) if F is None else (F @ Ak, F @ b)
            xo, *_ = sla.lstsq(FA, Fb, lapack_driver="gelsy", cond=1e-12)
            # the whitened fitted values are unique even though x is not
            assert rel(FA @ x, FA @ xo) < 1e-8, (k, rel(FA @ x, FA @ xo))
        if mode == "zero_column":
            assert dd.fit("gls", lam, lc, clipped_indices=kept)[100] == 1.0
        # a full-rank fit on the same context reports zero again
        dd.fit("wls", lam, lc)
        assert ctx.last_rank_deficiency() == 0
    finally:
        dd.close()


@pytest.fixture(scope="module", params=["real9", "unrot7"], ids=["real_rotated_d9_gls", "real_unrotated_d7_gls"])
def setup_headline(request, ctx):
    """The headline shapes of BASELINE.json (configs[2]/[3]: rotated d=9, M x N = 12912 x 6779; unrotated
    d=7, 11700 x 6189) on the reference's own design tables, with log-normal noise and shot-noise-level
    estimates.  The oracle's dense QR takes tens of seconds per estimator at this size."""
    import aces_b200

    if request.param == "real9":
        fd = aces_b200.rotated_design(9, full_covariance=True, noise="lognormal", seed=9)
        assert fd.matrix.shape == (12912, 6779)
    else:
        fd = aces_b200.unrotated_design(7, full_covariance=True, noise="lognormal", seed=7)
        assert fd.matrix.shape == (11700, 6189)
    dd = aces_b200.DeviceDesign(ctx, fd)
    eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
    yield fd, dd, eig, cov
    dd.close()


@pytest.mark.slow
def test_headline_fit_vs_oracle(setup_headline):
    """North star: OLS / WLS / GLS gate eigenvalues within 1e-9 (relative, fp64) of the reference's
    estimate_gate_noise core (src/estimate.jl:492-524, 711-743) at rotated d=9 and unrotated d=7 --
    noisy estimates, so the conditioning of the whitened system is what a real run sees."""
    import aces_b200

    fd, dd, eig, cov = setup_headline
    want = fo.estimate_gate_noise_core(fd, eig, cov)
    got = aces_b200.estimate_gate_noise_core(dd, eig, cov, clip_warning=False)
    assert np.array_equal(got["est_eigenvalues"], want["est_eigenvalues"])
    assert np.array_equal(got["clipped_indices"], want["clipped_indices"])
    assert dd.fallback_count() == 0
    for k in ("ols", "wls", "gls"):
        assert rel(got[k], want[k]) < RTOL, (k, rel(got[k], want[k]))
        assert np.max(np.abs(got[k] - want[k]) / np.abs(want[k])) < RTOL, k   # element-wise too
    # the estimators differ from each other by far more than the tolerance (the test is not vacuous)
    assert rel(got["gls"], got["ols"]) > 1e3 * RTOL


@pytest.mark.slow
def test_headline_fit_after_clipping_vs_oracle(setup_headline):
    """Same with clipped rows (src/estimate.jl:435-484, 727-735): 40 circuit eigenvalues below 0.1."""
    import aces_b200

    fd, dd, eig, cov = setup_headline
    eig = [[list(v) for v in t] for t in eig]
    rng = np.random.default_rng(5)
    multi = np.nonzero(np.diff(sp.csr_matrix(fd.matrix).indptr) >= 2)[0]
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    for r in rng.choice(multi, size=40, replace=False):
        t = int(np.searchsorted(off, r, side="right") - 1)
        eig[t][int(r - off[t])] = [0.05 for _ in eig[t][int(r - off[t])]]
    want = fo.estimate_gate_noise_core(fd, eig, cov)
    got = aces_b200.estimate_gate_noise_core(dd, eig, cov, clip_warning=False)
    assert len(got["clipped_indices"]) == fd.M - 40
    assert np.array_equal(got["clipped_indices"], want["clipped_indices"])
    for k in ("ols", "wls", "gls"):
        assert rel(got[k], want[k]) < RTOL, (k, rel(got[k], want[k]))


@pytest.mark.slow
def test_headline_covariances_vs_oracle(setup_headline):
    """K2 / K3 / K6 at the headline shapes: covariance values (1e-15), F'F against the oracle's block
    inverse (1e-9), GLS / WLS / OLS estimator covariances and their trace moments (1e-9)."""
    import aces_b200

    fd, dd, eig, cov = setup_headline
    lam_ens, cov_ens = aces_b200.mean_ensemble(eig), aces_b200.mean_ensemble(cov)
    got = dd.calc_covariance(lam_ens, cov_ens, weight_time=False)
    want = fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False)
    diff = sp.csc_matrix(got) - sp.csc_matrix(want)
    assert (abs(diff).max() if diff.nnz else 0.0) <= 1e-15 * abs(want).max()
    lam = np.concatenate(lam_ens)
    cl = fo.calc_covariance_log(want, lam)
    got_log = dd.calc_covariance(lam_ens, cov_ens, weight_time=False, log_form=True)
    dl = sp.csc_matrix(got_log) - cl
    assert np.sqrt(dl.multiply(dl).sum()) <= 1e-15 * np.sqrt(cl.multiply(cl).sum())
    F = dd.sparse_covariance_inv_factor(cl)
    Fo = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    P, Po = sp.csc_matrix(F.T @ F), sp.csc_matrix(Fo.T @ Fo)
    dP = P - Po
    assert np.sqrt(dP.multiply(dP).sum()) < RTOL * np.sqrt(Po.multiply(Po).sum())
    # estimator covariances of the true gate eigenvalues (src/merit.jl:556-629)
    g = fd.gate_eigenvalues
    cov_eig = fo.calc_covariance_eigenvalues(fd, g)
    clt = fo.calc_covariance_log(fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, g), cov_eig), fo.get_eigenvalues(fd, g))
    for k, ref in (("gls", fo.calc_gls_covariance), ("wls", fo.calc_wls_covariance), ("ols", fo.calc_ols_covariance)):
        w = ref(fd, clt)
        sigma, tr, trsq = dd.calc_ls_covariance(k, clt)
        assert rel(sigma, w) < RTOL, (k, rel(sigma, w))
        assert abs(tr - np.trace(w)) < RTOL * np.trace(w), k
        assert abs(trsq - np.sum(w * w.T)) < RTOL * np.sum(w * w.T), k
        del sigma, w


@pytest.mark.parametrize("kind,d", [("rotated", 9), ("unrotated", 7), ("real_rotated", 9), ("real_unrotated", 7)])
def test_full_size_fit_properties(ctx, kind, d):
    """BASELINE configs 3 / 4 shapes (rotated d=9: M = 12912, N = 6779; unrotated d=7): properties that
    do not need the oracle at this size."""
    import aces_b200

    if kind == "real_rotated":
        fd = aces_b200.rotated_design(d, full_covariance=True)
        assert fd.matrix.shape == (12912, 6779)
    elif kind == "real_unrotated":
        fd = aces_b200.unrotated_design(d, full_covariance=True)
    else:
        fd = aces_b200.synthetic_design(kind, d, full_covariance=True, with_matrix=True)
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        N = fd.matrix.shape[1]
        # (1) noiseless records: every gate eigenvalue exactly 1.0 (test/device_tests.jl:144-155)
        ones, c1 = np.ones(fd.M), np.ones(dd.C)
        for k in ("ols", "wls", "gls"):
            assert np.all(dd.fit(k, ones, c1) == 1.0), k
        # (2) exact recovery from noiseless circuit eigenvalues
        lam = np.exp(fd.matrix.astype(np.float64) @ np.log(fd.gate_eigenvalues))
        lc = np.concatenate([np.exp(r.astype(np.float64) @ np.log(fd.gate_eigenvalues)) for r in fd.cov_rows])
        for k in ("ols", "wls", "gls"):
            assert np.max(np.abs(dd.fit(k, lam, lc) - fd.gate_eigenvalues)) < 1e-11, k
        # (3) noisy estimates: the three estimators agree with the truth to shot noise and with each
        #     other to a few standard errors; GLS == WLS when the off-diagonal covariance vanishes
        #     (test/design_tests.jl:243-246), here by making lambda_pq = lambda_p * lambda_q exactly
        eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
        est = np.concatenate(aces_b200.mean_ensemble(eig))
        off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
        cprod = np.concatenate([est[off[t] + fd.cov_keys[t][:, 0]] * est[off[t] + fd.cov_keys[t][:, 1]]
                                for t in range(fd.tuple_number)])
        wls = dd.fit("wls", est, cprod)
        gls = dd.fit("gls", est, cprod)
        assert rel(gls, wls) < RTOL
        assert np.max(np.abs(wls - fd.gate_eigenvalues)) < 1e-3
        # (4) covariance moments: GLS is at least as good as WLS, which beats OLS (Gauss-Markov)
        lam_ens = [lam[off[t]:off[t + 1]] for t in range(fd.tuple_number)]
        cov_ens = [lc[dd._keep["ptr"][t]:dd._keep["ptr"][t + 1]] for t in range(fd.tuple_number)]
        cl = dd.calc_covariance(lam_ens, cov_ens, log_form=True)
        tr = {k: dd.calc_ls_covariance(k, cl, want_matrix=False)[1] for k in ("gls", "wls", "ols")}
        assert 0 < tr["gls"] <= tr["wls"] * (1 + 1e-9) <= tr["ols"] * (1 + 1e-9)
        assert N == len(wls)
    finally:
        dd.close()


def test_phase_timings_are_reported(ctx):
    import aces_b200

    fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True, with_matrix=True)
    dd = aces_b200.DeviceDesign(ctx, fd)
    eig, cov = noisy_ensembles(fd, seed=11)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    ctx.set_timing(True)
    dd.fit("gls", lam, lc)
    ph = dd.phase_ms()
    ctx.set_timing(False)
    assert {"covariance", "block factor", "gram", "potrf", "first solve", "refinement"} <= set(ph)
    assert all(v >= 0.0 for v in ph.values()) and dd.gram_flops() > 0
    dd.close()


@pytest.mark.parametrize("shards,ls_type", [(2, "gls"), (3, "gls"), (2, "wls"), (2, "ols")])
def test_tuple_sharded_fit_equals_unsharded(ctx, shards, ls_type):
    """SURVEY 8e "one large Gram": the tuple blocks of one fit split over shards, the partial Gram
    matrix and right-hand sides summed at the reduction points (here: shards are contexts on one GPU
    and the sum is done by torch; across GPUs it is shard.all_reduce_device = ncclAllReduce)."""
    import threading

    import torch

    import aces_b200
    from aces_b200 import shard

    fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    eig, cov = noisy_ensembles(fd, seed=21)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    # clip rows of the repeated tuples only: a real design stays full rank without them
    first_repeated = int(np.sum(fd.mapping_lengths[:9]))
    rows = np.arange(fd.M)
    kept = aces_b200.get_clipped_indices(np.where((rows % 17 == 3) & (rows >= first_repeated), 0.05, lam), warning=False)
    assert len(kept) < fd.M
    dd0 = aces_b200.DeviceDesign(ctx, fd)
    want = dd0.fit(ls_type, lam, lc)
    want_clipped = dd0.fit(ls_type, lam, lc, clipped_indices=kept)
    dd0.close()
    parts = shard.shard_tuples(fd.mapping_lengths, shards)
    assert sorted(t for p in parts for t in p) == list(range(fd.tuple_number))
    ctxs = [aces_b200.Context(0) for _ in range(shards)]
    dds = [aces_b200.DeviceDesign(c, fd) for c in ctxs]
    for dd, mine in zip(dds, parts):
        dd.set_tuple_shard([1 if t in mine else 0 for t in range(fd.tuple_number)])
    barrier = threading.Barrier(shards)
    slots = [None] * shards

    def reducer(i):
        def reduce(ptr, n):
            slots[i] = torch.as_tensor(shard._DeviceBuffer(ptr, n), device="cuda")
            barrier.wait()
            if i == 0:
                torch.cuda.synchronize()
                total = torch.stack(slots).sum(dim=0)
                for s in slots:
                    s.copy_(total)
                torch.cuda.synchronize()
            barrier.wait()
        return reduce

    for clipped, ref in ((None, want), (kept, want_clipped)):
        outs, errs = [None] * shards, []

        def work(i):
            try:
                outs[i] = dds[i].fit_sharded(ls_type, lam, lc, reducer(i), clipped_indices=clipped)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
                barrier.abort()

        ths = [threading.Thread(target=work, args=(i,)) for i in range(shards)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert not errs, errs
        for o in outs:
            assert np.array_equal(o, outs[0])            # every shard ends with the same answer
        assert rel(outs[0], ref) < 1e-12                 # and it is the unsharded fit (summation order differs)
    # a sharded design refuses the unsharded entry points; clearing the shard restores them
    with pytest.raises(aces_b200.AcesError):
        dds[0].fit(ls_type, lam, lc)
    dds[0].set_tuple_shard(None)
    assert rel(dds[0].fit(ls_type, lam, lc), want) < 1e-15
    for dd in dds:
        dd.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("shards,panel,ls_type", [(2, 128, "gls"), (3, 256, "gls"), (4, 128, "wls"), (2, 512, "gls")])
def test_tuple_sharded_fit_with_distributed_cholesky(ctx, shards, panel, ls_type):
    """SURVEY 8e, last row: the Cholesky factorisation of the summed Gram matrix distributed over the shards
    by block columns (aces_design_fit_dist_*): the owner factorises a block column, its full columns and
    block inverses are broadcast, every shard updates the block columns it owns.  Shards are contexts on one
    GPU here and the broadcast is a device copy; across GPUs it is shard.broadcast_device = ncclBroadcast.
    Every shard must end with the same factor (bit for bit) and the fit must equal the unsharded one."""
    import threading

    import torch

    import aces_b200
    from aces_b200 import shard

    fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    eig, cov = noisy_ensembles(fd, seed=21)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    dd0 = aces_b200.DeviceDesign(ctx, fd)
    want = dd0.fit(ls_type, lam, lc)
    dd0.close()
    parts = shard.shard_tuples(fd.mapping_lengths, shards)
    ctxs = [aces_b200.Context(0) for _ in range(shards)]
    dds = [aces_b200.DeviceDesign(c, fd) for c in ctxs]
    for dd, mine in zip(dds, parts):
        dd.set_tuple_shard([1 if t in mine else 0 for t in range(fd.tuple_number)])
        dd.DIST_PANEL = panel
    barrier = threading.Barrier(shards)
    slots = [None] * shards

    def reducer(i):
        def reduce(ptr, n):
            slots[i] = torch.as_tensor(shard._DeviceBuffer(ptr, n), device="cuda")
            barrier.wait()
            if i == 0:
                torch.cuda.synchronize()
                total = torch.stack(slots).sum(dim=0)
                for s in slots:
                    s.copy_(total)
                torch.cuda.synchronize()
            barrier.wait()
        return reduce

    def broadcaster(i):
        def broadcast(ptr, n, src):
            dds[i].ctx.synchronize()                     # the owner's factorisation / this shard's updates are done
            slots[i] = torch.as_tensor(shard._DeviceBuffer(ptr, n), device="cuda")
            barrier.wait()
            if i != src:
                slots[i].copy_(slots[src])
                torch.cuda.synchronize()
            barrier.wait()
        return broadcast

    outs, grams, errs = [None] * shards, [None] * shards, []

    def work(i):
        try:
            outs[i] = dds[i].fit_sharded(ls_type, lam, lc, reducer(i), distributed=(i, shards, broadcaster(i)))
            (g, n), _ = dds[i].fit_buffers()
            dds[i].ctx.synchronize()
            grams[i] = torch.as_tensor(shard._DeviceBuffer(g, n), device="cuda").clone()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            barrier.abort()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(shards)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    Np = int(round(np.sqrt(grams[0].numel())))
    low = torch.tril(torch.ones(Np, Np, dtype=torch.bool, device="cuda")).T.reshape(-1)   # column-major lower triangle
    for i in range(1, shards):
        assert torch.equal(grams[i][low], grams[0][low])     # the same factor on every shard
        assert np.array_equal(outs[i], outs[0])
    assert rel(outs[0], want) < 1e-12
    for dd in dds:
        dd.close()
    for c in ctxs:
        c.close()


def test_fit_matches_golden_fixture(ctx):
    """The committed known answers (tests/golden/fit_golden.npz: a literal dense transcription of the
    reference's formulas on the real example and rotated d=3 designs) through the CUDA path."""
    import aces_b200
    from test_fit_oracle import _golden_designs, _split

    for name, fd, g in _golden_designs():
        dd = aces_b200.DeviceDesign(ctx, fd)
        try:
            for k in ("ols", "wls", "gls"):
                got = dd.fit(k, g["lam"], g["lam_cov"])
                assert rel(got, g[k]) < RTOL, (name, k)
            ge = fd.gate_eigenvalues
            lam_true = np.exp(fd.matrix.astype(np.float64) @ np.log(ge))
            lc_true = np.concatenate([np.exp(r.astype(np.float64) @ np.log(ge)) for r in fd.cov_rows])
            lam_ens, cov_ens = _split(fd, lam_true, lc_true)
            ct = dd.calc_covariance(lam_ens, cov_ens, log_form=True)
            assert np.allclose(ct.diagonal(), g["true_covariance_log_diag"], rtol=1e-13)
            for k in ("gls", "wls", "ols"):
                _, tr, _ = dd.calc_ls_covariance(k, ct, want_matrix=False)
                assert abs(tr - float(g[f"tr_{k}"])) <= RTOL * float(g[f"tr_{k}"]), (name, k)
        finally:
            dd.close()
