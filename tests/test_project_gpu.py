"""GPU parity tests of the batched simplex projections against
oracle/project_oracle.py: the Euclidean projection (sort based, deterministic) is compared bit for bit;
the Mahalanobis projection to 1e-8, the reference's SCS tolerance (src/utils.jl:57-97)."""
import os
import sys

import numpy as np
import pytest

from fit_util import fo, noisy_ensembles

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import project_oracle as po  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,d,scale", [("rotated", 3, 1.0), ("rotated", 5, 30.0), ("unrotated", 3, 3.0), ("rotated", 9, 10.0)])
def test_simple_projection_bit_exact(ctx, kind, d, scale):
    import aces_b200

    fd = (aces_b200.rotated_design if kind == "rotated" else aces_b200.unrotated_design)(d, full_covariance=False,
                                                                                         noise="lognormal", seed=d)
    rng = np.random.default_rng(d)
    noisy = fd.gate_eigenvalues + rng.normal(0, scale * 1e-3, size=len(fd.gate_eigenvalues))
    assert set(np.diff(fd.gate_ptr)) <= {1, 3, 15}
    got = aces_b200.project.simple_project_gate_eigenvalues(ctx, noisy, fd.gate_ptr)
    want = po.simple_project_gate_eigenvalues(noisy, fd.gate_ptr)
    for g, w, name in zip(got, want, ("eigenvalues", "unprojected probabilities", "probabilities")):
        assert np.array_equal(g, w), name
    assert np.any(want[1] < 0)                      # the projection did something
    # the true eigenvalues are a fixed point (src/estimate.jl:744 relies on it)
    fix = aces_b200.project.simple_project_gate_eigenvalues(ctx, fd.gate_eigenvalues, fd.gate_ptr)
    assert np.allclose(fix[0], fd.gate_eigenvalues, atol=1e-15) and np.all(fix[2] >= 0)


def test_simple_projection_scattered_groups(ctx):
    import aces_b200

    rng = np.random.default_rng(4)
    N = 60
    perm = rng.permutation(N).astype(np.int32)
    gate_ptr = np.array([0, 15, 18, 19, 34, 37, 40, 55, 56, 59, 60])
    lam = 1 - rng.uniform(0, 0.05, size=N) + rng.normal(0, 0.02, size=N)
    got, up, pp = aces_b200.project.simple_project_gate_eigenvalues(ctx, lam, gate_ptr, perm)
    want, wup, wpp = po.simple_project_gate_eigenvalues(lam[perm], gate_ptr)
    assert np.array_equal(got[perm], want) and np.array_equal(up, wup) and np.array_equal(pp, wpp)
    with pytest.raises(aces_b200.AcesError):        # a gate of 5 eigenvalues has no Walsh-Hadamard block
        aces_b200.project.simple_project_gate_eigenvalues(ctx, lam[:5], np.array([0, 5]))


def test_nonnegative_projection_vs_oracle(ctx):
    import aces_b200

    rng = np.random.default_rng(6)
    vs, Qs = [], []
    for n in [1, 3, 15] * 200:
        B = rng.normal(size=(n + 2, n))
        Qs.append(B.T @ B + 0.05 * np.eye(n))
        vs.append(rng.normal(0.002, 0.01, size=n))
    ys = aces_b200.project.project_nonnegative_batched(ctx, vs, Qs, precision=1e-8)
    worst = 0.0
    for v, Q, y in zip(vs, Qs, ys):
        want = po.scs_project_nonnegative(v, Q, precision=1e-8)
        worst = max(worst, float(np.max(np.abs(y - want))))
        assert np.all(y >= 0)
    assert worst < 1e-8, worst
    # exactness (threshold 0): Karush-Kuhn-Tucker conditions to rounding
    ys = aces_b200.project.project_nonnegative_batched(ctx, vs, Qs, precision=0.0)
    for v, Q, y in zip(vs, Qs, ys):
        grad = Q @ (y - v)
        s = np.abs(Q @ v).max()
        assert np.all(grad >= -1e-11 * s) and abs(y @ grad) <= 1e-11 * s * max(np.abs(y).max(), 1e-300)
    with pytest.raises(aces_b200.NotPositiveDefinite):
        aces_b200.project.project_nonnegative_batched(ctx, [np.ones(2)], [np.array([[1.0, 2.0], [2.0, 1.0]])])
    assert aces_b200.project.project_nonnegative_batched(ctx, [], []) == []


@pytest.mark.parametrize("ls_type", ["wls", "gls"])
def test_split_projection_of_a_fit_vs_oracle(ctx, ls_type):
    """split_project_gate_eigenvalues on a real fit: precision matrix of the unprojected estimate
    (src/estimate.jl:745-764, src/merit.jl:754-770) from the oracle, projection on the GPU."""
    import scipy.sparse as sp

    import aces_b200

    fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    eig, cov = noisy_ensembles(fd, shots=10**6, seed=8)       # few shots: negative probabilities appear
    core = fo.estimate_gate_noise_core(fd, eig, cov)
    unproj = core[ls_type]
    cl = core["covariance_log"]
    A = sp.csr_matrix(fd.matrix)[core["clipped_indices"], :].astype(np.float64)
    if ls_type == "gls":
        inv = fo.sparse_covariance_inv(cl, fo.get_clipped_mapping_lengths(fd.mapping_lengths, core["clipped_indices"]))
    else:
        inv = sp.diags(1.0 / cl.diagonal())
    P = fo.calc_precision_matrix(A, unproj, inv).toarray()
    blocks = [P[a:b, a:b] for a, b in zip(fd.gate_ptr[:-1], fd.gate_ptr[1:])]
    got = aces_b200.project.split_project_gate_eigenvalues(ctx, unproj, blocks, fd.gate_ptr)
    want = po.split_project_gate_eigenvalues(unproj, P, fd.gate_ptr)
    assert np.any(want[1] < 0)
    assert np.max(np.abs(got[2] - want[2])) < 1e-8 and np.max(np.abs(got[0] - want[0])) < 1e-7
    assert np.all(got[2] >= 0)


@pytest.mark.parametrize("name,ls_type,shots", [("real3", "gls", 10**6), ("real3", "wls", 10**6), ("real3", "ols", 10**5),
                                                 ("example", "gls", 10**5), ("real5", "gls", 10**6)])
def test_full_projection_of_a_fit_vs_oracle(ctx, name, ls_type, shots):
    """full_project_gate_eigenvalues (src/estimate.jl:532-573), the reference's default below 5000 gate
    eigenvalues: the simultaneous Mahalanobis projection of all gates.  GPU: precision matrix from the fit's own
    Gram matrix, exact block principal pivoting; oracle: the reference's sparse precision matrix
    (calc_precision_matrix) and an exact NNLS solve of the same N-variable program."""
    import scipy.sparse as sp

    import aces_b200

    fd = (aces_b200.example_design(full_covariance=True) if name == "example"
          else aces_b200.rotated_design(int(name[-1]), full_covariance=True, noise="lognormal", seed=4))
    eig, cov = noisy_ensembles(fd, shots=shots, seed=8)       # few shots: negative probabilities appear
    core = fo.estimate_gate_noise_core(fd, eig, cov)
    unproj = core[ls_type]
    cl = core["covariance_log"]
    kept = core["clipped_indices"]
    A = sp.csr_matrix(fd.matrix)[kept, :].astype(np.float64)
    if ls_type == "gls":
        inv = fo.sparse_covariance_inv(cl, fo.get_clipped_mapping_lengths(fd.mapping_lengths, kept))
    elif ls_type == "wls":
        inv = sp.diags(1.0 / cl.diagonal())
    else:
        inv = sp.identity(A.shape[0])
    P = fo.calc_precision_matrix(A, unproj, inv).toarray()
    want = po.full_project_gate_eigenvalues(unproj, P, fd.gate_ptr)
    assert np.any(want[1] < 0)
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        est = np.concatenate(aces_b200.mean_ensemble(eig))
        covm = np.concatenate(aces_b200.mean_ensemble(cov))
        clipped = None if len(kept) == dd.M else kept
        un_gpu = dd.fit(ls_type, est, covm, clipped)
        assert np.linalg.norm(un_gpu - unproj) < 1e-9 * np.linalg.norm(unproj)
        got = aces_b200.project.full_project_gate_eigenvalues(dd, ls_type, un_gpu, fd.gate_ptr)
        its = aces_b200.project.full_project_gate_eigenvalues.last_iterations
        assert 1 <= its <= 60
        assert np.all(got[2] >= 0)
        assert np.max(np.abs(got[2] - want[2])) < 1e-8 and np.max(np.abs(got[0] - want[0])) < 1e-7
        # the joint solution is at least as close (Mahalanobis) as the per-gate one, and differs from it
        dist = lambda x: float((x - unproj) @ P @ (x - unproj))  # noqa: E731
        split = po.split_project_gate_eigenvalues(unproj, P, fd.gate_ptr)
        assert dist(got[0]) <= dist(split[0]) * (1 + 1e-9)
        # must follow a fit of the same estimator
        with pytest.raises(aces_b200.AcesError):
            aces_b200.project.full_project_gate_eigenvalues(dd, ls_type, un_gpu, fd.gate_ptr)
    finally:
        dd.close()


@pytest.mark.parametrize("name,shots,split", [("real3", 10**6, True), ("real5", 10**7, True), ("real3", 10**10, True),
                                               ("real3", 10**6, False), ("real5", 10**7, None)])
def test_estimate_gate_noise_projected_vs_oracle(ctx, name, shots, split):
    """The whole numeric path of estimate_gate_noise (src/estimate.jl:703-831): fits, precision matrix from the
    design-resident Gram matrix, projections -- against the oracles; split = true (per gate), false (all gates
    at once) and the reference's default (None: false below 5000 gate eigenvalues)."""
    import scipy.sparse as sp

    import aces_b200

    fd = aces_b200.rotated_design(int(name[-1]), full_covariance=True, noise="lognormal", seed=2)
    eig, cov = noisy_ensembles(fd, shots=shots, seed=5)
    dd = aces_b200.DeviceDesign(ctx, fd)
    try:
        got = aces_b200.estimate_gate_noise_projected(dd, eig, cov, clip_warning=False, split=split)
        core = fo.estimate_gate_noise_core(fd, eig, cov)
        ols = po.simple_project_gate_eigenvalues(core["ols"], fd.gate_ptr)
        assert np.max(np.abs(got["ols"] - ols[0])) < 1e-9
        pproj = not aces_b200.merit.isapprox(ols[0], core["ols"])
        assert got["precision_project"] == pproj
        assert pproj == (shots < 10**10)          # shot noise pushes probabilities negative, 1e10 shots do not
        A = sp.csr_matrix(fd.matrix)[core["clipped_indices"], :].astype(np.float64)
        cl = core["covariance_log"]
        for k in ("wls", "gls"):
            assert np.linalg.norm(got[k + "_unproj"] - core[k]) < 1e-9 * np.linalg.norm(core[k])
            if pproj:
                inv = (fo.sparse_covariance_inv(cl, fo.get_clipped_mapping_lengths(fd.mapping_lengths, core["clipped_indices"]))
                       if k == "gls" else sp.diags(1.0 / cl.diagonal()))
                P = fo.calc_precision_matrix(A, core[k], inv)
                # the precision-matrix slices themselves: 1e-9 (src/merit.jl:754-770)
                blocks = dd.precision_blocks(k, fd.gate_ptr, got[k + "_unproj"]) if (k == "gls" and split) else None
                if blocks is not None:
                    for (a, b), blk in zip(zip(fd.gate_ptr[:-1], fd.gate_ptr[1:]), blocks):
                        w = P[a:b, a:b].toarray()
                        assert np.linalg.norm(blk - w) <= 1e-9 * np.linalg.norm(w)
                want = (po.split_project_gate_eigenvalues if split else po.full_project_gate_eigenvalues)(core[k], P, fd.gate_ptr)
            else:
                want = po.simple_project_gate_eigenvalues(core[k], fd.gate_ptr)
            assert np.max(np.abs(got[k + "_probabilities"] - want[2])) < 2e-8, k
            assert np.max(np.abs(got[k] - want[0])) < 2e-7, k
        with pytest.raises(aces_b200.AcesError):   # the slices belong to the LAST fit's estimator
            dd.precision_blocks("ols", fd.gate_ptr)
        if not split:
            assert fd.matrix.shape[1] < 5000     # the default took the simultaneous projection
    finally:
        dd.close()
