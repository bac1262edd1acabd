"""CPU tests of the projection oracle (oracle/project_oracle.py) against the reference's own identities
(SURVEY.md section 8f row f1): project_simplex lands on the simplex and is idempotent
(src/utils.jl:47-48 asserts), a distribution already in the simplex is a fixed point of
simple_project_gate_eigenvalues (which is how estimate_gate_noise detects that no Mahalanobis
projection is needed, src/estimate.jl:744), the Walsh-Hadamard matrices are involutions up to 4^n, and
the Mahalanobis projection satisfies the Karush-Kuhn-Tucker conditions of its quadratic program."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import project_oracle as po  # noqa: E402


def test_wht_matrices():
    for k in (1, 3, 15):
        W, Wi = po.get_wht_matrix(k), po.get_wht_matrix(k, inverse=True)
        assert np.array_equal(W @ Wi, np.eye(k + 1))
        assert np.array_equal(W, W.T) and np.all(W[0] == 1.0)
    # one qubit: index a = x + 2z; X (1) and Z (2) anticommute, X and Y (3) anticommute
    W = po.get_wht_matrix(3)
    assert W[1, 2] == -1 and W[1, 3] == -1 and W[1, 1] == 1 and W[2, 3] == -1


def test_project_simplex_properties():
    rng = np.random.default_rng(0)
    for n in (2, 4, 16):
        for _ in range(50):
            p = rng.normal(1.0 / n, 0.3, size=n)
            q = po.project_simplex(p)
            assert np.all(q >= 0) and abs(q.sum() - 1) < 1e-12
            assert np.allclose(po.project_simplex(q), q, atol=1e-15)
            # Euclidean optimality: q - p is constant on the support and no larger off it
            sup = q > 0
            d = (q - p)[sup]
            assert np.ptp(d) < 1e-12 and np.all((q - p)[~sup] >= d[0] - 1e-12)
    assert np.array_equal(po.project_simplex(np.array([0.25, 0.25, 0.25, 0.25])), np.full(4, 0.25))


def test_simple_projection_fixed_point_and_clipping():
    rng = np.random.default_rng(1)
    gate_ptr = np.array([0, 15, 18, 19, 34, 37])
    probs = []
    for g in range(5):
        k = gate_ptr[g + 1] - gate_ptr[g]
        p = np.concatenate([[0.0], rng.uniform(0, 0.01, size=k)])
        p[0] = 1 - p.sum()
        probs.append(p)
    lam = np.concatenate([(po.get_wht_matrix(len(p) - 1) @ p)[1:] for p in probs])
    out, up, pp = po.simple_project_gate_eigenvalues(lam, gate_ptr)
    assert np.allclose(out, lam, atol=1e-15) and np.allclose(up, np.concatenate(probs), atol=1e-15)
    noisy = lam + rng.normal(0, 5e-3, size=len(lam))
    out, up, pp = po.simple_project_gate_eigenvalues(noisy, gate_ptr)
    assert np.any(up < 0) and np.all(pp >= 0)
    assert not np.allclose(out, noisy)


def test_mahalanobis_projection_kkt():
    rng = np.random.default_rng(2)
    for n in (1, 3, 15):
        for _ in range(20):
            B = rng.normal(size=(n + 3, n))
            Q = B.T @ B + 0.1 * np.eye(n)
            v = rng.normal(0, 1, size=n)
            y = po.scs_project_nonnegative(v, Q, precision=0.0)
            grad = Q @ (y - v)
            assert np.all(y >= 0) and np.all(grad >= -1e-9) and abs(y @ grad) < 1e-9
    # a diagonal precision matrix gives the Euclidean clamp
    v = np.array([0.3, -0.2, 0.0, 0.5])
    assert np.allclose(po.scs_project_nonnegative(v, np.diag([1.0, 2.0, 3.0, 4.0])), np.maximum(v, 0))
    # split projection with P = identity-like blocks keeps feasible gates unchanged
    gate_ptr = np.array([0, 3, 18])
    p1 = np.array([0.97, 0.01, 0.015, 0.005])
    p2 = np.concatenate([[0.9], np.full(15, 0.1 / 15)])
    lam = np.concatenate([(po.get_wht_matrix(3) @ p1)[1:], (po.get_wht_matrix(15) @ p2)[1:]])
    out, up, pp = po.split_project_gate_eigenvalues(lam, np.eye(18), gate_ptr)
    assert np.allclose(out, lam, atol=1e-12)


def test_full_projection_oracle_properties():
    """full_project_gate_eigenvalues (src/estimate.jl:532-573): with a precision matrix that is block diagonal
    by gate the simultaneous projection IS the per-gate one (split_project, :580-628); with correlations between
    the gates it satisfies the Karush-Kuhn-Tucker conditions of the joint problem and never does worse in the
    Mahalanobis distance than the split solution."""
    rng = np.random.default_rng(11)
    gate_ptr = np.array([0, 1, 4, 19, 22, 37])
    N = int(gate_ptr[-1])
    lam = np.concatenate([(po.get_wht_matrix(b - a) @ np.abs(rng.dirichlet(np.full(b - a + 1, 0.3))))[1:]
                          for a, b in zip(gate_ptr[:-1], gate_ptr[1:])])
    lam = lam + rng.normal(0, 0.02, size=N)          # noisy: some probabilities go negative
    blocks = []
    for a, b in zip(gate_ptr[:-1], gate_ptr[1:]):
        B = rng.normal(size=(b - a + 2, b - a))
        blocks.append(B.T @ B + 0.1 * np.eye(b - a))
    import scipy.linalg as sla
    Pd = sla.block_diag(*blocks)
    full = po.full_project_gate_eigenvalues(lam, Pd, gate_ptr, precision=0.0)
    split = po.split_project_gate_eigenvalues(lam, Pd, gate_ptr, precision=0.0)
    assert np.any(full[1] < 0)
    for x, y in zip(full, split):
        assert np.allclose(x, y, atol=1e-10)
    C = rng.normal(size=(N + 5, N))
    P = Pd + 0.3 * C.T @ C                            # gates correlated
    out, up, pp = po.full_project_gate_eigenvalues(lam, P, gate_ptr, precision=0.0)
    for a, b in zip(gate_ptr[:-1], gate_ptr[1:]):     # every gate's distribution sums to one (only the
                                                      # non-identity probabilities are constrained, as in the reference)
        g = np.searchsorted(gate_ptr, a)
        assert abs(pp[a + g:b + g + 1].sum() - 1.0) < 1e-12
    dist = lambda x: float((x - lam) @ P @ (x - lam))  # noqa: E731
    assert dist(out) <= dist(po.split_project_gate_eigenvalues(lam, P, gate_ptr, precision=0.0)[0]) * (1 + 1e-12)
    # KKT in the space of unpadded probabilities: Q (y - v) >= 0, complementary to y
    T = sla.block_diag(*[po.probs_transform_block(b - a) for a, b in zip(gate_ptr[:-1], gate_ptr[1:])])
    strip = np.concatenate([np.arange(a + g + 1, b + g + 1) for g, (a, b) in enumerate(zip(gate_ptr[:-1], gate_ptr[1:]))])
    y, v = pp[strip], up[strip]
    grad = (T @ P @ T.T) @ (y - v)
    s = np.abs((T @ P @ T.T) @ v).max()
    assert np.all(y >= 0) and np.all(grad >= -1e-9 * s) and abs(y @ grad) <= 1e-9 * s
