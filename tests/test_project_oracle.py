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
