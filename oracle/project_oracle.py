"""project_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy / scipy) restatement of the reference's simplex projections of the fitted gate
eigenvalues.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline legs may import this module.

Each function cites the reference code it follows (paths relative to the QuantumACES.jl root).
Gates are groups of consecutive gate eigenvalues: gate g owns indices gate_ptr[g] .. gate_ptr[g+1]-1.

Third-party arithmetic absent from /root/reference: scs_project_nonnegative calls SCS.jl "2"
(src/utils.jl:72-92), an iterative conic solver run to eps_abs = eps_rel = 1e-8.  Its problem --
min 1/2 x'Px + c'x, x + s = 0, s >= 0 with c = scaling * P v, result -x / scaling -- is the
non-negative quadratic program min (y - v)' P (y - v), y >= 0, restated here through LAPACK
Cholesky + scipy's Lawson-Hanson NNLS (an exact solver): parity with SCS holds to its tolerance
(parity unpinned: no Julia / SCS in this image).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt


def get_wht_matrix(n_eigenvalues: int, inverse: bool = False) -> np.ndarray:
    """src/noise.jl:285-312: symplectically ordered Walsh-Hadamard matrix of a gate with
    n_eigenvalues = 4^n - 1 eigenvalues, or [1 1; 1 -1] for SPAM / reset gates (one eigenvalue)."""
    P = n_eigenvalues + 1
    if P == 2:
        W = np.array([[1.0, 1.0], [1.0, -1.0]])
    else:
        n = {4: 1, 16: 2, 64: 3}[P]
        idx = np.arange(P)
        lo = (idx[:, None] >> np.arange(n)) & 1           # a[1:n]
        hi = (idx[:, None] >> (n + np.arange(n))) & 1     # a[(n+1):(2n)]
        sym = (lo @ hi.T + hi @ lo.T) & 1
        W = np.where(sym == 1, -1.0, 1.0)
    return W / P if inverse else W


def project_simplex(p: np.ndarray) -> np.ndarray:
    """src/utils.jl:36-50, operation by operation (the sums are sequential, as Julia's cumsum is for
    short vectors)."""
    p = np.asarray(p, dtype=np.float64)
    assert not np.any(np.isnan(p))
    s = np.sort(p)[::-1]
    c = np.empty_like(s)
    acc = 0.0
    for i, x in enumerate(s):
        acc = x if i == 0 else acc + x
        c[i] = acc
    sp = np.array([(1.0 / (i + 1)) * (1.0 - c[i]) for i in range(len(s))])
    last = np.nonzero(s + sp > 0)[0][-1]
    out = np.maximum(p + sp[last], 0.0)
    assert abs(out.sum() - 1.0) < 1e-8
    return out


def _seq_matvec(W: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Sparse mat-vec in Julia's order: every output accumulates its terms for ascending j."""
    out = np.zeros(W.shape[0])
    for j in range(W.shape[1]):
        out = out + W[:, j] * x[j]
    return out


def simple_project_gate_eigenvalues(unproj, gate_ptr):
    """src/estimate.jl:635-663.  Returns (gate eigenvalues, unprojected padded probabilities,
    projected padded probabilities); padded vectors are laid out gate after gate."""
    unproj = np.asarray(unproj, dtype=np.float64)
    G = len(gate_ptr) - 1
    out = unproj.copy()
    up, pp = [], []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        lam = np.concatenate([[1.0], unproj[a:b]])                    # pad_transform * g + pad_mask
        q = _seq_matvec(get_wht_matrix(b - a, inverse=True), lam)
        pr = project_simplex(q)
        out[a:b] = _seq_matvec(get_wht_matrix(b - a), pr)[1:]         # pad_transform' * wht_transform * p
        up.append(q)
        pp.append(pr)
    return out, np.concatenate(up), np.concatenate(pp)


def scs_project_nonnegative(v, Q, precision: float = 1e-8):
    """src/utils.jl:57-97: argmin (y - v)' Q (y - v) over y >= 0, entries below `precision` zeroed."""
    v = np.asarray(v, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    Q = np.triu(Q) + np.triu(Q, 1).T                                  # P = UpperTriangular(precision_matrix)
    L = np.linalg.cholesky(Q)
    y, _ = sopt.nnls(L.T, L.T @ v, maxiter=100 * len(v))
    y = y.copy()
    y[y < precision] = 0.0
    return y


def probs_transform_block(n_eigenvalues: int) -> np.ndarray:
    """The gate's block of probs_transform = pad_transform' * wht_transform * pad_transform_probs
    (src/estimate.jl:597): T[a, b] = W[a+1, b+1] - W[a+1, 0]."""
    W = get_wht_matrix(n_eigenvalues)
    return W[1:, 1:] - W[1:, :1]


def split_project_gate_eigenvalues(unproj, precision_matrix, gate_ptr, precision: float = 1e-8):
    """src/estimate.jl:580-628 with a dense or scipy-sparse N x N precision matrix."""
    unproj = np.asarray(unproj, dtype=np.float64)
    G = len(gate_ptr) - 1
    out = unproj.copy()
    up, pp = [], []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        n = b - a
        W, Winv = get_wht_matrix(n), get_wht_matrix(n, inverse=True)
        q = Winv @ np.concatenate([[1.0], unproj[a:b]])
        T = probs_transform_block(n)
        Pgg = precision_matrix[a:b, a:b]
        Pgg = Pgg.toarray() if hasattr(Pgg, "toarray") else np.asarray(Pgg)
        Qg = T @ Pgg @ T.T
        y = scs_project_nonnegative(q[1:], Qg, precision=precision)
        pr = np.concatenate([[1.0 - y.sum()], y])                    # pad_transform_probs * y + pad_mask
        out[a:b] = (W @ pr)[1:]
        up.append(q)
        pp.append(pr)
    return out, np.concatenate(up), np.concatenate(pp)


def full_project_gate_eigenvalues(unproj, precision_matrix, gate_ptr, precision: float = 1e-8):
    """src/estimate.jl:532-573 with a dense or scipy-sparse N x N precision matrix: ONE non-negative quadratic
    program over the unpadded probabilities of all gates, Q = probs_transform * P * probs_transform' (block
    diagonal transform, src/estimate.jl:553-556)."""
    unproj = np.asarray(unproj, dtype=np.float64)
    N, G = len(unproj), len(gate_ptr) - 1
    P = precision_matrix.toarray() if hasattr(precision_matrix, "toarray") else np.asarray(precision_matrix, dtype=np.float64)
    T = np.zeros((N, N))
    v = np.zeros(N)
    qs = []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        q = get_wht_matrix(b - a, inverse=True) @ np.concatenate([[1.0], unproj[a:b]])
        T[a:b, a:b] = probs_transform_block(b - a)
        v[a:b] = q[1:]
        qs.append(q)
    y = scs_project_nonnegative(v, T @ P @ T.T, precision=precision)
    out = unproj.copy()
    pp = []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        pr = np.concatenate([[1.0 - y[a:b].sum()], y[a:b]])
        out[a:b] = (get_wht_matrix(b - a) @ pr)[1:]
        pp.append(pr)
    return out, np.concatenate(qs), np.concatenate(pp)
