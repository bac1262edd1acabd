"""Host-side mirror of the reference's simplex projections of the fitted gate eigenvalues
(src/estimate.jl:580-663, src/utils.jl:36-97): same names and argument meaning; the per-gate work
runs on the GPU through aces_project_simplex / aces_project_nonnegative_batched.

Gates are groups of gate eigenvalues (GateIndex.indices, src/noise.jl:16-24): gate g owns
gate_idx[gate_ptr[g] .. gate_ptr[g+1]-1]; consecutive ranges when gate_idx is None.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import _ptr


def _groups(N, gate_ptr, gate_idx):
    gp = np.ascontiguousarray(gate_ptr, dtype=np.int64)
    gi = np.arange(N, dtype=np.int32) if gate_idx is None else np.ascontiguousarray(gate_idx, dtype=np.int32)
    assert gp[0] == 0 and gp[-1] == len(gi)
    return gp, gi


def simple_project_gate_eigenvalues(ctx: _lib.Context, est_unproj_gate_eigenvalues, gate_ptr, gate_idx=None):
    """simple_project_gate_eigenvalues(est_unproj_gate_eigenvalues, gate_data) (src/estimate.jl:635-663):
    returns (gate eigenvalues, unprojected padded probabilities, projected padded probabilities)."""
    un = np.ascontiguousarray(est_unproj_gate_eigenvalues, dtype=np.float64)
    N = len(un)
    gp, gi = _groups(N, gate_ptr, gate_idx)
    G = len(gp) - 1
    out = np.zeros(N)
    up, pp = np.zeros(len(gi) + G), np.zeros(len(gi) + G)
    ctx.check(ctx.lib.aces_project_simplex(ctx.handle, N, G, _ptr(gp), _ptr(gi), 0, _ptr(un), _ptr(out), _ptr(up), _ptr(pp)))
    return out, up, pp


def project_nonnegative_batched(ctx: _lib.Context, vectors, precision_blocks, precision: float = 1e-8):
    """scs_project_nonnegative (src/utils.jl:57-97) for many small problems in one launch: `vectors` and
    `precision_blocks` are lists of length-n vectors and n x n symmetric matrices (n <= 15)."""
    ptr = np.zeros(len(vectors) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(v) for v in vectors])
    v = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in vectors])) if len(vectors) else np.zeros(1)
    Q = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float64).ravel() for q in precision_blocks])) if len(vectors) else np.zeros(1)
    y = np.zeros(max(int(ptr[-1]), 1))
    bad = C.c_int32()
    ctx.check(ctx.lib.aces_project_nonnegative_batched(ctx.handle, len(vectors), _ptr(ptr), _ptr(Q), _ptr(v), float(precision),
                                                       _ptr(y), C.byref(bad)))
    return [y[ptr[i]:ptr[i + 1]] for i in range(len(vectors))]


def _wht(n_eig: int) -> np.ndarray:
    """Symplectically ordered Walsh-Hadamard matrix of a gate with n_eig eigenvalues (src/noise.jl:285-312)."""
    P = n_eig + 1
    if P == 2:
        return np.array([[1.0, 1.0], [1.0, -1.0]])
    n = {4: 1, 16: 2}[P]
    idx = np.arange(P)
    lo = (idx[:, None] >> np.arange(n)) & 1
    hi = (idx[:, None] >> (n + np.arange(n))) & 1
    return np.where(((lo @ hi.T + hi @ lo.T) & 1) == 1, -1.0, 1.0)


def split_project_gate_eigenvalues(ctx: _lib.Context, est_unproj_gate_eigenvalues, precision_blocks, gate_ptr,
                                   precision: float = 1e-8):
    """split_project_gate_eigenvalues (src/estimate.jl:580-628) with the precision matrix given by its
    per-gate diagonal blocks (all that function reads from it): `precision_blocks[g]` is
    est_precision_matrix[indices_g, indices_g].  The tiny per-gate transforms (T P T' with the gate's
    Walsh-Hadamard block) are formed on the host, the G quadratic programs solved in one launch."""
    un = np.asarray(est_unproj_gate_eigenvalues, dtype=np.float64)
    G = len(gate_ptr) - 1
    vs, Qs, qs, Ws = [], [], [], []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        W = _wht(b - a)
        q = (W / (b - a + 1)) @ np.concatenate([[1.0], un[a:b]])
        T = W[1:, 1:] - W[1:, :1]
        vs.append(q[1:])
        Qs.append(T @ np.asarray(precision_blocks[g]) @ T.T)
        qs.append(q)
        Ws.append(W)
    ys = project_nonnegative_batched(ctx, vs, Qs, precision)
    out = un.copy()
    pp = []
    for g in range(G):
        a, b = int(gate_ptr[g]), int(gate_ptr[g + 1])
        pr = np.concatenate([[1.0 - ys[g].sum()], ys[g]])
        out[a:b] = (Ws[g] @ pr)[1:]
        pp.append(pr)
    return out, np.concatenate(qs), np.concatenate(pp)


def full_project_gate_eigenvalues(dd, ls_type: str, est_unproj_gate_eigenvalues, gate_ptr, precision: float = 1e-8):
    """full_project_gate_eigenvalues(est_unproj_gate_eigenvalues, est_precision_matrix, gate_data; precision)
    (src/estimate.jl:532-573), the simultaneous projection of all gates (the reference's default below 5000
    gate eigenvalues), for the estimator of the LAST `dd.fit(ls_type, ...)`: the precision matrix
    (calc_precision_matrix, src/merit.jl:754-770) is rebuilt on the device from that fit's Gram matrix and never
    leaves it; the N-variable non-negative quadratic program is solved exactly by block principal pivoting
    (aces_design_project_full).  The per-gate transforms in and out are the host's, as in the split version.
    Returns (gate eigenvalues, unprojected padded probabilities, projected padded probabilities)."""
    from .merit import LS_TYPES

    un = np.ascontiguousarray(est_unproj_gate_eigenvalues, dtype=np.float64)
    N = len(un)
    gp, gi = _groups(N, gate_ptr, None)
    G = len(gp) - 1
    v, blocks, qs, Ws = np.zeros(N), [], [], []
    for g in range(G):
        a, b = int(gp[g]), int(gp[g + 1])
        W = _wht(b - a)
        q = (W / (b - a + 1)) @ np.concatenate([[1.0], un[a:b]])
        v[a:b] = q[1:]
        blocks.append((W[1:, 1:] - W[1:, :1]).ravel())     # block of probs_transform = pad' * W * pad_probs
        qs.append(q)
        Ws.append(W)
    T = np.ascontiguousarray(np.concatenate(blocks))
    y = np.zeros(N)
    its = C.c_int32()
    ctx = dd.ctx
    ctx.check(ctx.lib.aces_design_project_full(ctx.handle, dd.handle, LS_TYPES[ls_type], G, _ptr(gp), _ptr(gi), 0, _ptr(un),
                                               _ptr(T), _ptr(v), float(precision), _ptr(y), C.byref(its)))
    out = un.copy()
    pp = []
    for g in range(G):
        a, b = int(gp[g]), int(gp[g + 1])
        pr = np.concatenate([[1.0 - y[a:b].sum()], y[a:b]])
        out[a:b] = (Ws[g] @ pr)[1:]
        pp.append(pr)
    full_project_gate_eigenvalues.last_iterations = int(its.value)
    return out, np.concatenate(qs), np.concatenate(pp)
