"""Flat, array-only mirror of the `Design` fields the ACES estimation path consumes.

The reference keeps these in nested Julia structs (`Design`, `Mapping`, `Pauli`:
src/design.jl:11-19,45-60,95-114).  The hot path only reads a few derived tables, built once
per design by get_experiment_data / get_covariance_experiment_data (src/estimate.jl:308-428)
and get_covariance_mapping_data (src/merit.jl:144-163).  `FlatDesign` holds exactly those
tables; the Julia shim (julia/ACESB200.jl: `flatten_design`) produces the same arrays from a
real `Design`.  Indices here are 0-based except `meas_indices`, which keep the reference's
1-based qubit numbering.

`synthetic_design` generates designs of the surface-code shapes named in BASELINE.json
 because the real design generator (Clifford propagation, greedy
experiment packing) is out of scope and Julia is not available in this image.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp


@dataclass
class FlatDesign:
    n_qubits: int
    experiment_numbers: np.ndarray          # [T] X_t                 (d.experiment_numbers)
    shot_weights: np.ndarray                # [T] w_t                 (d.shot_weights)
    tuple_times: np.ndarray                 # [T] tau_t               (d.tuple_times)
    mapping_lengths: np.ndarray             # [T] m_t                 (length.(d.mapping_ensemble))
    meas_indices: List[List[np.ndarray]]    # [T][m_t] 1-based qubits (get_support(m.final))
    pauli_signs: List[np.ndarray]           # [T][m_t] bool           (m.final.pauli[2n+1])
    experiment_ensemble: List[List[np.ndarray]]  # [T][X_t] mapping indices of experiment j
    full_covariance: bool = False
    # off-diagonal covariance keys (sorted, p1 < p2) and their Paulis (src/merit.jl:149-154,
    # src/estimate.jl:388-400)
    cov_keys: List[np.ndarray] = field(default_factory=list)           # [T][C_t, 2]
    cov_meas_indices: List[List[np.ndarray]] = field(default_factory=list)
    cov_signs: List[np.ndarray] = field(default_factory=list)
    cov_experiment_ensemble: List[List[np.ndarray]] = field(default_factory=list)  # [T][X_t] key indices
    # experiment counts stored in covariance_dict (src/design.jl:803-817,844-853)
    diag_counts: List[np.ndarray] = field(default_factory=list)        # [T][m_t]
    cov_counts: List[np.ndarray] = field(default_factory=list)         # [T][C_t]
    matrix: Optional[sp.csc_matrix] = None  # d.matrix, M x N, int32
    gate_eigenvalues: Optional[np.ndarray] = None  # d.c.gate_eigenvalues (true values, synthetic)
    # design rows of the covariance Paulis (covariance_dict[key][1].design_row), CSR per tuple
    cov_rows: List = field(default_factory=list)
    # gates as groups of consecutive gate eigenvalues (GateIndex.indices, src/noise.jl:16-24): gate g owns
    # columns gate_ptr[g] .. gate_ptr[g+1]-1 of the design matrix (1, 3 or 15 of them)
    gate_ptr: Optional[np.ndarray] = None
    gate_types: Optional[List[str]] = None   # Gate.type of every gate, in the order of gate_ptr (src/tableau.jl:88-92)
    name: str = ""

    @property
    def tuple_number(self) -> int:
        return len(self.experiment_numbers)

    @property
    def M(self) -> int:
        return int(np.sum(self.mapping_lengths))

    @property
    def n_bytes(self) -> int:
        return (self.n_qubits + 7) // 8

    def experiment_index(self, t: int, j: int) -> int:
        return int(np.sum(self.experiment_numbers[:t])) + j

    @property
    def experiment_number(self) -> int:
        return int(np.sum(self.experiment_numbers))


def get_covariance_experiment_ensemble(fd: FlatDesign) -> List[List[np.ndarray]]:
    """src/estimate.jl:349-376: per experiment, the covariance keys whose two Paulis are both
    measured in it, in the order of the double loop over the (sorted) experiment."""
    out = []
    for t in range(fd.tuple_number):
        keys = fd.cov_keys[t] if fd.cov_keys else np.zeros((0, 2), dtype=np.int64)
        if len(keys) == 0:
            out.append([np.zeros(0, dtype=np.int64) for _ in range(int(fd.experiment_numbers[t]))])
            continue
        key_index = {(int(a), int(b)): i for i, (a, b) in enumerate(keys)}
        per = []
        for exp in fd.experiment_ensemble[t]:
            found = []
            E = len(exp)
            for i1 in range(E):
                for i2 in range(i1 + 1, E):
                    k = key_index.get((int(exp[i1]), int(exp[i2])))
                    if k is not None:
                        found.append(k)
            per.append(np.asarray(found, dtype=np.int64))
        out.append(per)
    return out


def shots_divide(fd: FlatDesign, shots_set) -> Tuple[np.ndarray, np.ndarray]:
    """src/simulate.jl:98-104: shots_divided[t, s] = ceil(S_s * w_t / X_t), and its row max."""
    shots_set = np.asarray(shots_set, dtype=np.int64)
    w = np.asarray(fd.shot_weights, dtype=np.float64)
    X = np.asarray(fd.experiment_numbers, dtype=np.float64)
    shots_divided_float = shots_set[None, :].astype(np.float64) * w[:, None] / X[:, None]
    shots_divided = np.ceil(shots_divided_float).astype(np.int64)
    shots_maximum = shots_divided.max(axis=1)
    return shots_divided, shots_maximum


def build_pauli_tables(fd: FlatDesign):
    """Flattens a design into the CSR tables of aces_tables_create.

    Experiment e = (t, j) in tuple-major order measures first its circuit-eigenvalue Paulis
    (experiment_ensemble[t][j], src/simulate.jl:196-198) and then, if the tuple has covariance
    keys, its covariance Paulis (src/simulate.jl:199-205) -- both sets in one table so that one
    launch covers them.  Returns (exp_ptr, pauli_ptr, pauli_bits, n_eig) where n_eig[e] is the
    number of circuit-eigenvalue Paulis of experiment e.
    """
    exp_ptr = [0]
    pauli_ptr = [0]
    bits: List[int] = []
    n_eig = []
    has_cov = bool(fd.cov_keys)
    for t in range(fd.tuple_number):
        for j in range(int(fd.experiment_numbers[t])):
            exp = fd.experiment_ensemble[t][j]
            for p in exp:
                sup = fd.meas_indices[t][int(p)]
                bits.extend(int(m) - 1 for m in sup)
                pauli_ptr.append(len(bits))
            n_eig.append(len(exp))
            cnt = len(exp)
            if has_cov and len(fd.cov_keys[t]) > 0:
                for c in fd.cov_experiment_ensemble[t][j]:
                    sup = fd.cov_meas_indices[t][int(c)]
                    bits.extend(int(m) - 1 for m in sup)
                    pauli_ptr.append(len(bits))
                cnt += len(fd.cov_experiment_ensemble[t][j])
            exp_ptr.append(exp_ptr[-1] + cnt)
    return (
        np.asarray(exp_ptr, dtype=np.int64),
        np.asarray(pauli_ptr, dtype=np.int64),
        np.asarray(bits, dtype=np.int32),
        np.asarray(n_eig, dtype=np.int64),
    )


# ---------------------------------------------------------------------------------------------
# Synthetic designs of the BASELINE.json shapes ("derived sizes")
# ---------------------------------------------------------------------------------------------

def surface_code_shape(kind: str, d: int) -> Dict[str, int]:
    """n, N, M, T of the default design for a distance-d code."""
    if kind == "rotated":
        return dict(n=2 * d * d - 1, N=88 * d * d - 36 * d - 25, M=168 * d * d - 72 * d - 48, T=16)
    if kind == "unrotated":
        n = (2 * d - 1) ** 2
        # 7 layers / 6 unique -> 7 basic + 5 repeated tuples; N, M from the survey table at d=7
        # scaled with n for other distances (only d=7 is a named configuration).
        N = int(round(6189 * n / 169))
        M = int(round(11700 * n / 169))
        return dict(n=n, N=N, M=M, T=12)
    raise ValueError(f"unknown code kind {kind!r}")


def _tuple_experiment_numbers(kind: str) -> np.ndarray:
    # 3 experiments for tuples whose widest gate acts on one qubit, 9 for two-qubit layers
    # (src/tuples.jl:56-68).  Rotated XZZX/CZ circuit: H, CZ, H, CZ, CZ, H, CZ, H, R
    # (src/circuits/rotated_planar.jl:306-325): the empty tuple + 8 single-layer tuples, then
    # the 7 repeated tuples (no repeated reset tuple).
    if kind == "rotated":
        return np.array([3, 3, 9, 3, 9, 9, 3, 9, 3, 3, 9, 3, 9, 9, 3, 9], dtype=np.int64)
    return np.array([3, 3, 9, 9, 9, 9, 3, 3, 9, 9, 9, 9], dtype=np.int64)


def synthetic_design(
    kind: str = "rotated",
    d: int = 9,
    full_covariance: bool = False,
    seed: int = 0x5EED,
    stress_weights: bool = False,
    with_matrix: bool = False,
) -> FlatDesign:
    """A design with the shapes of SURVEY.md section 8d.

    Per experiment: E_j = n + floor(n/2) circuit-eigenvalue Paulis (two thirds weight 1, one third
    weight 2 on neighbouring qubits) and, with full_covariance, floor(n/2) covariance Paulis
    (products of two measured Paulis).  Shot weights are proportional to X_t.  With
    stress_weights a few supports get weight 3..9 to exercise the wide-support path.
    """
    rng = np.random.default_rng(seed)
    shape = surface_code_shape(kind, d)
    n, M, T = shape["n"], shape["M"], shape["T"]
    X = _tuple_experiment_numbers(kind)
    assert len(X) == T
    E_target = n + n // 2
    # mapping lengths: proportional to X_t, every class (p mod X_t) non-empty, sum == M
    base = np.floor(M * X / X.sum()).astype(np.int64)
    base[: int(M - base.sum())] += 1
    m_t = base
    assert m_t.sum() == M and np.all(m_t >= X)
    meas_indices, signs, experiments = [], [], []
    for t in range(T):
        sup_t = []
        cls_len = max(int(m_t[t]) // int(X[t]), 1)
        for p in range(int(m_t[t])):
            # Position of the Pauli on the qubit line: class c = p mod X_t starts where class c-1
            # ends, so that an experiment (its own class plus the start of the next one) walks
            # over every qubit: weight 1 on each qubit, weight 2 on the disjoint neighbour pairs.
            r = (p // int(X[t]) + (p % int(X[t])) * cls_len) % E_target
            if r % 3 == 2:
                k = (r // 3) % max(n // 2, 1)
                sup = [2 * k + 1, min(2 * k + 2, n)]
                if sup[0] == sup[1]:
                    sup = [sup[0]]
            else:
                sup = [((2 * (r // 3) + r % 3) % n) + 1]
            if stress_weights and p % 29 == 7:
                w = 3 + (p // 29) % 7
                sup = sorted(set(int(q) for q in rng.integers(1, n + 1, size=w)))
            sup_t.append(np.asarray(sup, dtype=np.int64))
        meas_indices.append(sup_t)
        signs.append(rng.integers(0, 2, size=int(m_t[t])).astype(bool))
        # Experiment j measures its own class and, like the reference's greedy packing that re-adds
        # compatible Paulis of other experiments (src/design.jl:745-767), Paulis of the other classes
        # at the line positions its own class does not reach: E_target Paulis, every qubit touched.
        pos = (np.arange(int(m_t[t])) // int(X[t]) + (np.arange(int(m_t[t])) % int(X[t])) * cls_len) % E_target
        exps = []
        for j in range(int(X[t])):
            own = np.arange(j, int(m_t[t]), int(X[t]), dtype=np.int64)[:E_target]
            have = np.zeros(E_target, dtype=bool)
            have[pos[own]] = True
            extra = []
            for p in range(int(m_t[t])):
                if len(own) + len(extra) >= E_target:
                    break
                if p % int(X[t]) != j and not have[pos[p]]:
                    have[pos[p]] = True
                    extra.append(p)
            exps.append(np.sort(np.concatenate([own, np.asarray(extra, dtype=np.int64)])))
        experiments.append(exps)
    w_t = X / X.sum()
    fd = FlatDesign(
        n_qubits=n,
        experiment_numbers=X,
        shot_weights=w_t.astype(np.float64),
        tuple_times=np.ones(T, dtype=np.float64),
        mapping_lengths=m_t,
        meas_indices=meas_indices,
        pauli_signs=signs,
        experiment_ensemble=experiments,
        full_covariance=full_covariance,
        name=f"{kind}_d{d}" + ("_gls" if full_covariance else ""),
    )
    # diagonal counts: number of experiments measuring each mapping (src/design.jl:803-817)
    fd.diag_counts = []
    for t in range(T):
        c = np.zeros(int(m_t[t]), dtype=np.int64)
        for exp in experiments[t]:
            c[exp] += 1
        fd.diag_counts.append(c)
    if full_covariance:
        n_cov = n // 2
        fd.cov_keys, fd.cov_meas_indices, fd.cov_signs, fd.cov_counts = [], [], [], []
        for t in range(T):
            keyset = {}
            cls_len_t = max(int(m_t[t]) // int(X[t]), 1)
            for exp in experiments[t]:
                # pairs of Paulis measured together whose product is also estimated: neighbours on
                # the qubit line (products of Paulis on adjacent qubits, as in a real design)
                line = (exp // int(X[t]) + (exp % int(X[t])) * cls_len_t) % E_target
                byline = exp[np.argsort(line, kind="stable")]
                picks = rng.choice(len(byline) - 1, size=min(n_cov, len(byline) - 1), replace=False)
                for i in np.sort(picks):
                    a, b = int(byline[i]), int(byline[i + 1])
                    keyset[(min(a, b), max(a, b))] = 0
            keys = np.asarray(sorted(keyset.keys()), dtype=np.int64).reshape(-1, 2)
            fd.cov_keys.append(keys)
            cm, cs = [], []
            for a, b in keys:
                sa = set(int(q) for q in meas_indices[t][a])
                sb = set(int(q) for q in meas_indices[t][b])
                cm.append(np.asarray(sorted(sa ^ sb), dtype=np.int64))
                cs.append(bool(signs[t][a]) ^ bool(signs[t][b]))
            fd.cov_meas_indices.append(cm)
            fd.cov_signs.append(np.asarray(cs, dtype=bool))
        fd.cov_experiment_ensemble = get_covariance_experiment_ensemble(fd)
        for t in range(T):
            c = np.zeros(len(fd.cov_keys[t]), dtype=np.int64)
            for ce in fd.cov_experiment_ensemble[t]:
                c[ce] += 1
            fd.cov_counts.append(c)
    if with_matrix:
        _attach_synthetic_matrix(fd, shape["N"], rng)
    return fd


def _attach_synthetic_matrix(fd: FlatDesign, N: int, rng) -> None:
    """Synthetic design matrix of SURVEY.md section 8d ("Fit synthetic input").

    A is M x N Int32 in T row blocks.  N rows spread over the blocks form a permuted identity
    (one row e_j per gate eigenvalue: full column rank, like the SPAM rows of a real design);
    every other row has 2-6 entries inside its block's column window with values in
    {1, r, (r-1)/2, (r+1)/2}, r the repeat number of the block (src/design.jl:475-531: an
    entry counts how often a gate eigenvalue enters the circuit eigenvalue).  Gate eigenvalues
    are log-normal around 1 - O(r_1, r_2, r_m) (src/noises/lognormal.jl:157-235, std 0.5).
    The covariance Paulis get design rows row_p + row_q +- e_j so that lambda_pq differs from
    lambda_p * lambda_q by one gate eigenvalue (a small, well-conditioned off-diagonal).
    """
    T, M = fd.tuple_number, fd.M
    m_t = np.asarray(fd.mapping_lengths, dtype=np.int64)
    assert M >= N, "the synthetic design needs at least as many circuit as gate eigenvalues"
    # identity rows per block, proportional to the block size
    ident = np.floor(N * m_t / M).astype(np.int64)
    ident[: int(N - ident.sum())] += 1
    assert ident.sum() == N and np.all(ident <= m_t)
    col_start = np.concatenate([[0], np.cumsum(ident)])
    repeat = [1] * 9 + [63, 143, 63, 143, 143, 63, 143] if T == 16 else [1] * 7 + [63, 143, 143, 143, 183]
    perm = rng.permutation(N)
    rows, cols, vals = [], [], []
    row0 = 0
    for t in range(T):
        width = int(min(N, max(N // 4, 2 * ident[t])))
        lo = int(min(max(col_start[t] - (width - ident[t]) // 2, 0), N - width))
        r = repeat[t]
        choices = np.array([1, r, max((r - 1) // 2, 1), (r + 1) // 2], dtype=np.int64)
        for i in range(int(m_t[t])):
            if i < ident[t]:
                rows.append(row0 + i)
                cols.append(int(perm[col_start[t] + i]))
                vals.append(1)
            else:
                k = int(rng.integers(2, 7))
                cc = lo + rng.choice(width, size=k, replace=False)
                rows.extend([row0 + i] * k)
                cols.extend(int(perm[c]) for c in cc)
                vals.extend(int(v) for v in rng.choice(choices, size=k))
        row0 += int(m_t[t])
    A = sp.csc_matrix((np.asarray(vals, dtype=np.int32), (np.asarray(rows), np.asarray(cols))),
                      shape=(M, N), dtype=np.int32)
    A.sum_duplicates()
    A.sort_indices()
    fd.matrix = A
    # gate eigenvalues: 1 - error with log-normal errors of scale ~0.4% / repeat (keeps the
    # circuit eigenvalues of the repeated blocks away from the 0.1 clipping threshold)
    err = 0.004 / 30.0 * np.exp(0.5 * rng.standard_normal(N) - 0.125)
    clean = rng.permutation(N)[: max(N // 20, 1)]  # a few very clean gates (see cov_rows below)
    err[clean] *= 0.02
    fd.gate_eigenvalues = 1.0 - err
    if fd.full_covariance:
        Acsr = sp.csr_matrix(A).astype(np.int64)
        lam = np.exp(Acsr.astype(np.float64) @ np.log(fd.gate_eigenvalues))
        order = np.argsort(err)
        err_sorted = err[order]
        cov_rows = []
        row0 = 0
        for t in range(T):
            keys = fd.cov_keys[t]
            if len(keys):
                Rp = Acsr[row0 + keys[:, 0]]
                Rq = Acsr[row0 + keys[:, 1]]
                # The extra gate of a pair is one whose error is at most a fiftieth of the smaller
                # of the two variances 1 - lambda^2 (none if no gate is that clean): every
                # covariance block is then diagonally dominant, hence positive definite.
                var = 1.0 - lam[row0:row0 + int(m_t[t])] ** 2
                thr = 0.02 * np.minimum(var[keys[:, 0]], var[keys[:, 1]])
                n_ok = np.searchsorted(err_sorted, thr, side="right")
                has = n_ok > 0
                pick = np.maximum(n_ok - 1 - rng.integers(0, 50, size=len(keys)), 0)
                extra_cols = order[pick]
                sign = np.where(rng.integers(0, 2, size=len(keys)) == 1, 1, -1) * has
                extra = sp.csr_matrix((sign, (np.arange(len(keys)), extra_cols)), shape=(len(keys), N))
                extra.eliminate_zeros()
                cov_rows.append((Rp + Rq + extra).tocsr())
            else:
                cov_rows.append(sp.csr_matrix((0, N), dtype=np.int64))
            row0 += int(m_t[t])
        fd.cov_rows = cov_rows


def synthetic_estimates(fd: FlatDesign, shots: float = 10**8, seed: int = 0, noise: bool = True):
    """Synthetic per-experiment estimate ensembles for a design with a matrix: what
    simulate_stim_estimate returns for one budget ([t][pauli] -> one value per experiment that
    measures it), drawn around the true circuit eigenvalues exp(A log g) (src/design.jl:255-258)
    with the shot-noise variance of src/merit.jl:245-254.  Input generator for tests and bench.py
    (Stim sampling is not part of the hot path)."""
    rng = np.random.default_rng(seed)
    lg = np.log(fd.gate_eigenvalues)
    lam = np.exp(fd.matrix.astype(np.float64) @ lg)
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    lam_ens = [lam[off[t]:off[t + 1]] for t in range(fd.tuple_number)]
    eig, cov = [], []
    for t in range(fd.tuple_number):
        sig = np.sqrt(np.maximum(1 - lam_ens[t] ** 2, 1e-10) / (shots * fd.shot_weights[t] / fd.experiment_numbers[t]))
        eig_t = []
        for p in range(int(fd.mapping_lengths[t])):
            k = int(fd.diag_counts[t][p])
            v = lam_ens[t][p] + (sig[p] * rng.standard_normal(k) if noise else np.zeros(k))
            eig_t.append(list(np.minimum(v, 1.0)))
        eig.append(eig_t)
        cov_t = []
        if fd.full_covariance and len(fd.cov_keys[t]):
            # the pair estimate follows the two single estimates (they come from the same shots):
            # lambda_pq_hat = lambda_p_hat * lambda_q_hat * (true ratio) + a little noise, which
            # keeps the estimated covariance blocks positive definite as in a real experiment
            cov_lam = np.exp(fd.cov_rows[t].astype(np.float64) @ lg)
            for c, (p, q) in enumerate(fd.cov_keys[t]):
                k = max(int(fd.cov_counts[t][c]), 1)
                s = 1e-9 if noise else 0.0
                ratio = cov_lam[c] / (lam_ens[t][p] * lam_ens[t][q])
                base = np.mean(eig_t[p]) * np.mean(eig_t[q]) * ratio
                cov_t.append(list(np.minimum(base + s * rng.standard_normal(k), 1.0)))
        cov.append(cov_t)
    return eig, cov
