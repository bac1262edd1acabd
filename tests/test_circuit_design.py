"""CPU tests of circuit_design.py (the array restatement of the reference's design generator that
feeds the kernels with real tables): SURVEY.md section 8c item 8 (the hand-checkable 3-qubit
example of test/creation_tests.jl:63-88), the derived sizes of the section 8 table, and a
cross-check of the batched Pauli conjugation against a literal per-Pauli stabiliser tableau written
here after src/tableau.jl:13-403 and src/design.jl:538-611."""
import numpy as np
import pytest

import aces_b200
from aces_b200 import circuit_design as cd


# ---------------------------------------------------------------------------------------------
# literal tableau (Aaronson-Gottesman), one Pauli at a time -- the way the reference does it
# ---------------------------------------------------------------------------------------------

class Tableau:
    def __init__(self, n):
        self.n = n
        self.t = np.zeros((2 * n + 1, 2 * n + 1), dtype=np.int64)
        for i in range(2 * n):
            self.t[i, i] = 1

    def hadamard(self, q):
        t, n = self.t, self.n
        for i in range(2 * n):
            t[i, 2 * n] = (t[i, 2 * n] + t[i, q] * t[i, n + q]) % 2
            t[i, q], t[i, n + q] = t[i, n + q], t[i, q]

    def phase(self, q):
        t, n = self.t, self.n
        for i in range(2 * n):
            t[i, 2 * n] = (t[i, 2 * n] + t[i, q] * t[i, n + q]) % 2
            t[i, n + q] = (t[i, n + q] + t[i, q]) % 2

    def cx(self, c, q):
        t, n = self.t, self.n
        for i in range(2 * n):
            t[i, 2 * n] = (t[i, 2 * n] + t[i, c] * t[i, n + q] * ((t[i, q] + t[i, n + c] + 1) % 2)) % 2
            t[i, q] = (t[i, q] + t[i, c]) % 2
            t[i, n + c] = (t[i, n + c] + t[i, n + q]) % 2

    def cz(self, c, q):
        self.hadamard(q); self.cx(c, q); self.hadamard(q)

    def x(self, q):
        self.t[: 2 * self.n, 2 * self.n] = (self.t[: 2 * self.n, 2 * self.n] + self.t[: 2 * self.n, self.n + q]) % 2

    def z(self, q):
        self.t[: 2 * self.n, 2 * self.n] = (self.t[: 2 * self.n, 2 * self.n] + self.t[: 2 * self.n, q]) % 2

    def y(self, q):
        n = self.n
        self.t[: 2 * n, 2 * n] = (self.t[: 2 * n, 2 * n] + self.t[: 2 * n, q] + self.t[: 2 * n, n + q]) % 2

    @staticmethod
    def _row_phase(x1, z1, x2, z2):
        if x1 == 0 and z1 == 0:
            return 0
        if x1 == 0 and z1 == 1:
            return x2 * (1 - 2 * z2)
        if x1 == 1 and z1 == 0:
            return z2 * (2 * x2 - 1)
        return z2 - x2

    def row_sum(self, target, control):
        t, n = self.t, self.n
        tot = 2 * t[target, 2 * n] + 2 * t[control, 2 * n] + sum(
            self._row_phase(t[control, i], t[control, n + i], t[target, i], t[target, n + i]) for i in range(n))
        assert tot % 2 == 0
        t[target, 2 * n] = 0 if tot % 4 == 0 else 1
        t[target, : 2 * n] = (t[target, : 2 * n] + t[control, : 2 * n]) % 2

    def measure(self, q):
        t, n = self.t, self.n
        stab = np.nonzero(t[n:2 * n, q])[0]
        if len(stab) == 0:
            t[2 * n, :] = 0
            for i in range(n):
                if t[i, q] == 1:
                    self.row_sum(2 * n, n + i)
            return int(t[2 * n, 2 * n])
        p = n + int(stab[0])
        for i in range(2 * n):
            if t[i, q] == 1 and i != p and i != p - n:
                self.row_sum(i, p)
        t[p - n, :] = t[p, :]
        t[p, :] = 0
        t[p, n + q] = 1
        t[p, 2 * n] = 1          # the reference draws rand(Bool); reset! undoes either outcome
        return 1

    def reset(self, q):
        if self.measure(q):
            self.x(q)

    def apply(self, layer):
        for (kind, _, tg) in layer:
            q = [a - 1 for a in tg]
            if kind in ("CX", "CNOT"):
                self.cx(*q)
            elif kind == "CZ":
                self.cz(*q)
            elif kind == "H":
                self.hadamard(q[0])
            elif kind == "S":
                self.phase(q[0])
            elif kind == "X":
                self.x(q[0])
            elif kind == "Y":
                self.y(q[0])
            elif kind == "Z":
                self.z(q[0])
            elif kind == "R":
                self.reset(q[0])
            elif kind in ("I", "IM"):
                pass
            else:
                raise AssertionError(kind)


def literal_mapping(c, circuit_tuple, x0, z0):
    """calc_mapping (src/design.jl:538-611) for one Pauli: returns (final x, final z, sign, row, track)."""
    n = c.qubit_num
    t = Tableau(n)
    row = np.zeros(c.N, dtype=np.int64)
    sup = [q for q in range(n) if x0[q] or z0[q]]
    for q in sup:                                   # get_prep_layer + apply!: PX = H, PY = H S
        if x0[q]:
            t.hadamard(q)
            if z0[q]:
                t.phase(q)
    lead = n + sup[0]

    def assemble():
        for q in sup[1:]:
            t.row_sum(lead, n + q)

    def update(px, pz, layer):
        s = {q for q in range(n) if px[q] or pz[q]}
        for g in layer:
            tg = [a - 1 for a in g[2]]
            if not (s & set(tg)):
                continue
            if g[0] == "R":
                row[:] = 0
                idx = 1
            elif len(tg) == 2:
                idx = px[tg[0]] + 2 * px[tg[1]] + 4 * pz[tg[0]] + 8 * pz[tg[1]]
            else:
                idx = px[tg[0]] + 2 * pz[tg[0]]
            row[c.gate_offset[g] + idx - 1] += 1

    track = []
    for li in circuit_tuple:
        assemble()
        px, pz = t.t[lead, :n].copy(), t.t[lead, n:2 * n].copy()
        track.append((px | pz).astype(bool))
        update(px, pz, c.circuit[li])
        assemble()                                   # row_sum! twice restores the generators
        t.apply(c.circuit[li])
    assemble()
    fx, fz, fs = t.t[lead, :n].copy(), t.t[lead, n:2 * n].copy(), int(t.t[lead, 2 * n])
    track.append((fx | fz).astype(bool))
    for q in range(n):
        if fx[q] or fz[q]:
            kind = "MY" if (fx[q] and fz[q]) else ("MX" if fx[q] else "MZ")
            row[c.gate_offset[(kind, 0, (q + 1,))]] += 1
    return fx.astype(bool), fz.astype(bool), bool(fs), row, np.array(track)


def _circuit(kind):
    if kind == "example":
        circ, lt, n, par = cd.example_circuit()
        return cd.get_circuit(circ, lt, n, par, {"single_qubit": 29.0, "two_qubit": 29.0, "meas_reset": 660.0})
    if kind == "rotated3":
        return cd.get_circuit(*cd.rotated_planar_circuit(3))
    if kind == "unrotated2x3":
        return cd.get_circuit(*cd.unrotated_planar_circuit(2, 3))
    if kind == "rotated3_no_decoupling":
        return cd.get_circuit(*cd.rotated_planar_circuit(3, dynamically_decouple=False))
    if kind == "rotated3x4_standard_cx":
        return cd.get_circuit(*cd.rotated_planar_circuit(3, 4, check_type="standard", gate_type="cx"))
    raise AssertionError(kind)


@pytest.mark.parametrize("kind", ["example", "rotated3", "unrotated2x3", "rotated3_no_decoupling", "rotated3x4_standard_cx"])
def test_batched_conjugation_matches_literal_tableau(kind):
    c = _circuit(kind)
    rng = np.random.default_rng(5)
    tuples = cd.get_tuple_set(c)
    # long repeated tuples are cut to 5 periods to keep the literal tableau quick
    tuples = [t if len(t) <= 12 else t[: 5 * (2 if t[0] != t[1] else 1)] for t in tuples]
    for tup in tuples:
        gates = cd.get_gates([c.circuit[i] for i in tup])
        X0, Z0 = cd.get_pauli_prep_set(gates, c.qubit_num)
        ms = cd.calc_mappings(c, tup, X0, Z0)
        rows = ms.rows.toarray()
        pick = range(len(X0)) if kind == "example" else rng.choice(len(X0), size=12, replace=False)
        for p in pick:
            fx, fz, fs, row, track = literal_mapping(c, tup, X0[p], Z0[p])
            assert np.array_equal(fx, ms.final_x[p]) and np.array_equal(fz, ms.final_z[p]), (tup, p)
            assert fs == bool(ms.final_sign[p]), (tup, p)
            assert np.array_equal(row, rows[p]), (tup, p)
            assert np.array_equal(track, ms.track[:, p, :]), (tup, p)


def test_covariance_paulis_match_literal_tableau():
    c = _circuit("rotated3")
    tup = [1]                                        # the first CZ layer
    X0, Z0 = cd.get_pauli_prep_set(cd.get_gates([c.circuit[1]]), c.qubit_num)
    ms = cd.calc_mappings(c, tup, X0, Z0)
    exps = cd.calc_experiment_set(ms, cd.calc_consistency(ms))
    keys, counts, fx, fz, fs, rows = cd.calc_covariance_dict(c, tup, ms, exps, {})
    assert len(keys) == 84
    rows = rows.toarray()
    for k in range(0, len(keys), 7):
        p, q = keys[k]
        lx, lz, ls, lrow, _ = literal_mapping(c, tup, X0[p] ^ X0[q], Z0[p] ^ Z0[q])
        assert np.array_equal(lx, fx[k]) and np.array_equal(lz, fz[k]) and ls == bool(fs[k])
        assert np.array_equal(lrow, rows[k])
        # non-trivial: the product's row differs from the sum of the two rows (src/design.jl:868-872)
        assert not np.array_equal(lrow, (ms.rows[p] + ms.rows[q]).toarray()[0])
        assert counts[k] == sum(1 for e in exps if p in e and q in e)


def test_example_circuit_hand_checked_rows():
    """SURVEY 8c item 8.  Gates in label order: layer 1 = CZ(2,3), I(1); layer 2 = CZ(1,2), H(3);
    layer 3 = H(1), H(3), S(2); then MZ/MX/MY per qubit: N = 15+3 + 15+3 + 3+3+3 + 9 = 54."""
    c = _circuit("example")
    assert c.N == 54
    assert c.unique_layer_indices == [0, 1, 2]
    fd = aces_b200.example_design()
    assert fd.matrix.shape == (99, 54) and fd.tuple_number == 7
    assert list(fd.mapping_lengths) == [9, 18, 18, 9, 18, 18, 9]
    assert list(fd.experiment_numbers) == [3, 9, 9, 3, 9, 9, 3]
    assert np.linalg.matrix_rank(fd.matrix.toarray().astype(float)) == 54
    A = fd.matrix.toarray()
    off = c.gate_offset
    # tuple 1 (empty circuit): Pauli p on qubit q is measured directly: one entry, M{p}_q.
    for q in range(3):
        for a, kind in enumerate(("MX", "MZ", "MY")):      # prep order X, Z, Y (src/design.jl:359-360)
            want = np.zeros(54, dtype=int)
            want[off[(kind, 0, (q + 1,))]] = 1
            assert np.array_equal(A[3 * q + a], want)
    # tuple 2 = layer 1.  X on qubit 2: CZ(2,3) sees XI (index 1), maps it to X_2 Z_3: MX_2, MZ_3.
    base = 9
    want = np.zeros(54, dtype=int)
    want[off[("CZ", 1, (2, 3))] + 1 - 1] = 1
    want[off[("MX", 0, (2,))]] = 1
    want[off[("MZ", 0, (3,))]] = 1
    assert np.array_equal(A[base + 3 * 1 + 0], want)
    # Z on qubit 1 only meets the padded identity: I(1) index Z = 2, then MZ_1.
    want = np.zeros(54, dtype=int)
    want[off[("I", 1, (1,))] + 2 - 1] = 1
    want[off[("MZ", 0, (1,))]] = 1
    assert np.array_equal(A[base + 3 * 0 + 1], want)
    # weight-2 Pauli Y_2 Y_3 (signature 9 of src/design.jl:361-372) on the gate support: index 15;
    # CZ maps YY -> (Y Z)(Z Y) up to sign = X_2 X_3: MX_2, MX_3.
    want = np.zeros(54, dtype=int)
    want[off[("CZ", 1, (2, 3))] + 15 - 1] = 1
    want[off[("MX", 0, (2,))]] = 1
    want[off[("MX", 0, (3,))]] = 1
    assert np.array_equal(A[base + 9 + 8], want)
    assert list(fd.meas_indices[1][9 + 8]) == [2, 3]
    # tuple 5 = layer 1 repeated three times (src/tuples.jl: repeat numbers are odd, at least 3)
    tuples = cd.get_tuple_set(c)
    assert tuples[4] == [0, 0, 0] or tuples[4] == [0] * len(tuples[4])
    r = len(tuples[4])
    row = A[int(np.sum(fd.mapping_lengths[:4])) + 3 * 1 + 0]     # X on qubit 2 again
    # X_2 -> X_2 Z_3 -> X_2 -> ...: the CZ sees XI (1) and XZ (9) alternately
    assert row[off[("CZ", 1, (2, 3))] + 1 - 1] == (r + 1) // 2
    assert row[off[("CZ", 1, (2, 3))] + 9 - 1] == r // 2


@pytest.mark.parametrize("d", [3, 5])
def test_rotated_design_sizes_and_structure(d):
    """n = 2d^2-1, N = 88d^2-36d-25, M = 168d^2-72d-48, T = 16."""
    fd = aces_b200.rotated_design(d, full_covariance=True)
    n = 2 * d * d - 1
    assert fd.n_qubits == n and fd.tuple_number == 16
    assert fd.matrix.shape == (168 * d * d - 72 * d - 48, 88 * d * d - 36 * d - 25)
    assert list(fd.experiment_numbers) == [3, 3, 9, 3, 9, 3, 9, 9, 3, 3, 3, 3, 9, 9, 9, 9]
    assert abs(fd.shot_weights.sum() - 1.0) < 1e-12 and np.all(fd.shot_weights > 0)
    if d == 3:
        assert np.linalg.matrix_rank(fd.matrix.toarray().astype(float)) == fd.matrix.shape[1]
    for t in range(fd.tuple_number):
        m = int(fd.mapping_lengths[t])
        seen = np.zeros(m, dtype=int)
        for exp in fd.experiment_ensemble[t]:
            assert np.all(np.diff(exp) > 0)
            seen[exp] += 1
            # simultaneously measurable: one measurement basis per qubit, so no two measured Paulis
            # of an experiment overlap on a qubit with different letters; here: supports are sets of
            # qubits and every qubit is read once, so each Pauli's parity is well defined
            assert all(len(fd.meas_indices[t][p]) in (1, 2) for p in exp)
        assert np.array_equal(seen, fd.diag_counts[t]) and seen.min() >= 1
        keys = fd.cov_keys[t]
        for k, (p, q) in enumerate(keys):
            assert p < q
            both = sum(1 for e in fd.experiment_ensemble[t] if p in e and q in e)
            assert both == fd.cov_counts[t][k] >= 1
            # conjugation is a homomorphism: the product's final Pauli is the product of the finals
            sp_, sq_ = set(fd.meas_indices[t][p]), set(fd.meas_indices[t][q])
            assert set(fd.cov_meas_indices[t][k]) == sp_ ^ sq_
            assert bool(fd.cov_signs[t][k]) == (bool(fd.pauli_signs[t][p]) ^ bool(fd.pauli_signs[t][q]))
        for j, ce in enumerate(fd.cov_experiment_ensemble[t]):
            exp = set(int(p) for p in fd.experiment_ensemble[t][j])
            assert all(int(keys[k][0]) in exp and int(keys[k][1]) in exp for k in ce)
    # repeat numbers 183 / 143 / 63 at the default noise (src/tuples.jl:228-236) show up as entries
    assert {1, 63, 143, 183} <= set(np.unique(fd.matrix.data).tolist())


def test_unrotated_design_sizes():
    fd = aces_b200.unrotated_design(3, full_covariance=True)
    assert fd.n_qubits == 25 and fd.tuple_number == 12
    assert list(fd.experiment_numbers) == [3, 3, 9, 9, 9, 9, 3, 3, 9, 9, 9, 9]
    assert np.linalg.matrix_rank(fd.matrix.toarray().astype(float)) == fd.matrix.shape[1]


def test_real_tables_through_the_parity_oracle():
    """The flattened tables of a real design drive the C oracle exactly like the synthetic ones:
    odd-parity counts equal a numpy brute force over the unpacked record bits."""
    import oracle_bindings as ob

    fd = aces_b200.rotated_design(3, full_covariance=True)
    exp_ptr, pauli_ptr, bits, n_eig = aces_b200.build_pauli_tables(fd)
    rng = np.random.default_rng(9)
    rows, B = 1000, fd.n_bytes
    e = 20                                                   # an experiment of a two-qubit tuple
    rec = np.asfortranarray(rng.integers(0, 256, size=(rows, B), dtype=np.uint8))
    unpacked = np.unpackbits(rec, axis=1, bitorder="little")
    lists = [bits[pauli_ptr[k]:pauli_ptr[k + 1]] for k in range(exp_ptr[e], exp_ptr[e + 1])]
    want = np.array([int(np.sum(np.bitwise_xor.reduce(unpacked[:, l], axis=1))) for l in lists])
    # pauli_estimate (src/estimate.jl:239-254) returns shots - 2 * odd, from 1-based qubit lists
    got = np.array([ob.pauli_estimate(rec, np.asarray(l) + 1, rows) for l in lists])
    assert np.array_equal(got, rows - 2 * want)


@pytest.mark.parametrize("kind", ["rotated3_no_decoupling", "rotated3x4_standard_cx"])
def test_other_circuit_variants_give_full_rank_designs(kind):
    """The non-default circuit variants of src/circuits/rotated_planar.jl:306-343 (no dynamical
    decoupling; standard checks with CX gates) go through the same generator."""
    fd = cd.generate_design(_circuit(kind), full_covariance=True)
    A = fd.matrix.toarray().astype(float)
    assert np.linalg.matrix_rank(A) == A.shape[1]
    assert abs(fd.shot_weights.sum() - 1.0) < 1e-12
    for t in range(fd.tuple_number):
        seen = np.zeros(int(fd.mapping_lengths[t]), dtype=int)
        for exp in fd.experiment_ensemble[t]:
            seen[exp] += 1
        assert seen.min() >= 1


@pytest.mark.parametrize("kind,d,shape,tuples", [("rotated", 9, (12912, 6779), 16), ("unrotated", 7, (11700, 6189), 12)])
def test_headline_design_sizes(kind, d, shape, tuples):
    """The BASELINE.json shapes from the generator, diagonal covariance only."""
    fd = (aces_b200.rotated_design if kind == "rotated" else aces_b200.unrotated_design)(d, full_covariance=False)
    assert fd.matrix.shape == shape and fd.tuple_number == tuples
    assert fd.n_qubits == (2 * d * d - 1 if kind == "rotated" else (2 * d - 1) ** 2)
    sizes = [len(e) for t in range(fd.tuple_number) for e in fd.experiment_ensemble[t]]
    assert max(sizes) <= 256          # one 8-round column group per experiment in the K1 kernel
    assert all(len(s) in (1, 2) for t in range(fd.tuple_number) for s in fd.meas_indices[t])
