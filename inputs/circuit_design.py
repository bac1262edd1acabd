"""Real ACES designs for the circuits named in BASELINE.json, as `FlatDesign` tables.

Host-side INPUT GENERATOR (numpy), not part of the device path: the reference builds a `Design`
with `generate_design` (src/design.jl:978-1100), which is outside the hot path (SURVEY.md
section 8, "not replaced") and cannot run here (no Julia).  To feed the parity and fit kernels
with the real structure instead of look-alike synthetic tables, this module restates the design
generator in array form:

* circuits: `rotated_planar_circuit` (src/circuits/rotated_planar.jl:186-368),
  `unrotated_planar_circuit` (src/circuits/unrotated_planar.jl:150-300), and the 3-qubit example of
  test/creation_tests.jl:63-88;
* labelling and gate-eigenvalue indexing: `label_circuit` (src/circuit.jl:293-337),
  `get_total_gates` (:353-386), `get_gate_data` (src/noise.jl:131-260);
* tuple sets: `get_basic_tuple_set`, `get_tuple_set_data`, `get_tuple_set`,
  `get_tuple_set_params` (src/tuples.jl:42-133,194-348,432-447);
* Pauli propagation and design rows: `get_pauli_prep_set` (src/design.jl:357-412),
  `update_design_row!` (:475-531), `calc_mapping` (:538-611) with the gate actions of
  src/tableau.jl:141-270 -- all Paulis of a tuple are conjugated together as rows of one bit
  matrix (the reference runs one tableau per Pauli; conjugation of the tracked stabiliser product is
  the same map);
* experiment packing: `calc_consistency_set` (:642-699), `calc_experiment_set` (:706-777);
* covariance dictionary: `calc_covariance_dict` (:785-896).

Indices are 0-based internally; `FlatDesign.meas_indices` keeps the reference's 1-based qubits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from aces_b200.design import FlatDesign, get_covariance_experiment_ensemble

Gate = Tuple[str, int, Tuple[int, ...]]   # (type, index, 1-based targets), src/tableau.jl:88-92

_PAULI_ROT = ["XX", "XZ", "XY", "ZX", "ZZ", "ZY", "YX", "YZ", "YY"]
ONE_QUBIT_GATES = {"I", "X", "Y", "Z", "H", "S", "S_DAG", "IM", "M", "R"}          # tableau.jl:619-622
TWO_QUBIT_GATES = ({"II", "CX", "CNOT", "CZ"} | set(_PAULI_ROT)
                   | {"SQRT_" + p for p in _PAULI_ROT} | {"SQRT_" + p + "_DAG" for p in _PAULI_ROT})
_SUPPORTED_ACTIONS = {"I", "IM", "II", "X", "Y", "Z", "H", "S", "S_DAG", "CX", "CNOT", "CZ", "R"}


def _gate_key(g: Gate):
    # Base.isless(::Gate, ::Gate), src/tableau.jl:99-101: [type; index; targets] lexicographic
    return (g[0], g[1], tuple(g[2]))


# ---------------------------------------------------------------------------------------------
# layers and circuits
# ---------------------------------------------------------------------------------------------

def make_layer(gate_types, ranges, n: int) -> List[Gate]:
    """The three `make_layer` methods of src/tableau.jl:568-614."""
    if isinstance(gate_types, str):
        if len(ranges) and isinstance(ranges[0], (list, tuple, np.ndarray)):
            layer = [(gate_types, 0, tuple(int(q) for q in r)) for r in ranges]
        else:
            layer = [(gate_types, 0, (int(q),)) for q in ranges]
    else:
        assert len(gate_types) == len(ranges)
        layer = [(gt, 0, (int(q),)) for gt, rng in zip(gate_types, ranges) for q in rng]
    _check_layer(layer, n)
    return layer


def _check_layer(layer: List[Gate], n: int) -> None:
    targets = sorted(q for g in layer for q in g[2])
    assert all(1 <= q <= n for q in targets), "layer acts on a qubit that is out of bounds"
    assert len(set(targets)) == len(targets), "layer acts on the same target more than once"


def pad_layer(layer: List[Gate], n: int) -> List[Gate]:
    """src/tableau.jl:790-806: identity gates ("IM" next to mid-circuit measure/reset) appended
    for the idle qubits in ascending order."""
    id_type = "IM" if any(g[0] in ("M", "R") for g in layer) else "I"
    used = {q for g in layer for q in g[2]}
    return list(layer) + [(id_type, 0, (q,)) for q in range(1, n + 1) if q not in used]


def get_layer_times(layer_types: Sequence[str], layer_time_dict: Dict[str, float]) -> np.ndarray:
    """src/circuit.jl:237-257: one time per layer, then the measurement+reset time."""
    return np.array([layer_time_dict[t] for t in layer_types] + [layer_time_dict["meas_reset"]], dtype=np.float64)


_DEFAULT_TIMES = {"single_qubit": 29.0, "two_qubit": 29.0, "meas_reset": 660.0, "dynamical": 29.0, "mid_reset": 660.0}


def rotated_planar_circuit(v: int, h: Optional[int] = None, check_type: str = "xzzx", gate_type: str = "cz",
                           dynamically_decouple: bool = True, ancilla_measurement: bool = True,
                           pad_identity: bool = True):
    """src/circuits/rotated_planar.jl:186-368 (defaults of get_rotated_param, :148-165)."""
    h = v if h is None else h
    assert v >= 3 and h >= 3
    data = [(2 * i, 2 * j) for j in range(1, h + 1) for i in range(1, v + 1)]
    inner = [(2 * i + 1, 2 * j + 1) for j in range(1, h) for i in range(1, v)]
    boundary = ([(2 * i + 1, 1) for i in range(1, v)] + [(2 * v + 1, 2 * j + 1) for j in range(1, h)]
                + [(2 * v + 1 - 2 * i, 2 * h + 1) for i in range(1, v)] + [(1, 2 * h + 1 - 2 * j) for j in range(1, h)])
    ancilla = inner + boundary[0::2]
    qubits = data + ancilla
    n = len(qubits)
    assert n == 2 * v * h - 1
    inv = {q: i + 1 for i, q in enumerate(qubits)}
    data_idx = list(range(1, len(data) + 1))
    anc_idx = list(range(len(data) + 1, n + 1))
    anc_x = [i for i in anc_idx if sum(qubits[i - 1]) % 4 == 0]
    anc_z = [i for i in anc_idx if sum(qubits[i - 1]) % 4 == 2]
    layers_x: List[List[List[int]]] = [[], [], [], []]
    for a in anc_x:
        (r, c) = qubits[a - 1]
        for li, adj in enumerate([(r - 1, c - 1), (r + 1, c - 1), (r - 1, c + 1), (r + 1, c + 1)]):
            if adj in inv:
                layers_x[li].append([a, inv[adj]])
    layers_z: List[List[List[int]]] = [[], [], [], []]
    for a in anc_z:
        (r, c) = qubits[a - 1]
        for li, adj in enumerate([(r - 1, c - 1), (r - 1, c + 1), (r + 1, c - 1), (r + 1, c + 1)]):
            if adj in inv:
                layers_z[li].append([inv[adj], a])
    S, T2, DD, MR = "single_qubit", "two_qubit", "dynamical", "mid_reset"
    if check_type == "xzzx" and gate_type == "cz":
        cz = [make_layer("CZ", layers_z[k] + layers_x[k], n) for k in range(4)]
        if dynamically_decouple:
            circuit = [make_layer(["X", "H"], [data_idx, anc_idx], n), cz[0],
                       make_layer(["H", "X"], [data_idx, anc_idx], n), cz[1],
                       make_layer(["X", "X"], [data_idx, anc_idx], n), cz[2],
                       make_layer(["H", "X"], [data_idx, anc_idx], n), cz[3],
                       make_layer(["X", "H"], [data_idx, anc_idx], n)]
            layer_types = [S, T2, S, T2, DD, T2, S, T2, S]
        else:
            circuit = [make_layer("H", anc_idx, n), cz[0], make_layer("H", data_idx, n), cz[1], cz[2],
                       make_layer("H", data_idx, n), cz[3], make_layer("H", anc_idx, n)]
            layer_types = [S, T2, S, T2, T2, S, T2, S]
    elif check_type == "standard" and gate_type == "cx":
        cx = [make_layer("CX", layers_z[k] + layers_x[k], n) for k in range(4)]
        circuit = [make_layer("H", anc_x, n)] + cx + [make_layer("H", anc_x, n)]
        layer_types = [S, T2, T2, T2, T2, S]
        dynamically_decouple = False
    else:
        raise ValueError("unsupported pairing of check type and gate type")
    if ancilla_measurement:
        circuit.append(make_layer("R", anc_idx, n))
        layer_types.append(MR)
    if pad_identity:
        circuit = [pad_layer(l, n) for l in circuit]
    return circuit, layer_types, n, {"dynamically_decouple": dynamically_decouple}


def unrotated_planar_circuit(v: int, h: Optional[int] = None, ancilla_measurement: bool = True,
                             pad_identity: bool = True):
    """src/circuits/unrotated_planar.jl:150-300 (gate type :cx, the only one supported there)."""
    h = v if h is None else h
    rows, cols = 2 * v - 1, 2 * h - 1
    data = [(i, j) for j in range(1, cols + 1) for i in range(1, rows + 1) if (i + j) % 2 == 0]
    ancilla = [(i, j) for j in range(1, cols + 1) for i in range(1, rows + 1) if (i + j) % 2 == 1]
    qubits = data + ancilla
    n = len(qubits)
    assert n == rows * cols
    inv = {q: i + 1 for i, q in enumerate(qubits)}
    anc_idx = list(range(len(data) + 1, n + 1))
    anc_x = [i for i in anc_idx if qubits[i - 1][0] % 2 == 0]
    anc_z = [i for i in anc_idx if qubits[i - 1][1] % 2 == 0]
    layers_x: List[List[List[int]]] = [[], [], [], []]
    for a in anc_x:
        (r, c) = qubits[a - 1]
        for li, adj in enumerate([(r - 1, c), (r, c + 1), (r, c - 1), (r + 1, c)]):
            if adj in inv:
                layers_x[li].append([a, inv[adj]])
    layers_z: List[List[List[int]]] = [[], [], [], []]
    for a in anc_z:
        (r, c) = qubits[a - 1]
        for li, adj in enumerate([(r - 1, c), (r, c - 1), (r, c + 1), (r + 1, c)]):
            if adj in inv:
                layers_z[li].append([inv[adj], a])
    S, T2, MR = "single_qubit", "two_qubit", "mid_reset"
    circuit = ([make_layer("H", anc_x, n)] + [make_layer("CX", layers_z[k] + layers_x[k], n) for k in range(4)]
               + [make_layer("H", anc_x, n)])
    layer_types = [S, T2, T2, T2, T2, S]
    if ancilla_measurement:
        circuit.append(make_layer("R", anc_idx, n))
        layer_types.append(MR)
    if pad_identity:
        circuit = [pad_layer(l, n) for l in circuit]
    return circuit, layer_types, n, {"dynamically_decouple": False}


def example_circuit(pad_identity: bool = True):
    """The 3-qubit example of test/creation_tests.jl:63-88."""
    n = 3
    circuit = [[("CZ", 0, (2, 3))], [("CZ", 0, (1, 2)), ("H", 0, (3,))],
               [("H", 0, (1,)), ("S", 0, (2,)), ("H", 0, (3,))]]
    layer_types = ["two_qubit", "two_qubit", "single_qubit"]
    if pad_identity:
        circuit = [pad_layer(l, n) for l in circuit]
    return circuit, layer_types, n, {"dynamically_decouple": False}


# ---------------------------------------------------------------------------------------------
# labelling and gate data
# ---------------------------------------------------------------------------------------------

def label_circuit(circuit: List[List[Gate]]):
    """src/circuit.jl:293-337: layers sorted, gates labelled by the ordinal of the unique layer
    they appear in.  Returns (labelled circuit, 0-based unique layer indices)."""
    unl = [sorted([(g[0], 0, tuple(g[2])) for g in layer], key=_gate_key) for layer in circuit]
    unique_layers: List[List[Gate]] = []
    unique_idx: List[int] = []
    for l, layer in enumerate(unl):
        if layer not in unique_layers:
            unique_layers.append(layer)
            unique_idx.append(l)
    which_unique = [unique_layers.index(layer) for layer in unl]
    appear: Dict[Gate, List[int]] = {}
    for u, layer in enumerate(unique_layers):
        for g in layer:
            appear.setdefault(g, []).append(u)
    labelled = [[(g[0], appear[g].index(which_unique[l]) + 1, g[2]) for g in layer] for l, layer in enumerate(unl)]
    return labelled, unique_idx


def get_gates(circuit: List[List[Gate]]) -> List[Gate]:
    """src/circuit.jl:276-288: unique gates in order of first appearance."""
    seen, out = set(), []
    for layer in circuit:
        for g in layer:
            if g not in seen:
                seen.add(g)
                out.append(g)
    return out


def is_additive(g: Gate) -> bool:
    """src/tableau.jl:736-743 with strict = false."""
    return g[0] in ("PZ", "PX", "PY", "MZ", "MX", "MY", "M", "R", "IM")


@dataclass
class Circuit:
    """The fields of `Circuit` (src/circuit.jl:25-42) the design generator reads."""
    circuit: List[List[Gate]]            # labelled layers
    layer_types: List[str]
    layer_times: np.ndarray              # per layer, then measurement+reset
    qubit_num: int
    unique_layer_indices: List[int]      # 0-based
    total_gates: List[Gate]
    gate_offset: Dict[Gate, int]         # get_gate_index_dict (src/noise.jl:271-277), 0-based column of pauli_idx 1 minus 1
    gate_size: Dict[Gate, int]
    N: int
    params: Dict[str, object]
    noise: Dict[str, float]
    name: str = ""


def get_circuit(circuit, layer_types, n, params, layer_time_dict=None, noise=None, name="") -> Circuit:
    """`get_circuit` / `prepare_circuit` (src/circuit.jl:514-636) with noisy_prep = false,
    noisy_meas = true, combined = false (the defaults)."""
    labelled, unique_idx = label_circuit(circuit)
    total = get_gates(labelled)
    for q in range(1, n + 1):
        total += [("MZ", 0, (q,)), ("MX", 0, (q,)), ("MY", 0, (q,))]
    assert len(set(total)) == len(total)
    offset, size, N = {}, {}, 0
    for g in total:
        k = 1 if g[0] in ("MZ", "MX", "MY", "PZ", "PX", "PY", "M", "R") else 4 ** len(g[2]) - 1
        offset[g], size[g] = N, k
        N += k
    times = get_layer_times(layer_types, layer_time_dict or _DEFAULT_TIMES)
    noise = noise or {"r_1": 0.05 / 100, "r_2": 0.4 / 100, "r_m": 0.8 / 100}
    return Circuit(labelled, list(layer_types), times, n, unique_idx, total, offset, size, N, dict(params), dict(noise), name)


def depolarising_gate_eigenvalues(c: Circuit) -> np.ndarray:
    """Gate eigenvalues of depolarising noise (src/noises/depolarising.jl: every non-identity
    Pauli error of a k-qubit gate has probability r/(4^k-1), so each eigenvalue is
    1 - r 4^k/(4^k-1); measurement, reset and measurement-idle flips have eigenvalue 1 - 2 r)."""
    r_1, r_2, r_m = c.noise["r_1"], c.noise["r_2"], c.noise["r_m"]
    r_im, r_r = c.noise.get("r_im", r_m), c.noise.get("r_r", r_m)
    g_eig = np.empty(c.N, dtype=np.float64)
    for g in c.total_gates:
        o, k = c.gate_offset[g], c.gate_size[g]
        if g[0] in ("MZ", "MX", "MY"):
            g_eig[o] = 1 - 2 * r_m
        elif g[0] == "R":
            g_eig[o] = 1 - 2 * r_r
        elif g[0] == "IM":
            g_eig[o:o + k] = 1 - r_im * 4 / 3
        elif len(g[2]) == 1:
            g_eig[o:o + k] = 1 - r_1 * 4 / 3
        else:
            g_eig[o:o + k] = 1 - r_2 * 16 / 15
    return g_eig


def lognormal_gate_eigenvalues(c: Circuit, seed: int = 0, std_log: float = 0.5) -> np.ndarray:
    """Log-normal Pauli noise in the spirit of src/noises/lognormal.jl:157-235: every Pauli error
    probability is an independent log-normal draw whose gate-level sum has mean r_1 / r_2 / r_m;
    eigenvalues by the Walsh-Hadamard transform with the symplectic sign (numpy's generator, not
    Julia's: a statistically equivalent input, not a bit-identical one)."""
    rng = np.random.default_rng(seed)
    r_1, r_2, r_m = c.noise["r_1"], c.noise["r_2"], c.noise["r_m"]
    g_eig = np.empty(c.N, dtype=np.float64)
    mean_factor = math.exp(-0.5 * std_log * std_log)
    for g in c.total_gates:
        o, k = c.gate_offset[g], c.gate_size[g]
        if k == 1:
            g_eig[o] = 1 - 2 * r_m * mean_factor * math.exp(std_log * rng.standard_normal())
            continue
        r = r_1 if len(g[2]) == 1 else r_2
        if g[0] == "IM":
            r = c.noise.get("r_im", r_m)
        nq = len(g[2])
        p = np.zeros(4 ** nq)
        p[1:] = r / k * mean_factor * np.exp(std_log * rng.standard_normal(k))
        p[0] = 1 - p[1:].sum()
        # index a = x_1 + 2 x_2 + 4 z_1 + 8 z_2 (src/design.jl:505-517); eigenvalue of Pauli a is
        # sum_b (-1)^{<a,b>} p_b with the symplectic product
        idx = np.arange(4 ** nq)
        xb = (idx[:, None] >> np.arange(nq)) & 1
        zb = (idx[:, None] >> (nq + np.arange(nq))) & 1
        sym = (xb @ zb.T + zb @ xb.T) & 1
        lam = ((1 - 2 * sym) * p[None, :]).sum(1)
        g_eig[o:o + k] = lam[1:]
    return g_eig


# ---------------------------------------------------------------------------------------------
# tuple sets
# ---------------------------------------------------------------------------------------------

def get_tuple_set(c: Circuit, error_target: float = 0.1, add_circuit: bool = False) -> List[List[int]]:
    """`get_tuple_set(get_tuple_set_data(c))` (src/tuples.jl:42-45,194-348,432-447); tuples hold
    0-based layer indices."""
    basic = [[]] + [[i] for i in c.unique_layer_indices]
    r_1, r_2, r_m = c.noise["r_1"], c.noise["r_2"], c.noise["r_m"]
    lt = c.layer_types
    types: List[str] = []
    for t in lt:
        if t != "mid_reset" and t not in types:
            types.append(t)
    decoupled = bool(c.params.get("dynamically_decouple", False))
    rep = {"single_qubit": math.ceil((error_target - r_m) / r_1),
           "two_qubit": math.ceil((3 * error_target - 2 * r_m) / ((r_2 + r_1) if decoupled else r_2)),
           "dynamical": math.ceil((0.8 * error_target - r_m) / r_1)}
    circuit_tuple = [c.unique_layer_indices[[c.circuit[u] for u in c.unique_layer_indices].index(c.circuit[i])]
                     for i in range(len(c.circuit)) if lt[i] != "mid_reset"]
    circuit_error = sum({"single_qubit": r_1, "two_qubit": r_2, "dynamical": r_1}[lt[i]] for i in circuit_tuple)
    rep["circuit"] = math.ceil((3 * error_target - 2 * r_m) / circuit_error)
    repeat_numbers = []
    for t in types:
        k = rep[t]
        k = (k - 1 if k % 2 == 0 else k) if k >= 3 else 3
        repeat_numbers.append(k)
    if decoupled:
        by_type = [[i for i in range(len(lt)) if lt[i] == t and i in c.unique_layer_indices] for t in types]
        non_two = [i for t, ids in zip(types, by_type) if t != "two_qubit" for i in ids]
        two = [i for t, ids in zip(types, by_type) if t == "two_qubit" for i in ids]
        dyn = [i for t, ids in zip(types, by_type) if t == "dynamical" for i in ids]
        pairs = [(i, j) for j in dyn for i in two]
        repeat_tuples = [[i] for i in non_two] + [[i, j] for (i, j) in pairs]
        repeat_type = [types.index(lt[i]) for i in non_two] + [types.index(lt[i]) for (i, _) in pairs]
    else:
        ids = [i for i in c.unique_layer_indices if lt[i] in types]
        repeat_tuples = [[i] for i in ids]
        repeat_type = [types.index(lt[i]) for i in ids]
    if add_circuit:
        k = rep["circuit"]
        k = (k - 1 if k % 2 == 1 else k) if k >= 2 else 2
        repeat_numbers.append(k)
        repeat_tuples.append(circuit_tuple)
        repeat_type.append(len(repeat_numbers) - 1)
    tuple_set = sorted(basic, key=len)
    for tup, ti in zip(repeat_tuples, repeat_type):
        if repeat_numbers[ti] > 0:
            tuple_set.append(list(tup) * repeat_numbers[ti])
    out: List[List[int]] = []
    for t in tuple_set:
        if t not in out:
            out.append(t)
    for t in out:   # check_tuple!, src/tuples.jl:143-152
        kinds = {lt[i] for i in t}
        assert not ("mid_reset" in kinds and "two_qubit" in kinds)
    return out


def get_tuple_set_params(c: Circuit, tuple_set, experiment_numbers, weight_experiments: bool = True):
    """src/tuples.jl:75-133: tuple times in units of the harmonic mean of the basic experiment
    times, and the default shot weights."""
    meas_reset = c.layer_times[-1]
    times = c.layer_times[:-1]
    unnorm = np.array([meas_reset + sum(times[i] for i in t) for t in tuple_set], dtype=np.float64)
    basic = [[]] + [[i] for i in c.unique_layer_indices]
    basic_times = np.array([meas_reset + sum(times[i] for i in t) for t in basic], dtype=np.float64)
    basic_exp = np.array([3] + [3 ** max(len(g[2]) for g in c.circuit[i]) for i in c.unique_layer_indices], dtype=np.float64)
    harm = basic_exp.sum() / (basic_exp / basic_times).sum()
    tuple_times = unnorm / harm
    X = np.asarray(experiment_numbers, dtype=np.float64)
    w = (X / tuple_times) / (X / tuple_times).sum() if weight_experiments else (1 / tuple_times) / (1 / tuple_times).sum()
    return tuple_times, w


# ---------------------------------------------------------------------------------------------
# Pauli propagation
# ---------------------------------------------------------------------------------------------

_W1 = [(1, 0), (0, 1), (1, 1)]                                              # X, Z, Y
_W2 = [(1, 1, 0, 0), (0, 1, 1, 0), (1, 1, 1, 0), (1, 0, 0, 1), (0, 0, 1, 1),  # src/design.jl:361-372
       (1, 0, 1, 1), (1, 1, 0, 1), (0, 1, 1, 1), (1, 1, 1, 1)]


def get_pauli_prep_set(gates: List[Gate], n: int):
    """src/design.jl:357-412 as bit matrices X, Z of shape [3n + 9 l_2, n]."""
    sup2 = sorted({tuple(sorted(g[2])) for g in gates if len(g[2]) == 2})
    assert all(len(g[2]) in (1, 2) for g in gates)
    P = 3 * n + 9 * len(sup2)
    X = np.zeros((P, n), dtype=bool)
    Z = np.zeros((P, n), dtype=bool)
    for i in range(n):
        for a, (x, z) in enumerate(_W1):
            X[3 * i + a, i], Z[3 * i + a, i] = x, z
    for i, (q1, q2) in enumerate(sup2):
        for a, (x1, x2, z1, z2) in enumerate(_W2):
            r = 3 * n + 9 * i + a
            X[r, q1 - 1], X[r, q2 - 1], Z[r, q1 - 1], Z[r, q2 - 1] = x1, x2, z1, z2
    return X, Z


class _LayerPlan:
    """A layer split by gate class into index arrays, built once per distinct layer."""

    def __init__(self, layer: List[Gate], c: Circuit):
        one = [g for g in layer if len(g[2]) == 1 and g[0] != "R"]
        two = [g for g in layer if len(g[2]) == 2]
        rst = [g for g in layer if g[0] == "R"]
        for g in layer:
            if g[0] not in _SUPPORTED_ACTIONS:
                raise NotImplementedError(f"gate {g[0]} is not supported by this generator")
        self.one_t = np.array([g[2][0] - 1 for g in one], dtype=np.int64)
        self.one_off = np.array([c.gate_offset[g] for g in one], dtype=np.int64)
        self.two_t1 = np.array([g[2][0] - 1 for g in two], dtype=np.int64)
        self.two_t2 = np.array([g[2][1] - 1 for g in two], dtype=np.int64)
        self.two_off = np.array([c.gate_offset[g] for g in two], dtype=np.int64)
        self.rst_t = np.array([g[2][0] - 1 for g in rst], dtype=np.int64)
        self.rst_off = np.array([c.gate_offset[g] for g in rst], dtype=np.int64)
        self.act: Dict[str, np.ndarray] = {}
        for kind in ("H", "S", "S_DAG", "X", "Y", "Z"):
            self.act[kind] = np.array([g[2][0] - 1 for g in layer if g[0] == kind], dtype=np.int64)
        self.cx = np.array([[g[2][0] - 1, g[2][1] - 1] for g in layer if g[0] in ("CX", "CNOT")], dtype=np.int64).reshape(-1, 2)
        self.cz = np.array([[g[2][0] - 1, g[2][1] - 1] for g in layer if g[0] == "CZ"], dtype=np.int64).reshape(-1, 2)


def _parity(m: np.ndarray) -> np.ndarray:
    return (np.count_nonzero(m, axis=1) & 1).astype(bool)


def _hadamard(X, Z, s, t):
    s ^= _parity(X[:, t] & Z[:, t])          # src/tableau.jl:141-154
    tmp = X[:, t].copy()
    X[:, t] = Z[:, t]
    Z[:, t] = tmp


def _cx(X, Z, s, ctl, tgt):
    xc, zc, xt, zt = X[:, ctl], Z[:, ctl], X[:, tgt], Z[:, tgt]
    s ^= _parity(xc & zt & ~(xt ^ zc))       # src/tableau.jl:181-199
    X[:, tgt] = xt ^ xc
    Z[:, ctl] = zc ^ zt


def _apply_layer(plan: _LayerPlan, X, Z, s):
    """Conjugates every row Pauli by the layer (gate actions of src/tableau.jl:141-270, dispatch
    of apply! :409-556).  A mid-circuit reset maps a Pauli supported on the reset qubit to +Z
    there (measure! + x!, :353-403; such a Pauli has weight one because tuples never mix reset
    and two-qubit layers, src/tuples.jl:143-152)."""
    a = plan.act
    if len(a["H"]):
        _hadamard(X, Z, s, a["H"])
    for kind in ("S", "S_DAG"):
        t = a[kind]
        if len(t):
            s ^= _parity(X[:, t] & Z[:, t])  # phase!, :162-173
            Z[:, t] ^= X[:, t]
            if kind == "S_DAG":
                s ^= _parity(X[:, t])        # followed by z!
    if len(a["X"]):
        s ^= _parity(Z[:, a["X"]])
    if len(a["Z"]):
        s ^= _parity(X[:, a["Z"]])
    if len(a["Y"]):
        s ^= _parity(X[:, a["Y"]] ^ Z[:, a["Y"]])
    if len(plan.cx):
        _cx(X, Z, s, plan.cx[:, 0], plan.cx[:, 1])
    if len(plan.cz):
        _hadamard(X, Z, s, plan.cz[:, 1])    # cz! = H(target) CX H(target), :209-214
        _cx(X, Z, s, plan.cz[:, 0], plan.cz[:, 1])
        _hadamard(X, Z, s, plan.cz[:, 1])
    if len(plan.rst_t):
        t = plan.rst_t
        hit = (X[:, t] | Z[:, t]).any(axis=1)
        if hit.any():
            assert ((X[hit] | Z[hit]).sum(axis=1) == 1).all(), "reset of a Pauli of weight > 1"
            Z[hit] |= X[hit]
            X[hit] = False
            s[hit] = False


@dataclass
class MappingSet:
    """Array form of a vector of `Mapping`s (src/design.jl:45-60)."""
    init_x: np.ndarray
    init_z: np.ndarray
    final_x: np.ndarray
    final_z: np.ndarray
    final_sign: np.ndarray
    rows: sp.csr_matrix          # design rows, P x N
    track: np.ndarray            # spread_track as support bits, [L+1, P, n]


def calc_mappings(c: Circuit, circuit_tuple: List[int], X0: np.ndarray, Z0: np.ndarray,
                  plans: Optional[Dict[int, _LayerPlan]] = None, want_track: bool = True) -> MappingSet:
    """`calc_mapping` (src/design.jl:538-611) for all rows of (X0, Z0) at once; noisy_prep = false,
    noisy_meas = true."""
    P, n = X0.shape
    X, Z = X0.copy(), Z0.copy()
    s = np.zeros(P, dtype=bool)
    plans = {} if plans is None else plans
    ent_p: List[np.ndarray] = []
    ent_c: List[np.ndarray] = []
    ent_l: List[np.ndarray] = []
    last_reset = np.full(P, -1, dtype=np.int64)
    track = np.zeros((len(circuit_tuple) + 1, P, n), dtype=bool) if want_track else None
    for l, layer_idx in enumerate(circuit_tuple):
        if layer_idx not in plans:
            plans[layer_idx] = _LayerPlan(c.circuit[layer_idx], c)
        plan = plans[layer_idx]
        if want_track:
            track[l] = X | Z
        # update_design_row!, src/design.jl:475-531
        if len(plan.one_t):
            idx = X[:, plan.one_t].astype(np.int64) + 2 * Z[:, plan.one_t]
            p, g = np.nonzero(idx)
            ent_p.append(p); ent_c.append(plan.one_off[g] + idx[p, g] - 1); ent_l.append(np.full(len(p), l))
        if len(plan.two_t1):
            idx = (X[:, plan.two_t1].astype(np.int64) + 2 * X[:, plan.two_t2]
                   + 4 * Z[:, plan.two_t1] + 8 * Z[:, plan.two_t2])
            p, g = np.nonzero(idx)
            ent_p.append(p); ent_c.append(plan.two_off[g] + idx[p, g] - 1); ent_l.append(np.full(len(p), l))
        if len(plan.rst_t):
            hit = X[:, plan.rst_t] | Z[:, plan.rst_t]
            p, g = np.nonzero(hit)
            last_reset[p] = l                 # the row is zeroed, then the reset's own variable counted
            ent_p.append(p); ent_c.append(plan.rst_off[g]); ent_l.append(np.full(len(p), l))
        _apply_layer(plan, X, Z, s)
    if want_track:
        track[len(circuit_tuple)] = X | Z
    # measurement layer (get_meas_layer :447-469; SPAM gates take pauli_idx 1, :485-496)
    mz = np.array([c.gate_offset[("MZ", 0, (q,))] for q in range(1, n + 1)], dtype=np.int64)
    mx = np.array([c.gate_offset[("MX", 0, (q,))] for q in range(1, n + 1)], dtype=np.int64)
    my = np.array([c.gate_offset[("MY", 0, (q,))] for q in range(1, n + 1)], dtype=np.int64)
    p, q = np.nonzero(X | Z)
    col = np.where(X[p, q] & Z[p, q], my[q], np.where(X[p, q], mx[q], mz[q]))
    ent_p.append(p); ent_c.append(col); ent_l.append(np.full(len(p), len(circuit_tuple)))
    pp = np.concatenate(ent_p); cc = np.concatenate(ent_c); ll = np.concatenate(ent_l)
    keep = ll >= last_reset[pp]
    rows = sp.csr_matrix((np.ones(int(keep.sum()), dtype=np.int32), (pp[keep], cc[keep])), shape=(P, c.N))
    rows.sum_duplicates()
    rows.sort_indices()
    return MappingSet(X0, Z0, X, Z, s, rows, track)


# ---------------------------------------------------------------------------------------------
# experiments and covariance dictionary
# ---------------------------------------------------------------------------------------------

def _conflicts(X: np.ndarray, Z: np.ndarray) -> np.ndarray:
    """conflict[i, j]: some qubit carries different non-identity Paulis in rows i and j."""
    code = X.astype(np.int8) + 2 * Z.astype(np.int8)
    nz = (code > 0).astype(np.float32)
    both = nz @ nz.T
    same = np.zeros_like(both)
    for a in (1, 2, 3):
        o = (code == a).astype(np.float32)
        same += o @ o.T
    return (both - same) > 0.5


def calc_consistency(ms: MappingSet) -> np.ndarray:
    """src/design.jl:642-699 as a boolean matrix (diagonal true)."""
    ok = ~(_conflicts(ms.init_x, ms.init_z) | _conflicts(ms.final_x, ms.final_z))
    np.fill_diagonal(ok, True)
    return ok


def calc_experiment_set(ms: MappingSet, consistent: np.ndarray) -> List[np.ndarray]:
    """The greedy packing of src/design.jl:706-777 (stable order by final weight, most overlap
    first, first index on ties), then every other compatible Pauli re-added."""
    sup = ms.final_x | ms.final_z
    L = sup.shape[0]
    weight = sup.sum(axis=1)
    total = np.argsort(-weight, kind="stable")
    sup_t = sup[total].astype(np.int32)          # everything below is indexed by position in `total`
    cons_t = consistent[np.ix_(total, total)]
    unadded = np.ones(L, dtype=bool)
    experiments = []
    for first in range(L):
        if not unadded[first]:
            continue
        # `first` is the first unadded Pauli in the order `total`
        experiment = [first]
        unadded[first] = False
        in_exp = np.zeros(L, dtype=bool)
        in_exp[first] = True
        set_sup = sup_t[first] > 0
        overlap = sup_t[:, set_sup].sum(axis=1)   # |final support of candidate  n  support of the set|
        compatible = cons_t[first].copy()         # consistent with every Pauli of the experiment
        for phase in (0, 1):
            # phase 0: unadded Paulis; phase 1: all Paulis not in the experiment
            while True:
                cand = compatible & (unadded if phase == 0 else ~in_exp)
                if not cand.any():
                    break
                best = int(np.argmax(np.where(cand, overlap, -1)))   # most overlap, first in order on ties
                experiment.append(best)
                in_exp[best] = True
                if phase == 0:
                    unadded[best] = False
                new = (sup_t[best] > 0) & ~set_sup
                if new.any():
                    overlap += sup_t[:, new].sum(axis=1)
                    set_sup |= new
                compatible &= cons_t[best]
        experiments.append(np.sort(total[np.array(experiment, dtype=np.int64)]).astype(np.int64))
    return experiments


def calc_covariance_dict(c: Circuit, circuit_tuple, ms: MappingSet, experiments, plans):
    """Off-diagonal part of src/design.jl:785-896: for every pair of Paulis measured together
    whose spread tracks intersect, the mapping of their product; kept (with the number of
    experiments that measure both) when its design row differs from the sum of the two rows.
    Returns (keys [C,2] in Julia's CartesianIndex order, counts, final x, final z, final sign, rows)."""
    P = ms.init_x.shape[0]
    found = []
    for exp in experiments:
        sub = ms.track[:, exp, :].astype(np.float32)
        inter = np.zeros((len(exp), len(exp)), dtype=bool)
        for l in range(sub.shape[0]):
            inter |= (sub[l] @ sub[l].T) > 0.5
        i1, i2 = np.nonzero(np.triu(inter, k=1))
        found.append(exp[i2] * P + exp[i1])       # Julia's CartesianIndex order: second index major
    flat, counts = np.unique(np.concatenate(found) if found else np.zeros(0, dtype=np.int64), return_counts=True)
    if len(flat) == 0:
        z = np.zeros((0, ms.init_x.shape[1]), dtype=bool)
        return (np.zeros((0, 2), dtype=np.int64), np.zeros(0, dtype=np.int64), z, z, np.zeros(0, dtype=bool),
                sp.csr_matrix((0, c.N), dtype=np.int32))
    keys = np.stack([flat % P, flat // P], axis=1).astype(np.int64)
    counts = counts.astype(np.int64)
    # Pauli.+ (src/design.jl:26-37): bitwise sum; initial signs are all positive
    keep_all, fx, fz, fs, rows = [], [], [], [], []
    chunk = max(1, 40_000_000 // max(1, (len(circuit_tuple) + 2) * ms.init_x.shape[1]))
    for lo in range(0, len(keys), chunk):
        kk = keys[lo:lo + chunk]
        sx = ms.init_x[kk[:, 0]] ^ ms.init_x[kk[:, 1]]
        sz = ms.init_z[kk[:, 0]] ^ ms.init_z[kk[:, 1]]
        assert (sx | sz).any(axis=1).all()
        sm = calc_mappings(c, circuit_tuple, sx, sz, plans, want_track=False)
        diff = sm.rows - (ms.rows[kk[:, 0]] + ms.rows[kk[:, 1]])
        diff.eliminate_zeros()
        nontrivial = np.diff(diff.indptr) > 0
        keep_all.append(nontrivial)
        fx.append(sm.final_x[nontrivial]); fz.append(sm.final_z[nontrivial]); fs.append(sm.final_sign[nontrivial])
        rows.append(sm.rows[np.nonzero(nontrivial)[0]])
    keep = np.concatenate(keep_all)
    return (keys[keep], counts[keep], np.concatenate(fx), np.concatenate(fz), np.concatenate(fs),
            sp.vstack(rows).tocsr())


# ---------------------------------------------------------------------------------------------
# generate_design
# ---------------------------------------------------------------------------------------------

def generate_design(c: Circuit, tuple_set: Optional[List[List[int]]] = None, full_covariance: bool = True,
                    gate_eigenvalues: Optional[np.ndarray] = None, weight_experiments: bool = True) -> FlatDesign:
    """`generate_design` (src/design.jl:978-1100) emitting the flat tables of the estimation path."""
    tuple_set = get_tuple_set(c) if tuple_set is None else tuple_set
    n = c.qubit_num
    plans: Dict[int, _LayerPlan] = {}
    blocks, m_t, X_t = [], [], []
    meas_indices, signs, experiments_all, diag_counts = [], [], [], []
    cov_keys, cov_meas, cov_signs, cov_counts, cov_rows = [], [], [], [], []
    for circuit_tuple in tuple_set:
        tuple_gates = get_gates([c.circuit[i] for i in circuit_tuple])     # apply_tuple, src/circuit.jl:145-154
        X0, Z0 = get_pauli_prep_set(tuple_gates, n)
        ms = calc_mappings(c, circuit_tuple, X0, Z0, plans, want_track=full_covariance)
        experiments = calc_experiment_set(ms, calc_consistency(ms))
        blocks.append(ms.rows)
        m_t.append(X0.shape[0])
        X_t.append(len(experiments))
        fsup = ms.final_x | ms.final_z
        meas_indices.append([np.nonzero(fsup[p])[0].astype(np.int64) + 1 for p in range(X0.shape[0])])
        signs.append(ms.final_sign.copy())
        experiments_all.append(experiments)
        dc = np.zeros(X0.shape[0], dtype=np.int64)
        for e in experiments:
            dc[e] += 1
        assert dc.min() >= 1
        diag_counts.append(dc)
        if full_covariance:
            keys, counts, fx, fz, fs, rows = calc_covariance_dict(c, circuit_tuple, ms, experiments, plans)
            cov_keys.append(keys)
            cov_counts.append(counts)
            csup = fx | fz
            cov_meas.append([np.nonzero(csup[k])[0].astype(np.int64) + 1 for k in range(len(keys))])
            cov_signs.append(fs)
            cov_rows.append(rows.astype(np.int64))
    A = sp.vstack(blocks).astype(np.int32).tocsc()
    A.sort_indices()
    tuple_times, shot_weights = get_tuple_set_params(c, tuple_set, X_t, weight_experiments)
    fd = FlatDesign(
        n_qubits=n,
        experiment_numbers=np.asarray(X_t, dtype=np.int64),
        shot_weights=shot_weights,
        tuple_times=tuple_times,
        mapping_lengths=np.asarray(m_t, dtype=np.int64),
        meas_indices=meas_indices,
        pauli_signs=signs,
        experiment_ensemble=experiments_all,
        full_covariance=full_covariance,
        diag_counts=diag_counts,
        matrix=A,
        gate_eigenvalues=depolarising_gate_eigenvalues(c) if gate_eigenvalues is None else gate_eigenvalues,
        name=c.name + ("_gls" if full_covariance else ""),
    )
    # gate index ranges in the order of get_gate_index_dict (src/noise.jl:271-277)
    fd.gate_ptr = np.asarray([c.gate_offset[g] for g in c.total_gates] + [c.N], dtype=np.int64)
    assert np.all(np.diff(fd.gate_ptr) == [c.gate_size[g] for g in c.total_gates])
    fd.gate_types = [g[0] for g in c.total_gates]
    if full_covariance:
        fd.cov_keys, fd.cov_meas_indices, fd.cov_signs = cov_keys, cov_meas, cov_signs
        fd.cov_counts, fd.cov_rows = cov_counts, cov_rows
        fd.cov_experiment_ensemble = get_covariance_experiment_ensemble(fd)
    return fd


def _with_noise(c: Circuit, noise: str, seed: int, kw) -> Dict:
    if noise == "lognormal":
        kw = dict(kw, gate_eigenvalues=lognormal_gate_eigenvalues(c, seed=seed))
    elif noise != "depolarising":
        raise ValueError(f"unknown noise model {noise!r}")
    return kw


def rotated_design(d: int, full_covariance: bool = True, noise: str = "depolarising", seed: int = 0, **kw) -> FlatDesign:
    """`generate_design(get_circuit(get_rotated_param(d), noise_param))` with the default tuple set."""
    circuit, layer_types, n, params = rotated_planar_circuit(d)
    c = get_circuit(circuit, layer_types, n, params, name=f"rotated_planar_{d}_{d}")
    return generate_design(c, full_covariance=full_covariance, **_with_noise(c, noise, seed, kw))


def unrotated_design(d: int, full_covariance: bool = True, noise: str = "depolarising", seed: int = 0, **kw) -> FlatDesign:
    """`generate_design(get_circuit(get_unrotated_param(d), noise_param))` with the default tuple set."""
    circuit, layer_types, n, params = unrotated_planar_circuit(d)
    c = get_circuit(circuit, layer_types, n, params, name=f"unrotated_planar_{d}_{d}")
    return generate_design(c, full_covariance=full_covariance, **_with_noise(c, noise, seed, kw))


def example_design(full_covariance: bool = True, **kw) -> FlatDesign:
    circuit, layer_types, n, params = example_circuit()
    times = {"single_qubit": 29.0, "two_qubit": 29.0, "meas_reset": 660.0}
    return generate_design(get_circuit(circuit, layer_types, n, params, times, name="example_circuit"),
                           full_covariance=full_covariance, **kw)
