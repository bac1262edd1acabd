"""Generates tests/golden/parity_golden.npz.

The reference is Julia and cannot be executed in this image, so the fixture is produced by a
literal pure-Python transcription of pauli_estimate / estimate_experiment
(reference src/estimate.jl:239-254, 284-301) -- independent of the C oracle in oracle/ and of
the CUDA kernel, both of which are tested against it.  Run:  python tests/golden/make_parity_golden.py
"""
import os

import numpy as np


def fld(a, b):
    return a // b


def pauli_estimate(stim_counts, pauli_meas_indices, shots_value):
    # stim_counts[s, c] with 1-based s and c as in the Julia source
    pauli_shots = 0
    for s in range(1, shots_value + 1):
        pauli_shot = 0
        for m in pauli_meas_indices:
            pauli_shot += (int(stim_counts[s - 1, (1 + fld(m - 1, 8)) - 1]) >> ((m - 1) % 8)) & 1
        pauli_shots += 1 - 2 * (pauli_shot % 2)
    return pauli_shots


def estimate_experiment(counts, experiment_meas_indices, experiment_pauli_signs, shots_value):
    E = len(experiment_meas_indices)
    assert E == len(experiment_pauli_signs)
    est = np.zeros(E, dtype=np.float64)
    ps = np.zeros(E, dtype=np.int64)
    for idx in range(E):
        pauli_shots = pauli_estimate(counts, experiment_meas_indices[idx], shots_value)
        est[idx] = ((-1) ** int(experiment_pauli_signs[idx])) * pauli_shots / shots_value
        ps[idx] = pauli_shots
    return est, ps


def main():
    rng = np.random.default_rng(20261019)
    cases = {}
    specs = [  # (n_qubits, rows, prefixes)
        (17, 203, [1, 77, 203]),
        (49, 1037, [16, 500, 1037]),
        (161, 333, [128, 129, 333]),
        (9, 2500, [1024, 2048, 2500]),
    ]
    for ci, (n, rows, prefixes) in enumerate(specs):
        B = (n + 7) // 8
        counts = rng.integers(0, 256, size=(rows, B), dtype=np.uint8)
        E = 12
        supports = []
        for p in range(E):
            w = [1, 1, 2, 2, 3, 4, 5, 7, 9, 2, 1, 0][p]
            supports.append(sorted(rng.choice(n, size=min(w, n), replace=False) + 1))
        supports[9] = [supports[9][0], supports[9][0]]  # a repeated qubit cancels mod 2
        signs = rng.integers(0, 2, size=E).astype(np.uint8)
        ptr = np.zeros(E + 1, dtype=np.int64)
        for i, s in enumerate(supports):
            ptr[i + 1] = ptr[i] + len(s)
        flat = np.asarray([q for s in supports for q in s], dtype=np.int32)
        est = np.zeros((len(prefixes), E))
        ps = np.zeros((len(prefixes), E), dtype=np.int64)
        for k, S in enumerate(prefixes):
            est[k], ps[k] = estimate_experiment(counts, supports, signs, S)
        cases[f"c{ci}_n"] = np.int64(n)
        cases[f"c{ci}_counts"] = counts
        cases[f"c{ci}_ptr"] = ptr
        cases[f"c{ci}_idx"] = flat
        cases[f"c{ci}_signs"] = signs
        cases[f"c{ci}_prefixes"] = np.asarray(prefixes, dtype=np.int64)
        cases[f"c{ci}_est"] = est
        cases[f"c{ci}_pauli_shots"] = ps
    cases["n_cases"] = np.int64(len(specs))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parity_golden.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
