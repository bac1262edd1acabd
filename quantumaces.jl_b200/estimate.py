"""Host-side mirror of the reference's estimation entry points for the K1 path.

Same names, argument meaning and error behaviour as the Julia functions they stand for
(src/estimate.jl:284-301, src/simulate.jl:195-234); every count comes from the CUDA kernel
through the C ABI (aces_parity_counts / aces_estimate_experiment).  Nothing here computes a
parity on the CPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .design import FlatDesign, build_pauli_tables, shots_divide


def _csr(lists: Sequence[Sequence[int]]):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    flat = np.zeros(int(ptr[-1]), dtype=np.int32)
    for i, l in enumerate(lists):
        flat[ptr[i]:ptr[i + 1]] = l
    return ptr, flat


def estimate_experiment(
    ctx: _lib.Context,
    counts: np.ndarray,
    experiment_meas_indices: Sequence[Sequence[int]],
    experiment_pauli_signs: Sequence[bool],
    shots_value: int,
) -> np.ndarray:
    """estimate_experiment(counts::Matrix{UInt8}, experiment_meas_indices, experiment_pauli_signs,
    shots_value) of src/estimate.jl:284-301.  `counts` is the shots x bytes uint8 matrix (any
    memory order; Fortran order is passed without a copy), indices are 1-based qubits."""
    E = len(experiment_meas_indices)
    # src/estimate.jl:292
    assert E == len(experiment_pauli_signs), "The number of signs does not match the number of circuit eigenvalues."
    ptr, flat = _csr(experiment_meas_indices)
    return _lib.estimate_experiment_raw(ctx, counts, ptr, flat, np.asarray(experiment_pauli_signs, dtype=np.uint8), shots_value)


class DesignEstimator:
    """Per-design state of the K1 path: the device-resident Pauli tables (built once, the work
    get_experiment_data / get_covariance_experiment_data repeat every repetition in the
    reference, src/simulate.jl:124-135) and the scatter maps of src/simulate.jl:214-233."""

    def __init__(self, ctx: _lib.Context, fd: FlatDesign):
        self.ctx = ctx
        self.fd = fd
        exp_ptr, pauli_ptr, pauli_bits, n_eig = build_pauli_tables(fd)
        self.exp_ptr, self.n_eig = exp_ptr, n_eig
        self.tables = _lib.PauliTables(ctx, fd.n_qubits, exp_ptr, pauli_ptr, pauli_bits)
        # signs in table order
        sg = []
        for t in range(fd.tuple_number):
            for j in range(int(fd.experiment_numbers[t])):
                exp = fd.experiment_ensemble[t][j]
                sg.append(np.asarray(fd.pauli_signs[t])[exp].astype(np.int64))
                if fd.cov_keys and len(fd.cov_keys[t]) > 0:
                    ce = fd.cov_experiment_ensemble[t][j]
                    sg.append(np.asarray(fd.cov_signs[t])[ce].astype(np.int64))
        self.sign_factor = 1 - 2 * np.concatenate(sg) if sg else np.zeros(0, dtype=np.int64)
        self.tuple_of_experiment = np.repeat(np.arange(fd.tuple_number), fd.experiment_numbers)

    def experiment_sizes(self) -> np.ndarray:
        return np.diff(self.exp_ptr)

    def odd_counts(self, matrices, rows, ld, prefix_shots, experiment_ids=None, out=None,
                   accumulate=False, asynchronous=False, rowmajor=False):
        """One launch over a batch of sample matrices (aces_parity_counts; `rowmajor`: Stim's C-ordered
        [shots x bytes] records through aces_parity_counts_rowmajor, `ld` = bytes between shots)."""
        if experiment_ids is None:
            experiment_ids = np.arange(self.fd.experiment_number, dtype=np.int32)
        return _lib.parity_counts(self.ctx, self.tables, experiment_ids, matrices, rows, ld,
                                  prefix_shots, out=out, accumulate=accumulate,
                                  asynchronous=asynchronous, rowmajor=rowmajor)

    def prepare(self, matrices, rows, ld, prefix_shots, out, experiment_ids=None, accumulate=False,
                asynchronous=False, rowmajor=False) -> "_lib.PreparedBatch":
        """Marshals the arguments of odd_counts once; `.launch()` repeats the call."""
        if experiment_ids is None:
            experiment_ids = np.arange(self.fd.experiment_number, dtype=np.int32)
        return _lib.PreparedBatch(self.ctx, self.tables, experiment_ids, matrices, rows, ld, prefix_shots, out,
                                  accumulate=accumulate, asynchronous=asynchronous, rowmajor=rowmajor)

    def simulate_estimate(self, samples: Sequence[np.ndarray], shots_set: Sequence[int]):
        """The estimation half of simulate_stim_estimate (src/simulate.jl:195-234): given one
        sample matrix per experiment (tuple-major order, shots_maximum[t] rows each), returns
        (est_eigenvalues_experiment_ensemble_set, est_covariance_experiment_ensemble_set) with
        the reference's nesting [s][t][pauli_idx] -> list of per-experiment estimates."""
        fd = self.fd
        shots_divided, shots_maximum = shots_divide(fd, shots_set)
        K = len(shots_set)
        n_exp = fd.experiment_number
        assert len(samples) == n_exp
        rows, ld, mats, prefix = [], [], [], np.zeros((n_exp, K), dtype=np.int64)
        # Stim hands back C-ordered records: taken as they are (row-major ingest) when every sample is
        # C-contiguous; Fortran-ordered matrices (the reference's Matrix{UInt8}) go in place as well
        rowmajor = all(sm.flags.c_contiguous and not sm.flags.f_contiguous for sm in samples)
        for e, sm in enumerate(samples):
            t = int(self.tuple_of_experiment[e])
            if not rowmajor and not sm.flags.f_contiguous:
                sm = np.asfortranarray(sm)
            assert sm.shape[0] >= shots_maximum[t], "sample matrix has fewer rows than shots_maximum"
            mats.append(sm)
            rows.append(sm.shape[0])
            ld.append(sm.shape[1] if rowmajor else max(sm.shape[0], 1))
            prefix[e, :] = shots_divided[t, :]
        # budgets need not be sorted in the reference; the kernel wants nested prefixes
        order = np.argsort(np.asarray(shots_set), kind="stable")
        odd = self.odd_counts(mats, rows, ld, prefix[:, order], rowmajor=rowmajor)
        return self.scatter(odd, prefix[:, order], order)

    def estimate_ensemble(self, jobs, job_indices, pauli_sign_ensemble, covariance_pauli_sign_ensemble,
                          experiment_shots: int):
        """The estimation loop of estimate_stim_ensemble (src/device.jl:178-229; the Matrix{UInt8} half of
        estimate_qiskit_ensemble, :272-323, is the same): a randomised design runs thousands of small
        circuits (64 ... 10^4 shots each) in jobs; circuit c of job b ran experiment j of tuple i under
        randomisation r -- job_indices[b][c] = (i, j, r), 0-based -- with the Pauli signs
        pauli_sign_ensemble[i][j][r] (and covariance_pauli_sign_ensemble[i][j][r]).  Where the reference
        calls estimate_experiment twice per circuit, here ALL circuits of a job go through one
        aces_parity_counts launch (n_items = circuits of the job, K = 1, both Pauli sets of the experiment in
        the same pass; small host matrices are packed into one pinned upload).  Returns
        (est_eigenvalues_experiment_ensemble, est_covariance_experiment_ensemble): [i][pauli_idx] -> list."""
        fd = self.fd
        has_cov = bool(fd.cov_keys)
        eig_ens = [[[] for _ in range(int(m))] for m in fd.mapping_lengths]
        cov_ens = [[[] for _ in range(len(fd.cov_keys[t]) if has_cov else 0)] for t in range(fd.tuple_number)]
        exp0 = np.concatenate([[0], np.cumsum(fd.experiment_numbers)]).astype(np.int64)
        S = int(experiment_shots)
        for counts_list, indices in zip(jobs, job_indices):
            # src/device.jl:192
            assert len(counts_list) == len(indices), "The number of circuits in the job does not match the number of counts."
            if not len(indices):
                continue
            eids = np.asarray([exp0[i] + j for (i, j, _) in indices], dtype=np.int32)
            mats = [m if m.flags.f_contiguous else np.asfortranarray(m) for m in counts_list]
            rows = [m.shape[0] for m in mats]
            ld = [max(m.shape[0], 1) for m in mats]
            prefix = np.full((len(mats), 1), S, dtype=np.int64)
            odd = self.odd_counts(mats, rows, ld, prefix, experiment_ids=eids)
            off = 0
            for (i, j, r), e in zip(indices, eids):
                E = int(self.exp_ptr[e + 1] - self.exp_ptr[e])
                ne = int(self.n_eig[e])
                pauli_shots = S - 2 * odd[off:off + E]
                off += E
                sg = np.asarray(pauli_sign_ensemble[i][j][r], dtype=np.int64)
                assert len(sg) == ne, "The number of signs does not match the number of circuit eigenvalues."
                est = ((1 - 2 * sg) * pauli_shots[:ne]) / np.float64(S)      # src/estimate.jl:297-298
                for idx, pidx in enumerate(fd.experiment_ensemble[i][j]):
                    eig_ens[i][int(pidx)].append(float(est[idx]))
                if has_cov and len(fd.cov_keys[i]) > 0:
                    csg = np.asarray(covariance_pauli_sign_ensemble[i][j][r], dtype=np.int64)
                    assert len(csg) == E - ne
                    cest = ((1 - 2 * csg) * pauli_shots[ne:]) / np.float64(S)
                    for idx, cidx in enumerate(fd.cov_experiment_ensemble[i][j]):
                        cov_ens[i][int(cidx)].append(float(cest[idx]))
        return eig_ens, cov_ens

    def scatter(self, odd: np.ndarray, prefix_sorted: np.ndarray, order: np.ndarray):
        """Forms ((-1)^sign * pauli_shots) / shots_value (src/estimate.jl:297-298) from the odd
        counts and pushes each estimate into [s][t][pauli_idx] (src/simulate.jl:214-233)."""
        fd = self.fd
        K = prefix_sorted.shape[1]
        eig_set = [[[[] for _ in range(int(m))] for m in fd.mapping_lengths] for _ in range(K)]
        has_cov = bool(fd.cov_keys)
        cov_set = [
            [[[] for _ in range(len(fd.cov_keys[t]) if has_cov else 0)] for t in range(fd.tuple_number)]
            for _ in range(K)
        ]
        off = 0
        for e in range(fd.experiment_number):
            t = int(self.tuple_of_experiment[e])
            j = e - int(np.sum(fd.experiment_numbers[:t]))
            E = int(self.exp_ptr[e + 1] - self.exp_ptr[e])
            ne = int(self.n_eig[e])
            sf = self.sign_factor[self.exp_ptr[e]:self.exp_ptr[e + 1]]
            exp = fd.experiment_ensemble[t][j]
            cexp = fd.cov_experiment_ensemble[t][j] if has_cov and len(fd.cov_keys[t]) > 0 else []
            for kk in range(K):
                s = int(order[kk])
                S = int(prefix_sorted[e, kk])
                o = odd[off + kk * E: off + (kk + 1) * E]
                pauli_shots = S - 2 * o
                est = (sf * pauli_shots) / np.float64(S)
                for idx, pidx in enumerate(exp):
                    eig_set[s][t][int(pidx)].append(float(est[idx]))
                for idx, cidx in enumerate(cexp):
                    cov_set[s][t][int(cidx)].append(float(est[ne + idx]))
            off += E * K
        return eig_set, cov_set
