"""Shared helpers of the fit tests: synthetic noisy estimates for a FlatDesign."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fit_oracle as fo  # noqa: E402


def noisy_ensembles(fd, shots=10**8, seed=0, noise=True):
    """Per-experiment estimate ensembles ([t][pauli] -> list, one value per experiment that
    measures it) drawn around the true circuit eigenvalues exp(A log g)."""
    rng = np.random.default_rng(seed)
    lam_ens = fo.get_eigenvalues_ensemble(fd, fd.gate_eigenvalues)
    cov_lam = fo.calc_covariance_eigenvalues(fd, fd.gate_eigenvalues) if fd.full_covariance else None
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
            for c, (p, q) in enumerate(fd.cov_keys[t]):
                k = max(int(fd.cov_counts[t][c]), 1)
                s = 1e-9 if noise else 0.0
                ratio = cov_lam[t][c] / (lam_ens[t][p] * lam_ens[t][q])
                base = np.mean(eig_t[p]) * np.mean(eig_t[q]) * ratio
                cov_t.append(list(np.minimum(base + s * rng.standard_normal(k), 1.0)))
        cov.append(cov_t)
    return eig, cov
