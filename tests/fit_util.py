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
    import aces_b200

    return aces_b200.synthetic_estimates(fd, shots=shots, seed=seed, noise=noise)
