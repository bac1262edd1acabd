"""Times the fit entry points at a surface-code shape (development probe; bench.py is the
reported measurement).  python tools/fit_probe.py [d] [kind]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aces_b200  # noqa: E402
from fit_util import fo, noisy_ensembles  # noqa: E402


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    kind = sys.argv[2] if len(sys.argv) > 2 else "rotated"
    fd = aces_b200.synthetic_design(kind, d, full_covariance=True, with_matrix=True)
    eig, cov = noisy_ensembles(fd, seed=3)
    lam = np.concatenate(fo.mean_ensemble(eig))
    t0 = time.perf_counter()
    cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
    F = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
    print(f"cpu oracle covariance + factor: {time.perf_counter() - t0:.2f} s, F nnz {F.nnz}")
    dfac = sp.diags(cl.diagonal() ** -0.5)
    with aces_b200.Context(0) as ctx:
        for name, fac in (("ols", None), ("wls", dfac), ("gls", F)):
            for rep in range(3):
                ctx.set_timing(True)
                t0 = time.perf_counter()
                g = aces_b200.estimate_gate_eigenvalues(ctx, fd.matrix, lam, fac)
                dt = time.perf_counter() - t0
                print(f"{name} fit #{rep}: {dt * 1e3:.1f} ms wall, gram kernel {ctx.last_kernel_ms():.2f} ms, "
                      f"err vs truth {np.max(np.abs(g - fd.gate_eigenvalues)):.3e}")
        t0 = time.perf_counter()
        S = aces_b200.calc_gls_covariance_from_factor(ctx, fd.matrix, fd.gate_eigenvalues, F)
        print(f"gls covariance: {(time.perf_counter() - t0) * 1e3:.1f} ms, trace {np.trace(S):.6e}")
    if d <= 5:
        t0 = time.perf_counter()
        ref = fo.estimate_gate_eigenvalues(fd.matrix, lam, F)
        print(f"oracle gls {time.perf_counter() - t0:.2f} s, rel diff {np.linalg.norm(g - ref) / np.linalg.norm(ref):.3e}")


if __name__ == "__main__":
    main()
