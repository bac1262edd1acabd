"""Times the design-resident fit entry points at a surface-code shape (development probe;
bench.py is the reported measurement).  python tools/fit_probe.py [d] [kind] [--cov]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aces_b200  # noqa: E402
from fit_util import fo, noisy_ensembles  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    d = int(args[0]) if len(args) > 0 else 9
    kind = args[1] if len(args) > 1 else "rotated"
    fd = aces_b200.synthetic_design(kind, d, full_covariance=True, with_matrix=True)
    eig, cov = noisy_ensembles(fd, seed=11)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    with aces_b200.Context(0) as ctx:
        dd = aces_b200.DeviceDesign(ctx, fd)
        for name in ("ols", "wls", "gls"):
            for rep in range(3):
                t0 = time.perf_counter()
                g = dd.fit(name, lam, lc)
                dt = time.perf_counter() - t0
                print(f"{name} fit #{rep}: {dt * 1e3:.2f} ms wall, err vs truth {np.max(np.abs(g - fd.gate_eigenvalues)):.3e}")
        if "--cov" in sys.argv:
            glam = fo.get_eigenvalues(fd, fd.gate_eigenvalues)
            cl = fo.calc_covariance_log(fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, fd.gate_eigenvalues),
                                                           fo.calc_covariance_eigenvalues(fd, fd.gate_eigenvalues)), glam)
            for name in ("gls", "wls", "ols"):
                for rep in range(2):
                    t0 = time.perf_counter()
                    _, tr, trsq = dd.calc_ls_covariance(name, cl, want_matrix=False)
                    print(f"{name} covariance moments #{rep}: {(time.perf_counter() - t0) * 1e3:.1f} ms, trace {tr:.6e}")
        dd.close()


if __name__ == "__main__":
    main()
