"""One fit at a surface-code shape on the real design (for ncu launch lists).
python tools/fit_once.py [d] [gls|wls|ols]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
fd = aces_b200.rotated_design(d, full_covariance=True, noise="lognormal", seed=9)
eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
lam = np.concatenate(aces_b200.mean_ensemble(eig))
lc = np.concatenate(aces_b200.mean_ensemble(cov))
with aces_b200.Context(0) as ctx:
    dd = aces_b200.DeviceDesign(ctx, fd)
    g = dd.fit(sys.argv[2] if len(sys.argv) > 2 else "gls", lam, lc)
    print("fit ok", float(np.max(np.abs(g - fd.gate_eigenvalues))))
