"""One GLS fit + one GLS covariance at a surface-code shape (for ncu launch lists)."""
import os, sys
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import aces_b200
from fit_util import fo, noisy_ensembles
d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
fd = aces_b200.synthetic_design("rotated", d, full_covariance=True, with_matrix=True)
eig, cov = noisy_ensembles(fd, seed=11)
lam = np.concatenate(fo.mean_ensemble(eig))
cl = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
F = fo.sparse_covariance_inv_factor(cl, fd.mapping_lengths)
with aces_b200.Context(0) as ctx:
    dd = aces_b200.DeviceDesign(ctx, fd)
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    g = dd.fit(sys.argv[2] if len(sys.argv) > 2 else "gls", lam, lc)
    print("fit ok", float(np.max(np.abs(g - fd.gate_eigenvalues))))
