"""Development probe: time of the Gram Cholesky (phase "potrf" of a WLS fit at real rotated d=9) with
parts of the factorisation skipped (ACES_POTRF_SKIP bit mask: 1 diagonal kernels, 2 critical tiles,
4 second-stream work, 8 look-ahead update, 16 side-stream update).  Results are wrong when skipping."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

fd = aces_b200.rotated_design(9, full_covariance=False)
eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
lam = np.concatenate(aces_b200.mean_ensemble(eig))
ctx = aces_b200.Context(0)
dd = aces_b200.DeviceDesign(ctx, fd)
ctx.set_timing(True)
for mask in [0, 1, 2, 4, 8, 16, 24, 30, 7, 6, 31 - 1, 31 - 16, 31 - 8]:
    os.environ["ACES_POTRF_SKIP"] = str(mask)
    ts = []
    for _ in range(4):
        try:
            dd.fit("wls", lam, None)
        except aces_b200.AcesError:
            pass
        ts.append(dd.phase_ms().get("potrf"))
    print(f"skip mask {mask:2d}: potrf {sorted(ts)[1]:.3f} ms", flush=True)
