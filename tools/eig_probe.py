"""Times aces_symmetric_eigenvalues against LAPACK on the host.  The matrix is device-resident for the timed
calls (the entry point takes host or device pointers), so the figure is the solver, not the PCIe copy."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces_b200  # noqa: E402
from aces_b200._lib import _ptr  # noqa: E402

sizes = [int(s) for s in sys.argv[1:] if not s.startswith("-")] or [1024, 2048, 4096, 6779]
ctx = aces_b200.Context(0)
for n in sizes:
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = np.asfortranarray((a + a.T) / 2)
    host = aces_b200.eigvals(ctx, a)
    dev = torch.from_numpy(np.ascontiguousarray(a)).cuda()      # symmetric: row- and column-major coincide
    w = np.zeros(n)
    reps = 3
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.check(ctx.lib.aces_symmetric_eigenvalues(ctx.handle, n, C.c_void_p(dev.data_ptr()), n, _ptr(w)))
    gpu = (time.perf_counter() - t0) / reps
    assert np.array_equal(w, host)
    want = None
    cpu = float("nan")
    if "--no-cpu" not in sys.argv:
        t0 = time.perf_counter()
        want = np.linalg.eigvalsh(a)
        cpu = time.perf_counter() - t0
    err = np.max(np.abs(w - want)) / np.abs(want).max() if want is not None else float("nan")
    traffic = 4.0 * n ** 3 / 3
    print(f"n={n}: gpu {gpu * 1e3:.1f} ms device-resident ({traffic / gpu / 1e9:.0f} GB/s on the 4n^3/3 algorithmic bytes), "
          f"LAPACK eigvalsh {cpu * 1e3:.0f} ms, max err {err:.2e} of |A|", flush=True)
ctx.close()
