"""Times aces_symmetric_eigenvalues (host matrix in, eigenvalues out) against LAPACK on the host."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces_b200  # noqa: E402

sizes = [int(s) for s in sys.argv[1:]] or [1024, 2048, 4096, 6779]
ctx = aces_b200.Context(0)
for n in sizes:
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = np.asfortranarray((a + a.T) / 2)
    aces_b200.eigvals(ctx, a)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        got = aces_b200.eigvals(ctx, a)
    gpu = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    want = np.linalg.eigvalsh(a)
    cpu = time.perf_counter() - t0
    err = np.max(np.abs(got - want)) / np.abs(want).max()
    traffic = 8.0 * n ** 3 / 3
    print(f"n={n}: gpu {gpu * 1e3:.1f} ms (trailing products {traffic / gpu / 1e9:.0f} GB/s equivalent), "
          f"LAPACK eigvalsh {cpu * 1e3:.0f} ms, max err {err:.2e} of |A|", flush=True)
ctx.close()
