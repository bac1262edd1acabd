"""Fits per second with several fit contexts running concurrently on one GPU (independent fits:
budgets x repetitions, src/simulate.jl:393).  python tools/fit_concurrent.py [d] [threads ...]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
counts = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]
fd = aces_b200.synthetic_design("rotated", d, full_covariance=True, with_matrix=True)
eig, cov = aces_b200.synthetic_estimates(fd, seed=11)
lam = np.concatenate(aces_b200.mean_ensemble(eig))
lc = np.concatenate(aces_b200.mean_ensemble(cov))
ref = None
for c in counts:
    ctxs = [aces_b200.Context(0) for _ in range(c)]
    dds = [aces_b200.DeviceDesign(x, fd) for x in ctxs]
    for name in ("gls", "wls"):
        outs = [None] * c

        def work(i, n):
            for _ in range(n):
                outs[i] = dds[i].fit(name, lam, lc)

        ths = [threading.Thread(target=work, args=(i, 2)) for i in range(c)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        n = 10
        ths = [threading.Thread(target=work, args=(i, n)) for i in range(c)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        dt = time.perf_counter() - t0
        if ref is None:
            ref = {}
        ref.setdefault(name, outs[0])
        ok = all(np.array_equal(o, ref[name]) for o in outs)
        print(f"{c} concurrent contexts, {name}: {c * n / dt:7.1f} fits/s ({1e3 * dt / n:6.2f} ms per round), identical results: {ok}", flush=True)
    for x in dds:
        x.close()
    for x in ctxs:
        x.close()
