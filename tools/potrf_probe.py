"""Development probe: the Gram Cholesky (phase "potrf" of a fit at real rotated d=9, N = 6784) as the
persistent task-DAG kernel against the previous multi-stream version (ACES_POTRF_STREAMS=1, read once per
process: the two are timed in separate processes), with the fitted gate eigenvalues compared."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag):
    import aces_b200

    d = int(os.environ.get("PROBE_D", "9"))
    fd = aces_b200.rotated_design(d, full_covariance=True, noise="lognormal", seed=9)
    eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    ctx = aces_b200.Context(0)
    dd = aces_b200.DeviceDesign(ctx, fd)
    ctx.set_timing(True)
    out = {}
    for name in ("wls", "gls"):
        ts, phases = [], None
        for _ in range(6):
            g = dd.fit(name, lam, lc)
            phases = dd.phase_ms()
            ts.append(phases.get("potrf"))
        print(f"[{tag}] {name}: potrf {sorted(ts)[2]:.3f} ms (min {min(ts):.3f}); phases {({k: round(v, 3) for k, v in phases.items()})}", flush=True)
        out[name] = g
    np.savez(os.path.join(ROOT, "gpurun_out", f"potrf_probe_{tag}.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for tag, env in (("dag", {}), ("streams", {"ACES_POTRF_STREAMS": "1"})):
        subprocess.run([sys.executable, __file__, tag], env=dict(os.environ, **env), timeout=300, check=False)
    a = np.load(os.path.join(ROOT, "gpurun_out", "potrf_probe_dag.npz"))
    b = np.load(os.path.join(ROOT, "gpurun_out", "potrf_probe_streams.npz"))
    for k in ("wls", "gls"):
        print(f"{k}: |dag - streams| / |streams| = {np.linalg.norm(a[k] - b[k]) / np.linalg.norm(b[k]):.3e}")
