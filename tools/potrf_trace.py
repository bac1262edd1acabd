"""Development tool: per-task timeline of the persistent Cholesky kernel (ACES_POTRF_TRACE=1) on the Gram
matrix of a WLS fit at real rotated d (default 9): chain timing, task durations by type, worker utilisation."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["ACES_POTRF_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
fd = aces_b200.rotated_design(d, full_covariance=False)
eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
lam = np.concatenate(aces_b200.mean_ensemble(eig))
ctx = aces_b200.Context(0)
dd = aces_b200.DeviceDesign(ctx, fd)
for _ in range(3):
    dd.fit("wls", lam, None)
n = C.c_int64()
ctx.check(ctx.lib.aces_potrf_trace(ctx.handle, None, 0, C.byref(n)))
tr = np.zeros((n.value, 9), dtype=np.int64)
ctx.check(ctx.lib.aces_potrf_trace(ctx.handle, tr.ctypes.data, n.value, C.byref(n)))
t0 = tr[:, 6].min()
grab, ready, done = (tr[:, 6] - t0) / 1e3, (tr[:, 7] - t0) / 1e3, (tr[:, 8] - t0) / 1e3
typ = tr[:, 0]
print(f"tasks {len(tr)}, makespan {done.max():.1f} us, potrf phase {dd.phase_ms().get('potrf')} ms (timing off)")
names = {0: "F", 1: "S", 2: "U", 3: "S slab", 4: "U slab"}
for k, nm in names.items():
    m = typ == k
    if m.any():
        run, wait = (done - ready)[m], (ready - grab)[m]
        print(f"  {nm:7s} n={m.sum():6d} run mean {run.mean():7.2f} med {np.median(run):7.2f} max {run.max():7.2f} us | "
              f"wait mean {wait.mean():7.2f} max {wait.max():8.2f} | total run {run.sum() / 1e3:8.2f} ms, wait {wait.sum() / 1e3:8.2f} ms")
um = typ == 2
for kb in (1, 2, 3, 4):
    m = um & ((tr[:, 5] - tr[:, 4]) == kb)
    if m.any():
        print(f"  U K={128 * kb}: n={m.sum()} run mean {(done - ready)[m].mean():.2f} us -> {2 * 128 * 128 * 128 * kb / (done - ready)[m].mean() / 1e6:.2f} TFLOP/s per SM x148 = "
              f"{148 * 2 * 128 * 128 * 128 * kb / (done - ready)[m].mean() / 1e6:.1f}")
busy = (done - ready).sum()
print(f"  busy {busy / 1e3:.2f} ms of {148 * done.max() / 1e3:.2f} worker-ms ({100 * busy / (148 * done.max()):.1f} %); "
      f"waiting on dependencies {(ready - grab).sum() / 1e3:.2f} worker-ms")
f = np.nonzero(typ == 0)[0]
f = f[np.argsort(tr[f, 3])]
fd_, fr = done[f], ready[f]
gap = fr[1:] - fd_[:-1]
print(f"  chain: F run mean {(fd_ - fr).mean():.2f} us; gap F(k) done -> F(k+1) ready mean {gap.mean():.2f} med {np.median(gap):.2f} max {gap.max():.2f} us; "
      f"sum F {((fd_ - fr).sum()) / 1e3:.2f} ms + gaps {gap.sum() / 1e3:.2f} ms")
q = len(f) // 4
for a in range(0, len(f) - 1, max(q, 1)):
    b = min(a + q, len(f) - 1)
    print(f"    blocks {a:3d}..{b:3d}: gap mean {gap[a:b].mean():7.2f} us, F wait for grab->ready {(fr - grab[f])[a:b].mean():7.2f}")
np.save(os.path.join(ROOT, "gpurun_out", "potrf_trace.npy"), tr)

# where does the gap go?  tasks feeding F(k+1), relative to the end of F(k)
for k in (20, 21):
    fk = f[k]
    base = done[fk]
    print(f"  -- step {k}: F({k}) done at {base:.1f}; F({k + 1}) ready at +{ready[f[k + 1]] - base:.1f}")
    sel = np.nonzero(((tr[:, 2] == k + 1) & (tr[:, 3] == k + 1) & (typ != 0)) | ((tr[:, 2] == k + 1) & (tr[:, 3] == k) & ((typ == 1) | (typ == 3))))[0]
    sel = sel[np.argsort(done[sel])][-14:]
    for t in sel:
        print(f"     {names[int(typ[t])]:7s} tile ({tr[t, 2]},{tr[t, 3]}) k[{tr[t, 4]},{tr[t, 5]}) slab {tr[t, 1]}: grab +{grab[t] - base:8.1f} ready +{ready[t] - base:8.1f} done +{done[t] - base:8.1f}")

# the private list of helper 0 (row k+1 group, slab 0) and helper 8 (row k+2 group, slab 0) around step 20
nbk = int(tr[:, 2].max()) + 1
nn = C.c_int64()
ctx.lib.aces_potrf_schedule(nbk, 148, None, 0, C.byref(nn))
sch = np.zeros((nn.value, 6), dtype=np.int32)
ctx.lib.aces_potrf_schedule(nbk, 148, sch.ctypes.data, nn.value, C.byref(nn))
cls = sch[:, 1] >> 8
assert len(sch) == len(tr)
base = done[f[19]]
for h in (1, 9):
    idx = np.nonzero(cls == h)[0]
    idx = idx[(done[idx] > base) & (grab[idx] < done[f[21]])]
    print(f"  -- helper {h - 1}: tasks between F(19) done (t = 0) and F(21) done; F(20) ready +{ready[f[20]] - base:.1f} done +{done[f[20]] - base:.1f}")
    for t in idx:
        print(f"     {names[int(typ[t])]:7s} tile ({tr[t, 2]},{tr[t, 3]}) k[{tr[t, 4]},{tr[t, 5]}): grab +{grab[t] - base:8.1f} ready +{ready[t] - base:8.1f} done +{done[t] - base:8.1f}")
# late bulk tasks: what the helpers were waiting for
bulk_crit = np.nonzero((cls == 0) & (typ == 2) & ((tr[:, 2] - tr[:, 3]) <= 1) & (tr[:, 3] >= 19) & (tr[:, 3] <= 23))[0]
for t in bulk_crit:
    print(f"     bulk U tile ({tr[t, 2]},{tr[t, 3]}) k[{tr[t, 4]},{tr[t, 5]}): list pos {t}, grab +{grab[t] - base:8.1f} ready +{ready[t] - base:8.1f} done +{done[t] - base:8.1f}")
