"""K1 development probe: which property of a real design costs time -- the tables (Pauli weights,
experiment sizes) or the shot split (unequal rows per experiment)?  Runs the d=9 WLS workload with
the four combinations of {real, synthetic} tables x {real, uniform} shot weights."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

ctx = aces_b200.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
real = aces_b200.rotated_design(9, full_covariance=False)
syn = aces_b200.synthetic_design("rotated", 9, full_covariance=False)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for tables, weights in (("real", "real"), ("real", "uniform"), ("synthetic", "real"), ("synthetic", "uniform")):
    fd = real if tables == "real" else syn
    saved = fd.shot_weights.copy()
    X = fd.experiment_numbers.astype(np.float64)
    fd.shot_weights = real.shot_weights.copy() if weights == "real" else X / X.sum()
    est = aces_b200.DesignEstimator(ctx, fd)
    shots = 10**8
    shots_set = [shots // 100, shots // 10, shots][3 - K:]
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    rows = [int(shots_max[int(est.tuple_of_experiment[e])]) for e in range(n_exp)]
    ld = [(r + 127) // 128 * 128 for r in rows]
    prefix = np.stack([shots_divided[int(est.tuple_of_experiment[e])] for e in range(n_exp)])
    g = torch.Generator(device="cuda").manual_seed(1)
    dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
    ptrs = [t.data_ptr() for t in dev]
    sizes = est.experiment_sizes()
    out = torch.zeros(int(np.sum(sizes) * K), dtype=torch.int64, device="cuda")
    nbytes = sum(rows[e] * B for e in range(n_exp))
    batch = est.prepare(ptrs, rows, ld, prefix, out.data_ptr(), asynchronous=True)
    for _ in range(5):
        batch.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 30
    e0.record(stream)
    for _ in range(steps):
        batch.launch()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"tables={tables:9s} weights={weights:7s} K={K} rows {min(rows)}..{max(rows)} E {int(sizes.min())}..{int(sizes.max())} "
          f"mean E {float(np.mean(sizes)):.0f}  {nbytes / 1e9:.2f} GB  {ms:.3f} ms/step  {nbytes / ms / 1e6:.0f} GB/s "
          f"({100 * nbytes / ms / 1e6 / 6547.2:.1f}%)", flush=True)
    fd.shot_weights = saved
    del dev, out, est, batch
    torch.cuda.empty_cache()
