"""K1 throughput at the BASELINE.json shapes (development probe; bench.py is the reported number).
python tools/k1_shapes.py [kind:d:total_shots:full_cov ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402

cases = sys.argv[1:] or ["rotated:3:6e8:0", "rotated:5:3e8:1", "rotated:9:1e8:0", "rotated:9:1e8:1",
                         "unrotated:7:1e8:1", "rotated:17:1.25e8:0",
                         "real:3:6e8:0", "real:5:3e8:1", "real:9:1e8:0", "real:9:1e8:1", "realun:7:1e8:1", "real:17:1.25e8:0"]
ctx = aces_b200.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
assert stream.cuda_stream != 0
ctx.set_stream(stream.cuda_stream)
for case in cases:
    kind, d, shots, full = case.split(":")
    d, shots, full = int(d), int(float(shots)), bool(int(full))
    if kind.startswith("real"):   # the reference's own tables (circuit_design.py)
        fd = (aces_b200.unrotated_design if kind == "realun" else aces_b200.rotated_design)(d, full_covariance=full)
    else:
        fd = aces_b200.synthetic_design(kind, d, full_covariance=full)
    est = aces_b200.DesignEstimator(ctx, fd)
    shots_set = [shots // 100, shots // 10, shots]
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    rows = [int(shots_max[int(est.tuple_of_experiment[e])]) for e in range(n_exp)]
    ld = [(r + 127) // 128 * 128 for r in rows]
    prefix = np.stack([shots_divided[int(est.tuple_of_experiment[e])] for e in range(n_exp)])
    g = torch.Generator(device="cuda").manual_seed(1)
    dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
    ptrs = [t.data_ptr() for t in dev]
    sizes = est.experiment_sizes()
    out = torch.zeros(int(np.sum(sizes) * 3), dtype=torch.int64, device="cuda")
    nbytes = sum(rows[e] * B for e in range(n_exp))
    rbytes = sum(rows[e] * est.tables.columns_read(e) for e in range(n_exp))  # what the kernel reads
    for _ in range(3):
        est.odd_counts(ptrs, rows, ld, prefix, out=out.data_ptr(), asynchronous=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record(stream)
    for _ in range(steps):
        est.odd_counts(ptrs, rows, ld, prefix, out=out.data_ptr(), asynchronous=True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{case:24s} n={fd.n_qubits:4d} B={B:3d} E={int(sizes.max()):4d} matrix {nbytes / 1e9:6.2f} GB, read {rbytes / 1e9:6.2f} GB  "
          f"{ms:8.3f} ms/step  {rbytes / ms / 1e6:8.1f} GB/s read ({100 * rbytes / ms / 1e6 / 6547.2:5.1f}% of measured HBM peak), "
          f"{nbytes / ms / 1e6:8.1f} GB/s of matrix", flush=True)
    del dev, out, est
    torch.cuda.empty_cache()
