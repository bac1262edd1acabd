"""K1 at d=9: GPU time per step and host enqueue time per step versus the number of steps."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200
ctx = aces_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); assert stream.cuda_stream != 0; ctx.set_stream(stream.cuda_stream)
fd = aces_b200.rotated_design(9, full_covariance=False)
est = aces_b200.DesignEstimator(ctx, fd)
shots_divided, shots_max = aces_b200.shots_divide(fd, [10**6, 10**7, 10**8])
n_exp, B = fd.experiment_number, fd.n_bytes
rows = [int(shots_max[int(est.tuple_of_experiment[e])]) for e in range(n_exp)]
ld = [(r + 127) // 128 * 128 for r in rows]
prefix = np.stack([shots_divided[int(est.tuple_of_experiment[e])] for e in range(n_exp)])
g = torch.Generator(device="cuda").manual_seed(1)
dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
out = torch.zeros(int(np.sum(est.experiment_sizes()) * 3), dtype=torch.int64, device="cuda")
batch = est.prepare([t.data_ptr() for t in dev], rows, ld, prefix, out.data_ptr(), asynchronous=True)
for _ in range(5): batch.launch()
torch.cuda.synchronize()
for steps in (5, 10, 20, 50, 100, 200, 20, 10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(steps): batch.launch()
    e1.record(stream); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"steps {steps:4d}: gpu {e0.elapsed_time(e1) / steps:.4f} ms/step, host enqueue {1e3 * th / steps:.4f} ms/step", flush=True)
