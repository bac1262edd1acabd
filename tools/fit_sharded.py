"""Tuple-sharded fit over the GPUs of one node (SURVEY 8e "one large Gram"): run under torchrun.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/fit_sharded.py [rotated|unrotated] [d] [gls|wls]
Every rank builds the same real design, works on its tuple blocks, and the partial Gram matrix and
right-hand sides are summed with ncclAllReduce (shard.all_reduce_device).  Rank 0 prints one JSON
line: sharded and single-GPU time per fit, and the relative difference of the two answers."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aces_b200  # noqa: E402
from aces_b200 import shard  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "rotated"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 9
ls = sys.argv[3] if len(sys.argv) > 3 else "gls"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29511")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
gen = aces_b200.rotated_design if kind == "rotated" else aces_b200.unrotated_design
# rank 0 generates the design (minutes at d = 17) and shares it through a file; the others wait
import pickle  # noqa: E402

shipped = os.path.join(ROOT, "build", f"aces_design_{kind}_{d}.pkl")   # optional: tools/make_design_pickle.py, made beforehand
path = shipped if os.path.exists(shipped) else f"/tmp/aces_design_{kind}_{d}.pkl"
if rank == 0 and path != shipped:
    fd = gen(d, full_covariance=True, noise="lognormal", seed=9)
    with open(path + ".tmp", "wb") as f:
        pickle.dump(fd, f, protocol=4)
    os.replace(path + ".tmp", path)
dist.barrier()
if rank != 0 or path == shipped:
    with open(path, "rb") as f:
        fd = pickle.load(f)
eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
lam = np.concatenate(aces_b200.mean_ensemble(eig))
lc = np.concatenate(aces_b200.mean_ensemble(cov))
ctx = aces_b200.Context(local)
dd = aces_b200.DeviceDesign(ctx, fd)


def timed(fn):
    fn()
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, 1e3 * float(t.item()) / steps


single, ms_single = timed(lambda: dd.fit(ls, lam, lc))
sharded, ms_sharded = timed(lambda: shard.fit_tuple_sharded(dd, ls, lam, lc))
plain, ms_plain = timed(lambda: shard.fit_tuple_sharded(dd, ls, lam, lc, distributed_cholesky=True, lookahead=False))
distributed, ms_dist = timed(lambda: shard.fit_tuple_sharded(dd, ls, lam, lc, distributed_cholesky=True))
assert np.array_equal(plain, distributed)
NP = -(-dd.N // 128) * 128
gathered = [None] * world
dist.all_gather_object(gathered, sharded.tobytes() + distributed.tobytes())
if rank == 0:
    same = all(g == gathered[0] for g in gathered)
    print(json.dumps({
        "workload": f"{kind}_d{d} {ls.upper()} fit, real design, tuple blocks over {world} GPU(s)",
        "M": int(fd.M), "N": int(fd.matrix.shape[1]), "tuples": int(fd.tuple_number), "n_gpus": world,
        "tuple_shards": shard.shard_tuples(fd.mapping_lengths, world),
        "ms_per_fit_single_gpu": ms_single, "ms_per_fit_tuple_sharded": ms_sharded,
        "ms_per_fit_tuple_sharded_distributed_cholesky": ms_dist,
        "ms_per_fit_tuple_sharded_distributed_cholesky_no_lookahead": ms_plain,
        "distributed_cholesky": {"block_columns": int(-(-dd.N // 128) * 128 // dd.DIST_PANEL + (1 if (-(-dd.N // 128) * 128) % dd.DIST_PANEL else 0)),
                                 "panel_columns": dd.DIST_PANEL,
                                 "broadcast_bytes_per_fit": int(8 * sum(min(dd.DIST_PANEL, NP - c) * (NP - c) for c in range(0, NP, dd.DIST_PANEL)) + 8 * NP * 128),
                                 "allreduce_bytes_per_fit": int(8 * sum(min(dd.DIST_PANEL, NP - c) * (NP - c) for c in range(0, NP, dd.DIST_PANEL))),
                                 "collectives": "ONE ncclAllReduce of the packed lower trapezoids of the partial Gram matrices; per block column "
                                                "ncclBroadcast of its packed lower trapezoid + its block inverses, on a second stream "
                                                "(look-ahead: the next block column is factorised and sent while the current panel is applied)"},
        "rel_diff_distributed_vs_single": float(np.linalg.norm(distributed - single) / np.linalg.norm(single)),
        "gram_allreduce_bytes": int(8 * dd.fit_buffers()[0][1]),
        "rel_diff_sharded_vs_single": float(np.linalg.norm(sharded - single) / np.linalg.norm(single)),
        "all_ranks_identical": bool(same),
        "max_abs_error_vs_truth": float(np.max(np.abs(sharded - fd.gate_eigenvalues))),
    }))
dd.close()
ctx.close()
dist.barrier()
dist.destroy_process_group()
