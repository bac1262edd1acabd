"""calc_ensemble_scaling's repetition loop as a replica farm over the ranks of a torchrun launch (NCCL):
python -m torch.distributed.run --nproc-per-node N tools/merit_farm_nccl.py [d] [instances]"""
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
from aces_b200 import merit, shard  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 5
R = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
fd = aces_b200.rotated_design(d, full_covariance=True, noise="lognormal", seed=9)
rng = np.random.default_rng(1)
g0 = np.asarray(fd.gate_eigenvalues, dtype=np.float64)
sets = [1.0 - (1.0 - g0) * rng.uniform(0.5, 1.5, size=len(g0)) for _ in range(R)]   # same on every rank
ctx = aces_b200.Context(local)
dd = aces_b200.DeviceDesign(ctx, fd)
tf = (aces_b200.get_transform(fd, "marginal"), aces_b200.get_transform(fd, "relative"))
fn = lambda g: merit.calc_merit(dd, gate_eigenvalues=g, transforms=tf)  # noqa: E731
fn(sets[0])
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
table = shard.merit_farm(fn, sets)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    k = shard.MERIT_FARM_FIELDS.index("gls_expectation")
    gls = table[:, k]
    print(json.dumps({"what": "merit_farm (calc_merit replicas, src/scaling.jl:510-536)", "design": f"rotated d={d}", "N": int(dd.N),
                      "instances": R, "n_gpus": world, "seconds": dt, "merits_per_s": R / dt,
                      "gls_expectation_mean": float(gls.mean()), "gls_rel_sem": float(gls.std(ddof=1) / (gls.mean() * np.sqrt(R))),
                      "table_checksum": float(np.sum(table))}))
dd.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
