"""Sharding of the ACES estimation path over GPUs (one process per GPU, torch.distributed).

The path partitions along three axes, none of which needs a data-path
collective -- only the small count / estimate vectors are exchanged:

* repetitions (src/simulate.jl:371) are independent: round-robin over ranks, results gathered;
* experiments of one repetition (src/simulate.jl:145-155) are independent sample matrices:
  balanced over ranks by bytes, the odd-parity counts combined with all_reduce(sum, int64);
* shots of one experiment are additive: contiguous row ranges per rank with the budget prefixes
  (src/simulate.jl:206-213) clamped into each range, counts combined with all_reduce(sum, int64).

Integer sums are exact, so the combined counts equal the single-GPU counts bit for bit whatever
the sharding.  Backend: NCCL over NVLink for device tensors, gloo for the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_repetitions(repetitions: int, world: int, rank: int) -> List[int]:
    """Repetition indices (0-based) handled by `rank`: round-robin."""
    assert 0 <= rank < world
    return list(range(rank, repetitions, world))


def shard_experiments(experiment_bytes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time split of experiments by sample-matrix bytes
    (shots_maximum[t] * ceil(n / 8)).  Returns the experiment ids of every rank, ascending."""
    order = np.argsort(-np.asarray(experiment_bytes, dtype=np.int64), kind="stable")
    load = np.zeros(world, dtype=np.int64)
    out: List[List[int]] = [[] for _ in range(world)]
    for e in order:
        r = int(np.argmin(load))
        out[r].append(int(e))
        load[r] += int(experiment_bytes[int(e)])
    return [sorted(x) for x in out]


def shard_shots(rows: int, prefix_shots: Sequence[int], world: int, rank: int,
                align: int = 128) -> Tuple[int, int, np.ndarray]:
    """Contiguous row range [lo, hi) of `rank` for a sample matrix of `rows` shots and the budget
    prefixes restricted to it: local_prefix[k] = clamp(prefix[k] - lo, 0, hi - lo).  Range starts
    are multiples of `align` rows so that every shard can be read in place through the TMA path
    (16-byte aligned columns).  The per-rank counts add up to the counts of the whole matrix."""
    assert 0 <= rank < world and rows >= 0
    per = -(-rows // world)
    per = -(-per // align) * align
    lo = min(rank * per, rows)
    hi = min(lo + per, rows)
    pre = np.asarray(prefix_shots, dtype=np.int64)
    return lo, hi, np.clip(pre - lo, 0, hi - lo)


def all_reduce_counts(counts, group=None):
    """In-place all_reduce(sum) of an int64 count tensor (device tensor: NCCL; CPU tensor: gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def all_gather_vectors(vec, group=None):
    """Gathers one equally sized vector per rank (per-repetition counts or estimates) into a
    [world, len] tensor on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec.reshape(1, -1)
    world = dist.get_world_size(group)
    out = torch.empty((world, vec.numel()), dtype=vec.dtype, device=vec.device)
    dist.all_gather_into_tensor(out.reshape(-1), vec.reshape(-1).contiguous(), group=group)
    return out


def shard_tuples(mapping_lengths: Sequence[int], world: int) -> List[List[int]]:
    """Tuple blocks of one fit over `world` shards ("one large Gram"): greedy by the
    cubic cost of a block (its Cholesky factor, inverse and whitening), largest first."""
    cost = [float(m) ** 3 for m in mapping_lengths]
    order = sorted(range(len(cost)), key=lambda t: -cost[t])
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for t in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(t)
        load[r] += cost[t]
    return [sorted(x) for x in out]


class PeerExchange:
    """The count exchange of the sharded K1 path as kernels over peer memory (csrc/exchange.cu): every rank owns
    an exchange buffer that all ranks of the node map through CUDA IPC; `push` stores this rank's counts into
    its slot of every rank's buffer over NVLink and releases a flag, `pull` waits for the flags of all ranks and
    sums the slots (reduce) or lays them side by side (gather).  Replaces ncclAllReduce / ncclAllGather of the
    0.4 - 1.4 MB count vectors, whose latency (60 - 100 us on eight GPUs) is as long as a sharded repetition's
    parity kernel.  The handles travel once, through torch.distributed's object gather (host)."""

    def __init__(self, ctx, n_counts: int, group=None, peers=None):
        import ctypes as C

        import torch.distributed as dist

        self.ctx, self.n = ctx, int(n_counts)
        lib = ctx.lib
        if peers is not None:              # (rank, [device pointers]): ranks emulated inside one process (tests)
            self.rank, bases = peers
            self.world = len(bases)
            self._own, self._opened = None, []
        else:
            on = dist.is_available() and dist.is_initialized()
            self.world = dist.get_world_size(group) if on else 1
            self.rank = dist.get_rank(group) if on else 0
            nbytes = int(lib.aces_exchange_bytes(self.world, self.n))
            ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
            ctx.check(lib.aces_peer_alloc(ctx.handle, nbytes, C.byref(ptr), handle))
            self._own = ptr.value
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
            bases, self._opened = [], []
            for r in range(self.world):
                if r == self.rank:
                    bases.append(self._own)
                    continue
                q = C.c_void_p()
                buf = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                ctx.check(lib.aces_peer_open(ctx.handle, buf, C.byref(q)))
                bases.append(q.value)
                self._opened.append(q.value)
            if self.world > 1:
                dist.barrier(group=group)   # every rank has mapped every buffer before anyone pushes
        arr = (C.c_void_p * self.world)(*bases)
        h = C.c_void_p()
        ctx.check(lib.aces_exchange_create(ctx.handle, self.world, self.rank, self.n, arr, C.byref(h)))
        self.handle = h

    @staticmethod
    def buffer_bytes(ctx, world: int, n_counts: int) -> int:
        return int(ctx.lib.aces_exchange_bytes(world, n_counts))

    def push(self, local_counts_ptr: int, stream: int = None) -> None:
        self.ctx.check(self.ctx.lib.aces_exchange_push(self.ctx.handle, self.handle, local_counts_ptr, stream))

    def pull(self, out_ptr: int, reduce: bool, stream: int = None) -> None:
        self.ctx.check(self.ctx.lib.aces_exchange_pull(self.ctx.handle, self.handle, 1 if reduce else 0, out_ptr, stream))

    def attach(self, on: bool = True) -> None:
        """Fuses the push into K1: while attached, the prefix kernel of every asynchronous parity launch on this
        context publishes the counts to all ranks itself (aces_exchange_attach); only `pull` remains per step."""
        self.ctx.check(self.ctx.lib.aces_exchange_attach(self.ctx.handle, self.handle if on else None))

    def set_timeout(self, milliseconds: int) -> None:
        self.ctx.check(self.ctx.lib.aces_exchange_set_timeout(self.ctx.handle, self.handle, int(milliseconds)))

    def status(self) -> None:
        self.ctx.check(self.ctx.lib.aces_exchange_status(self.ctx.handle, self.handle))

    def close(self) -> None:
        lib = self.ctx.lib
        if getattr(self, "handle", None):
            lib.aces_exchange_destroy(self.ctx.handle, self.handle)
            self.handle = None
            for q in self._opened:
                lib.aces_peer_close(self.ctx.handle, q)
            if self._own:
                lib.aces_peer_free(self.ctx.handle, self._own)


MERIT_FARM_FIELDS = tuple(f"{ls}{est}_{what}" for ls in ("gls", "wls", "ols") for est in ("", "_marginal", "_relative")
                          for what in ("expectation", "variance"))


def merit_farm(merit_fn, gate_eigenvalue_sets, group=None) -> np.ndarray:
    """The repetition loop of calc_ensemble_scaling (src/scaling.jl:510-536: one calc_merit per random noise
    instance of the same design, up to 10^4 per distance) as a replica farm: instance r runs on rank
    r % world (`merit_fn(gate_eigenvalues) -> dict`, e.g. lambda g: merit.calc_merit(dd, gate_eigenvalues=g)
    on that rank's DeviceDesign), the 18 moments of every instance (MERIT_FARM_FIELDS) are gathered on every
    rank.  No data-path collective: one all_gather of [instances / world, 18] doubles at the end.  The
    convergence test of the reference (relative standard error of the GLS / WLS expectations) is host
    logic on the returned table."""
    import torch
    import torch.distributed as dist

    sets = [np.asarray(g, dtype=np.float64) for g in gate_eigenvalue_sets]
    R = len(sets)
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)
    per = -(-R // world)
    local = np.full((per, len(MERIT_FARM_FIELDS)), np.nan, dtype=np.float64)
    for slot, r in enumerate(shard_repetitions(R, world, rank)):
        m = merit_fn(sets[r])
        local[slot] = [m[k] for k in MERIT_FARM_FIELDS]
    if not on:
        return local[:R]
    backend = dist.get_backend(group)
    t = torch.from_numpy(local.reshape(-1))
    if backend == "nccl":
        t = t.cuda()
    allv = all_gather_vectors(t, group).cpu().numpy().reshape(world, per, len(MERIT_FARM_FIELDS))
    out = np.empty((R, len(MERIT_FARM_FIELDS)), dtype=np.float64)
    for r in range(R):
        out[r] = allv[r % world, r // world]
    return out


class _DeviceBuffer:
    """A raw device pointer as a __cuda_array_interface__ object (float64, 1-D)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def all_reduce_device(ptr: int, n: int, group=None) -> None:
    """In-place sum over the ranks of `n` doubles at device pointer `ptr` (NCCL over NVLink)."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()


class _DeviceBufferI32:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def broadcast_device(ptr: int, n: int, src: int, group=None, sync: bool = True) -> None:
    """Broadcast of `n` doubles at device pointer `ptr` from rank `src` of `group` (ncclBroadcast over NVLink)."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda")
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    if sync:
        torch.cuda.current_stream().synchronize()


def all_reduce_device_i32(ptr: int, group=None) -> None:
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(_DeviceBufferI32(ptr, 1), device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()


def _staging(dd, name: str, n: int):
    """A float64 device buffer of at least n elements cached on the DeviceDesign (packing area of the collectives)."""
    import torch

    t = getattr(dd, name, None)
    if t is None or t.numel() < n:
        t = torch.empty(n, dtype=torch.float64, device="cuda")
        setattr(dd, name, t)
    return t


def fit_tuple_sharded(dd, ls_type: str, est_eigenvalues, est_covariance_eigenvalues=None, group=None,
                      distributed_cholesky: bool = False, lookahead: bool = True, **kw):
    """One fit with the tuple blocks sharded over the ranks of `group`: block factors and partial
    Gram matrices per rank, ncclAllReduce of the Gram matrix and the right-hand sides, then the Cholesky
    factorisation either replicated or -- `distributed_cholesky` -- distributed by block columns with
    ncclBroadcast of every factorised panel (DeviceDesign.factor_distributed); the triangular solves are
    replicated.  `dd` is the rank's DeviceDesign; the shard is set here.

    With `distributed_cholesky` the library's kernels and the collectives are ordered on the device (torch
    streams and events: no host synchronisation per panel); with `lookahead` the broadcasts run on a second
    stream and the next block column is factorised and sent while the current panel is still being applied."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = shard_tuples(dd.fd.mapping_lengths, world)[rank]
    on = [1 if t in mine else 0 for t in range(dd.fd.tuple_number)]
    dd.set_tuple_shard(on)
    try:
        if not distributed_cholesky:
            return dd.fit_sharded(ls_type, est_eigenvalues, est_covariance_eigenvalues,
                                  lambda p, n: all_reduce_device(p, n, group), **kw)
        stream = getattr(dd, "_dist_stream", None)
        if stream is None:
            stream = dd._dist_stream = torch.cuda.Stream()
            dd._comm_stream = torch.cuda.Stream()
        comm = dd._comm_stream
        dd.ctx.set_stream(stream.cuda_stream)
        try:
            with torch.cuda.stream(stream):
                pw = dd.DIST_PANEL

                def columns(p, Np):   # the column-major matrix as [column][row]
                    return torch.as_tensor(_DeviceBuffer(p, Np * Np), device="cuda").view(Np, Np)

                def reduce(p, n):
                    t = torch.as_tensor(_DeviceBuffer(p, n), device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

                def reduce_lower(p, Np):
                    # The partial Gram matrices only hold their lower triangle: pack the lower trapezoid of every
                    # block column (rows col0 .. Np) into one contiguous buffer, ONE all-reduce of half the bytes,
                    # unpack.  (The strict upper parts inside the diagonal blocks travel too: 1 % at d = 17.)
                    G = columns(p, Np)
                    pieces = [(c, min(pw, Np - c)) for c in range(0, Np, pw)]
                    total = sum(n * (Np - c) for c, n in pieces)
                    stage = _staging(dd, "_gram_stage", total)
                    off = 0
                    for c, n in pieces:
                        stage[off:off + n * (Np - c)].view(n, Np - c).copy_(G[c:c + n, c:])
                        off += n * (Np - c)
                    dist.all_reduce(stage[:total], op=dist.ReduceOp.SUM, group=group)
                    off = 0
                    for c, n in pieces:
                        G[c:c + n, c:].copy_(stage[off:off + n * (Np - c)].view(n, Np - c))
                        off += n * (Np - c)

                reduce.lower = reduce_lower

                def bcast(p, n, src):
                    broadcast_device(p, n, src, group, sync=False)

                def bcast_panel(p, Np, col0, ncols, src):
                    view = columns(p, Np)[col0:col0 + ncols, col0:]
                    stage = _staging(dd, "_panel_stage", pw * Np)[:ncols * (Np - col0)].view(ncols, Np - col0)
                    if rank == src:
                        stage.copy_(view)
                    dist.broadcast(stage, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
                    if rank != src:
                        view.copy_(stage)

                bcast.panel = bcast_panel

                def reduce_int(p):
                    t = torch.as_tensor(_DeviceBufferI32(p, 1), device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

                return dd.fit_sharded(ls_type, est_eigenvalues, est_covariance_eigenvalues, reduce,
                                      distributed=(rank, world, bcast, reduce_int, (stream, comm) if lookahead else None), **kw)
        finally:
            stream.synchronize()
            dd.ctx.set_stream(None)
    finally:
        dd.set_tuple_shard(None)
