"""CPU tests of the multi-GPU host logic (quantumaces.jl_b200/shard.py): world_size-2 runs over the
gloo backend.  Each rank reduces its shard with the CPU oracle (there is no GPU here), the shards
are combined with the same collectives bench.py / a multi-GPU run use, and the result must equal
the unsharded counts bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import aces_b200
import oracle_bindings as ob
from aces_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True, stress_weights=True)
    exp_ptr, pauli_ptr, pauli_bits, _ = aces_b200.build_pauli_tables(fd)
    shots_divided, shots_max = aces_b200.shots_divide(fd, [3_000, 20_000, 90_001])
    tuple_of = np.repeat(np.arange(fd.tuple_number), fd.experiment_numbers)
    rng = np.random.default_rng(2)
    mats, prefix, sups = [], [], []
    for e in range(fd.experiment_number):
        t = int(tuple_of[e])
        mats.append(np.asfortranarray(rng.integers(0, 256, size=(int(shots_max[t]), fd.n_bytes), dtype=np.uint8)))
        prefix.append(shots_divided[t])
        sups.append([(pauli_bits[pauli_ptr[p]:pauli_ptr[p + 1]] + 1).tolist()
                     for p in range(int(exp_ptr[e]), int(exp_ptr[e + 1]))])
    return fd, mats, np.asarray(prefix, dtype=np.int64), sups


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import merit_oracle as mo  # noqa: E402


def _odd(mat, sup, prefix):
    return ob.simulate_estimate(mat, sup, np.zeros(len(sup), dtype=np.uint8), prefix)[1].reshape(-1)


def _worker(rank, world, port, mode, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fd, mats, prefix, sups = _problem()
        sizes = [len(s) * prefix.shape[1] for s in sups]
        off = np.concatenate([[0], np.cumsum(sizes)])
        counts = torch.zeros(int(off[-1]), dtype=torch.int64)
        if mode == "experiments":
            mine = shard.shard_experiments([m.shape[0] * m.shape[1] for m in mats], world)[rank]
            for e in mine:
                counts[off[e]:off[e + 1]] = torch.from_numpy(_odd(mats[e], sups[e], prefix[e]))
            shard.all_reduce_counts(counts)
        elif mode == "shots":
            for e in range(len(mats)):
                lo, hi, local = shard.shard_shots(mats[e].shape[0], prefix[e], world, rank, align=128)
                if hi > lo:
                    counts[off[e]:off[e + 1]] = torch.from_numpy(_odd(mats[e][lo:hi], sups[e], local))
            shard.all_reduce_counts(counts)
        else:  # repetitions: every rank reduces whole repetitions, the vectors are gathered
            reps = 4
            mine = shard.shard_repetitions(reps, world, rank)
            assert len(mine) == reps // world
            per_rep = []
            for r in mine:
                v = torch.zeros(int(off[-1]), dtype=torch.int64)
                for e in range(len(mats)):
                    rolled = np.asfortranarray(np.roll(mats[e], r, axis=0))
                    v[off[e]:off[e + 1]] = torch.from_numpy(_odd(rolled, sups[e], prefix[e]))
                per_rep.append(shard.all_gather_vectors(v))  # [world, len]: repetitions r0 + rank
            counts = torch.stack(per_rep)  # [reps / world, world, len]
        if rank == 0:
            ret["counts"] = counts.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["experiments", "shots", "repetitions"])
def test_world_size_2_gloo(mode):
    import torch.multiprocessing as mp

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, mode, ret), nprocs=world, join=True)
    got = ret["counts"]
    fd, mats, prefix, sups = _problem()
    if mode == "repetitions":
        for i in range(got.shape[0]):
            for w in range(world):
                r = i * world + w
                want = np.concatenate([_odd(np.asfortranarray(np.roll(mats[e], r, axis=0)), sups[e], prefix[e])
                                       for e in range(len(mats))])
                assert np.array_equal(got[i, w], want)
    else:
        want = np.concatenate([_odd(mats[e], sups[e], prefix[e]) for e in range(len(mats))])
        assert np.array_equal(got, want)


def test_shard_helpers():
    assert shard.shard_repetitions(20, 8, 3) == [3, 11, 19]
    parts = shard.shard_experiments([9, 1, 1, 1, 8, 2, 2, 7], 3)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum([9, 1, 1, 1, 8, 2, 2, 7][e] for e in p) for p in parts]
    assert max(loads) - min(loads) <= 2
    rows, pre = 1_041_667, [10_417, 104_167, 1_041_667]
    total = np.zeros(3, dtype=np.int64)
    covered = 0
    for r in range(8):
        lo, hi, local = shard.shard_shots(rows, pre, 8, r)
        assert lo % 128 == 0 and 0 <= lo <= hi <= rows
        assert np.all(np.diff(local) >= 0) and local[-1] == hi - lo
        total += local
        covered += hi - lo
    assert covered == rows and np.array_equal(total, pre)
    # more ranks than rows: empty shards are valid
    assert [shard.shard_shots(5, [5], 4, r)[:2] for r in range(4)] == [(0, 5), (5, 5), (5, 5), (5, 5)]


def test_shard_tuples_partition_is_balanced():
    from aces_b200 import shard

    m = [483, 483, 1131, 483, 1131, 483, 1131, 1131, 483, 483, 483, 483, 1131, 1131, 1131, 1131]   # rotated d=9
    for world in (1, 2, 3, 4, 8):
        parts = shard.shard_tuples(m, world)
        assert sorted(t for p in parts for t in p) == list(range(len(m)))
        cost = [sum(float(m[t]) ** 3 for t in p) for p in parts]
        assert max(cost) <= sum(cost) / world + max(float(x) ** 3 for x in m)


def _merit_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fd, sets = _merit_problem()
        table = shard.merit_farm(lambda g: mo.calc_merit(fd, gate_eigenvalues=g), sets)
        if rank == 0:
            ret["table"] = table
    finally:
        dist.destroy_process_group()


def _merit_problem():
    import aces_b200

    fd = aces_b200.example_design(full_covariance=True)
    rng = np.random.default_rng(8)
    g0 = np.asarray(fd.gate_eigenvalues, dtype=np.float64)
    return fd, [1.0 - (1.0 - g0) * rng.uniform(0.5, 1.5, size=len(g0)) for _ in range(5)]


def test_merit_farm_world_size_2_gloo():
    """calc_ensemble_scaling's repetition loop as replicas: 5 noise instances over 2 ranks
    (uneven: 3 + 2), every rank ends up with the whole table, equal to the serial one."""
    import torch.multiprocessing as mp

    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_merit_worker, args=(2, port, ret), nprocs=2, join=True)
    fd, sets = _merit_problem()
    serial = shard.merit_farm(lambda g: mo.calc_merit(fd, gate_eigenvalues=g), sets)
    assert serial.shape == (5, 18) and np.all(np.isfinite(serial))
    assert np.array_equal(ret["table"], serial)
