"""CPU test of the static task order of the persistent Cholesky kernel (aces_potrf_schedule): the order
must be a topological order of the tile DAG under exactly the dependency counters the kernel waits on --
executing the list sequentially must never find an unmet dependency (that is the kernel's deadlock-freedom
argument) -- and must contain every tile task exactly once."""
import ctypes as C

import numpy as np
import pytest


def schedule(nb, workers):
    import aces_b200

    lib = aces_b200._lib.load()
    n = C.c_int64()
    assert lib.aces_potrf_schedule(nb, workers, None, 0, C.byref(n)) == 0
    out = np.zeros((n.value, 6), dtype=np.int32)
    assert lib.aces_potrf_schedule(nb, workers, out.ctypes.data, n.value, C.byref(n)) == 0
    return out


@pytest.mark.parametrize("nb,workers", [(1, 148), (2, 148), (5, 20), (9, 148), (53, 148), (53, 19), (70, 148)])
def test_schedule_never_deadlocks_and_covers_the_dag(built, nb, workers):
    """Emulates the kernel's protocol: the chain CTA runs F(k) as soon as block k has all its updates; ONE
    CTA of the near pool and ONE of the far pool walk their lists in order and wait for the head's
    dependencies.  If no list head is ready while tasks remain, the kernel would spin for ever.  Every task
    must run exactly once with exactly the dependency counters the kernel waits on."""
    tasks = schedule(nb, workers)
    flist, lists = tasks[-nb:], tasks[:-nb]
    assert np.all(flist[:, 0] == 0) and np.array_equal(flist[:, 3], np.arange(nb)) and not np.any(lists[:, 0] == 0)
    cls = lists[:, 1] >> 8
    lists[:, 1] &= 255
    assert np.all(np.diff(cls) >= 0) and cls.max(initial=0) <= 1           # [far | near]
    queues = [lists[cls == c] for c in range(2)]
    assert np.all(queues[0][:, 0] <= 2)                                    # every slab task is near
    cnt = np.zeros((nb, nb), dtype=np.int64)
    sdone = np.zeros((nb, nb), dtype=np.int64)
    fdone = np.zeros(nb, dtype=np.int64)
    next_f = 0

    def ready(tk):
        typ, _, i, j, k0, k1 = (int(x) for x in tk)
        if typ in (1, 3):
            return fdone[j] == 1 and cnt[i, j] >= 8 * j
        return cnt[i, j] >= 8 * k0 and all(sdone[i, k] == 8 and sdone[j, k] == 8 for k in range(k0, k1))

    def run(tk):
        typ, _, i, j, k0, k1 = (int(x) for x in tk)
        if typ in (1, 3):
            assert i > j and cnt[i, j] == 8 * j
            sdone[i, j] += 8 if typ == 1 else 1
            assert sdone[i, j] <= 8
        else:
            assert typ in (2, 4) and i >= j and 0 <= k0 < k1 <= j and k1 - k0 <= 4
            assert (cnt[i, j] == 8 * k0) if typ == 2 else (8 * k0 <= cnt[i, j] < 8 * k1)
            cnt[i, j] += 8 * (k1 - k0) if typ == 2 else (k1 - k0)

    pos = [0] * 2
    while next_f < nb or any(pos[c] < len(queues[c]) for c in range(2)):
        progressed = False
        while next_f < nb and cnt[next_f, next_f] == 8 * next_f:
            fdone[next_f] = 1
            next_f += 1
            progressed = True
        for c in range(2):
            while pos[c] < len(queues[c]) and ready(queues[c][pos[c]]):
                run(queues[c][pos[c]])
                pos[c] += 1
                progressed = True
        assert progressed, f"deadlock: list positions {pos}, F {next_f}/{nb}"
    for j in range(nb):
        for i in range(j, nb):
            assert cnt[i, j] == 8 * j
            if i > j:
                assert sdone[i, j] == 8


def test_critical_tiles_are_split(built):
    nb = 53
    tasks = schedule(nb, 148)[:-nb]
    # S(k+1, k), S(k+2, k) run as eight 16-row slabs; so do the recent updates of tiles (j, j) and (j+1, j),
    # the very last k-block on its own
    assert np.sum(tasks[:, 0] == 3) == 8 * ((nb - 1) + (nb - 2))
    slabs = tasks[tasks[:, 0] == 4]
    last = slabs[slabs[:, 5] == slabs[:, 3]]
    assert np.all(last[:, 5] - last[:, 4] == 1) and len(last) == 8 * ((nb - 1) + (nb - 2))
    assert np.all((tasks[tasks[:, 0] >= 3][:, 1] >> 8) == 1)       # every slab task is in the near list
