"""GPU test of the peer-memory count exchange (csrc/exchange.cu): ranks emulated inside one process as contexts on
one GPU whose exchange buffers are plain device allocations (across processes they are mapped through CUDA IPC and
the stores travel over NVLink: tools / bench.py under torchrun).  Push and pull kernels, flags, slot parities."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n", [(2, 7), (3, 1000), (8, 48097)])
def test_push_pull_reduce_and_gather(ctx, world, n):
    import torch

    import aces_b200
    from aces_b200 import shard

    ctxs = [aces_b200.Context(0) for _ in range(world)]
    nbytes = shard.PeerExchange.buffer_bytes(ctx, world, n)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    xs = [shard.PeerExchange(ctxs[r], n, peers=(r, [b.data_ptr() for b in bufs])) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(5)
    for step in range(5):                      # several steps: both slot parities, epochs, ticket reset
        local = [torch.from_numpy(rng.integers(0, 2 ** 40, size=n, dtype=np.int64)).cuda() for _ in range(world)]
        out_sum = [torch.full((n,), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        reduce = step % 2 == 0
        out = out_sum if reduce else [torch.full((world * n,), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
        # pulls are queued FIRST on some ranks: they must wait for the pushes of the others (different streams)
        order = list(range(world))
        for r in order[::2]:
            xs[r].push(local[r].data_ptr(), streams[r].cuda_stream)
        for r in order:
            if r % 2 == 1:
                xs[r].push(local[r].data_ptr(), streams[r].cuda_stream)
            xs[r].pull(out[r].data_ptr(), reduce, streams[r].cuda_stream)
        torch.cuda.synchronize()
        want = torch.stack(local).sum(dim=0) if reduce else torch.cat(local)
        for r in range(world):
            xs[r].status()
            assert torch.equal(out[r], want)
    for x in xs:
        x.close()
    for c in ctxs:
        c.close()


def test_pull_without_a_peer_times_out_instead_of_hanging(ctx):
    import torch

    import aces_b200
    from aces_b200 import shard

    n = 100
    bufs = [torch.zeros(shard.PeerExchange.buffer_bytes(ctx, 2, n), dtype=torch.uint8, device="cuda") for _ in range(2)]
    x0 = shard.PeerExchange(ctx, n, peers=(0, [b.data_ptr() for b in bufs]))
    x0.set_timeout(300)
    local = torch.ones(n, dtype=torch.int64, device="cuda")
    out = torch.zeros(n, dtype=torch.int64, device="cuda")
    x0.push(local.data_ptr())
    x0.pull(out.data_ptr(), True)          # rank 1 never pushes
    torch.cuda.synchronize()
    with pytest.raises(aces_b200.AcesError):
        x0.status()
    x0.close()


def test_push_fused_into_the_prefix_kernel(ctx):
    """aces_exchange_attach: K1's prefix kernel publishes the counts to every rank as it forms them (no separate
    push launch).  Two emulated ranks reduce different sample matrices of the same design; after the pull each
    holds the sum of both count vectors, bit for bit, over several steps (both slot parities)."""
    import torch

    import aces_b200
    from aces_b200 import shard

    fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True)
    world = 2
    ctxs = [aces_b200.Context(0) for _ in range(world)]
    ests = [aces_b200.DesignEstimator(c, fd) for c in ctxs]
    shots_divided, shots_max = aces_b200.shots_divide(fd, [3000, 20000])
    n_exp = fd.experiment_number
    rows = [int(shots_max[int(ests[0].tuple_of_experiment[e])]) for e in range(n_exp)]
    prefix = np.stack([shots_divided[int(ests[0].tuple_of_experiment[e])] for e in range(n_exp)])
    ld = [(r + 127) // 128 * 128 for r in rows]
    n_counts = int(np.sum(ests[0].experiment_sizes()) * prefix.shape[1])
    nbytes = shard.PeerExchange.buffer_bytes(ctx, world, n_counts)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    xs = [shard.PeerExchange(ctxs[r], n_counts, peers=(r, [b.data_ptr() for b in bufs])) for r in range(world)]
    gen = torch.Generator(device="cuda").manual_seed(3)
    for step in range(4):
        mats = [[torch.randint(0, 256, (fd.n_bytes, ld[e]), dtype=torch.uint8, device="cuda", generator=gen) for e in range(n_exp)]
                for _ in range(world)]
        plain = []
        for r in range(world):        # reference: the same launches without the exchange
            o = torch.zeros(n_counts, dtype=torch.int64, device="cuda")
            ests[r].prepare([m.data_ptr() for m in mats[r]], rows, ld, prefix, o.data_ptr(), asynchronous=True).launch()
            ctxs[r].synchronize()
            plain.append(o)
        launches = [c.kernel_launches() for c in ctxs]
        outs, sums = [], []
        for r in range(world):
            xs[r].attach()
            o = torch.zeros(n_counts, dtype=torch.int64, device="cuda")
            ests[r].prepare([m.data_ptr() for m in mats[r]], rows, ld, prefix, o.data_ptr(), asynchronous=True).launch()
            outs.append(o)
        for r in range(world):
            s = torch.full((n_counts,), -1, dtype=torch.int64, device="cuda")
            xs[r].pull(s.data_ptr(), True)
            sums.append(s)
        for r in range(world):
            ctxs[r].synchronize()
            xs[r].status()
            xs[r].attach(False)
            assert ctxs[r].kernel_launches() - launches[r] == 3          # parity + prefix(+push) + pull: no push launch
            assert torch.equal(outs[r], plain[r])                        # the local counts are unchanged
            assert torch.equal(sums[r], plain[0] + plain[1])
    for x in xs:
        x.close()
    for c in ctxs:
        c.close()
