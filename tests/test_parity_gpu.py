"""GPU parity tests for K1: the CUDA kernel (through the C ABI) against the CPU oracle, the
golden fixture and size-independent properties.  Integer work: the bar is bit-exact."""
import numpy as np
import pytest

import oracle_bindings as ob
from golden_util import parity_cases

pytestmark = pytest.mark.gpu


def rand_supports(rng, n, E, max_w=9):
    sup = []
    for p in range(E):
        w = int(rng.choice([1, 1, 1, 2, 2, 2, 3, 4, 5, 7, max_w]))
        sup.append(sorted((rng.choice(n, size=min(w, n), replace=False) + 1).tolist()))
    return sup


def test_golden_fixture(ctx):
    import aces_b200

    for case in parity_cases():
        for k, S in enumerate(case["prefixes"]):
            est = aces_b200.estimate_experiment(ctx, case["counts"], case["supports"], case["signs"], int(S))
            assert np.array_equal(est, case["est"][k])


@pytest.mark.parametrize("n,rows", [(17, 1), (17, 15), (17, 16), (17, 4097), (49, 1023), (49, 70001),
                                    (161, 1025), (161, 33333), (577, 257), (577, 5000), (1200, 3000)])
def test_estimate_experiment_vs_oracle(ctx, n, rows):
    import aces_b200

    rng = np.random.default_rng(n * 1000 + rows)
    B = (n + 7) // 8
    counts = np.asfortranarray(rng.integers(0, 256, size=(rows, B), dtype=np.uint8))
    sup = rand_supports(rng, n, 97)
    sup[3] = []                       # identity: eigenvalue (+-)1
    sup[4] = [sup[5][0], sup[5][0]]   # repeated qubit cancels
    signs = rng.integers(0, 2, size=len(sup))
    for S in sorted({1, rows, max(1, rows // 3)}):
        got = aces_b200.estimate_experiment(ctx, counts, sup, signs, S)
        want, _ = ob.estimate_experiment(counts, sup, signs, S)
        assert np.array_equal(got, want)


def test_row_major_input_is_accepted(ctx):
    import aces_b200

    rng = np.random.default_rng(3)
    counts = rng.integers(0, 256, size=(777, 7), dtype=np.uint8)  # C order, as numpy/Stim produce
    sup = rand_supports(rng, 49, 20)
    got = aces_b200.estimate_experiment(ctx, counts, sup, [0] * 20, 777)
    want, _ = ob.estimate_experiment(counts, sup, [0] * 20, 777)
    assert np.array_equal(got, want)


def test_noiseless_known_answer(ctx):
    # test/device_tests.jl:144-155 through the CUDA path
    import aces_b200

    rec = np.zeros((12345, 21), dtype=np.uint8, order="F")
    sup = [[1], [2, 3], [161], [5, 6, 7, 8], []]
    assert np.all(aces_b200.estimate_experiment(ctx, rec, sup, [0] * 5, 12345) == 1.0)
    got = aces_b200.estimate_experiment(ctx, rec, sup, [1, 0, 1, 0, 1], 12345)
    assert np.array_equal(got, np.array([-1.0, 1.0, -1.0, 1.0, -1.0]))


def test_error_behaviour(ctx):
    import aces_b200

    rec = np.zeros((10, 3), dtype=np.uint8, order="F")
    with pytest.raises(AssertionError):      # src/estimate.jl:292
        aces_b200.estimate_experiment(ctx, rec, [[1], [2]], [0], 10)
    with pytest.raises(aces_b200.AcesError):  # qubit outside the record (Julia: BoundsError)
        aces_b200.estimate_experiment(ctx, rec, [[25]], [0], 10)
    with pytest.raises(aces_b200.AcesError):  # more shots than rows (Julia: BoundsError)
        aces_b200.estimate_experiment(ctx, rec, [[1]], [0], 11)


def make_batch(rng, est, fd, shots_set, pad_ld=0):
    import aces_b200

    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    mats, rows, ld, prefix = [], [], [], []
    for e in range(fd.experiment_number):
        t = int(est.tuple_of_experiment[e])
        S = int(shots_max[t])
        full = np.asfortranarray(rng.integers(0, 256, size=(S + pad_ld, fd.n_bytes), dtype=np.uint8))
        mats.append(full[:S] if pad_ld else full)
        rows.append(S)
        ld.append(S + pad_ld)
        prefix.append(shots_divided[t])
    return mats, rows, ld, np.asarray(prefix, dtype=np.int64)


@pytest.mark.parametrize("d,gls,stress", [(3, False, False), (3, True, True), (5, True, False),
                                          ("real3", True, False), ("real5", True, False), ("real5", False, False),
                                          ("unrot3", True, False),
                                          # the headline shapes on the reference's own tables: rotated d=9 WLS
                                          # (RMAX=8 / BMAX=21), d=9 GLS (449 Paulis: RMAX=16 / BMAX=21), unrotated
                                          # d=7 (22 byte columns: BMAX=25), rotated d=17 (73 columns: column groups)
                                          ("real9", False, False), ("real9", True, False), ("unrot7", True, False),
                                          ("unrot7", False, False), ("real17", False, False)])
def test_design_batch_vs_oracle(ctx, d, gls, stress):
    import aces_b200

    shots_set = [3_000, 41_111, 200_003]
    if isinstance(d, str):   # real designs (circuit_design.py): the reference's own tables
        kind, d = d.rstrip("0123456789"), int(d[len(d.rstrip("0123456789")):])
        fd = (aces_b200.rotated_design if kind == "real" else aces_b200.unrotated_design)(d, full_covariance=gls)
        if d >= 7:               # ~100 experiments: about 2e4 shots each (several tiles per stream)
            shots_set = [30_000, 411_111, 2_000_003]
    else:
        fd = aces_b200.synthetic_design("rotated", d, full_covariance=gls, stress_weights=stress)
    est = aces_b200.DesignEstimator(ctx, fd)
    rng = np.random.default_rng(d)
    mats, rows, ld, prefix = make_batch(rng, est, fd, shots_set)
    got = est.odd_counts([m.ctypes.data for m in mats], rows, ld, prefix)
    want = ob.oracle_odd_counts(est, mats, prefix)
    assert np.array_equal(got, want)
    # prefix properties: nondecreasing in k and bounded by the prefix length
    off = 0
    for e in range(fd.experiment_number):
        E = est.tables.experiment_size(e)
        blk = got[off:off + 3 * E].reshape(3, E)
        assert np.all(np.diff(blk, axis=0) >= 0)
        assert np.all(blk <= prefix[e][:, None])
        off += 3 * E


def test_simulate_estimate_matches_reference_nesting(ctx):
    import aces_b200

    fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True)
    est = aces_b200.DesignEstimator(ctx, fd)
    rng = np.random.default_rng(9)
    shots_set = [10_000, 2_000, 77_777]  # deliberately unsorted budgets
    mats, rows, ld, prefix = make_batch(rng, est, fd, shots_set)
    eig_set, cov_set = est.simulate_estimate(mats, shots_set)
    shots_divided, _ = aces_b200.shots_divide(fd, shots_set)
    # restate src/simulate.jl:195-234 with the oracle
    e = 0
    for t in range(fd.tuple_number):
        want_eig = [[[] for _ in range(int(fd.mapping_lengths[t]))] for _ in shots_set]
        want_cov = [[[] for _ in range(len(fd.cov_keys[t]))] for _ in shots_set]
        for j in range(int(fd.experiment_numbers[t])):
            exp = fd.experiment_ensemble[t][j]
            cexp = fd.cov_experiment_ensemble[t][j]
            for s in range(len(shots_set)):
                S = int(shots_divided[t, s])
                ee, _ = ob.estimate_experiment(mats[e], [fd.meas_indices[t][p] for p in exp],
                                               fd.pauli_signs[t][exp], S)
                for idx, p in enumerate(exp):
                    want_eig[s][int(p)].append(ee[idx])
                ce, _ = ob.estimate_experiment(mats[e], [fd.cov_meas_indices[t][c] for c in cexp],
                                               fd.cov_signs[t][cexp], S)
                for idx, c in enumerate(cexp):
                    want_cov[s][int(c)].append(ce[idx])
            e += 1
        for s in range(len(shots_set)):
            assert eig_set[s][t] == want_eig[s]
            assert cov_set[s][t] == want_cov[s]


def test_prefix_edge_cases(ctx):
    import aces_b200

    rng = np.random.default_rng(21)
    n, rows = 49, 5000
    sup = rand_supports(rng, n, 40)
    ptr, flat = ob._csr(sup)
    tables = aces_b200.PauliTables(ctx, n, [0, 40], ptr, flat - 1)
    m = np.asfortranarray(rng.integers(0, 256, size=(rows, 7), dtype=np.uint8))
    for prefixes in ([0, 0, 5000], [1, 1, 1], [1024, 2048, 4096], [1023, 1025, 4999], [5000], [0],
                     [17, 17, 17, 900, 5000]):
        got = aces_b200._lib.parity_counts(ctx, tables, [0], [m], [rows], [rows], [prefixes])
        K = len(prefixes)
        want = np.zeros((K, 40), dtype=np.int64)
        for k, S in enumerate(prefixes):
            if S > 0:
                _, ps = ob.estimate_experiment(m, sup, [0] * 40, S)
                want[k] = (S - ps) // 2
        assert np.array_equal(got.reshape(K, 40), want)
    with pytest.raises(aces_b200.AcesError):
        aces_b200._lib.parity_counts(ctx, tables, [0], [m], [rows], [rows], [[10, 5]])
    with pytest.raises(aces_b200.AcesError):
        aces_b200._lib.parity_counts(ctx, tables, [0], [m], [rows], [rows], [[5001]])
    # empty batch and zero-row matrix
    got = aces_b200._lib.parity_counts(ctx, tables, [0], [m], [0], [0], [[0, 0]])
    assert np.array_equal(got, np.zeros(80, dtype=np.int64))


def test_batched_accumulate_equals_concatenated(ctx):
    # src/simulate.jl:174-186: batches are vcat'ed; accumulating per-batch counts must agree
    import aces_b200

    rng = np.random.default_rng(33)
    n = 161
    sup = rand_supports(rng, n, 300)
    ptr, flat = ob._csr(sup)
    tables = aces_b200.PauliTables(ctx, n, [0, 300], ptr, flat - 1)
    batches = ob.batch_shots(10_000, n, 600_000)
    assert len(batches) > 1
    mats = [np.asfortranarray(rng.integers(0, 256, size=(int(b), 21), dtype=np.uint8)) for b in batches]
    whole = np.asfortranarray(np.concatenate(mats, axis=0))
    prefixes = np.array([700, 4100, 10_000])
    want = aces_b200._lib.parity_counts(ctx, tables, [0], [whole], [10_000], [10_000], [prefixes])
    acc = np.zeros(3 * 300, dtype=np.int64)
    start = 0
    for mtx in mats:
        local = np.clip(prefixes - start, 0, mtx.shape[0])
        aces_b200._lib.parity_counts(ctx, tables, [0], [mtx], [mtx.shape[0]], [mtx.shape[0]], [local],
                                     out=acc, accumulate=True)
        start += mtx.shape[0]
    assert np.array_equal(acc, want)
    _, odd = ob.simulate_estimate(whole, sup, [0] * 300, prefixes)
    assert np.array_equal(want.reshape(3, 300), odd)


@pytest.mark.parametrize("E", [513, 1500, 2049, 5000])
def test_many_paulis_per_experiment(ctx, E):
    # more Paulis than threads: several per thread (PPT 2/4) and Pauli slices beyond 2048
    import aces_b200

    rng = np.random.default_rng(E)
    n, rows = 97, 2100
    sup = rand_supports(rng, n, E)
    ptr, flat = ob._csr(sup)
    tables = aces_b200.PauliTables(ctx, n, [0, E], ptr, flat - 1)
    m = np.asfortranarray(rng.integers(0, 256, size=(rows, 13), dtype=np.uint8))
    got = aces_b200._lib.parity_counts(ctx, tables, [0], [m], [rows], [rows], [[100, rows]])
    _, odd = ob.simulate_estimate(m, sup, [0] * E, [100, rows])
    assert np.array_equal(got.reshape(2, E), odd)


def test_device_resident_inputs(ctx):
    # device matrices: in place (aligned ld) and through the pitched copy (odd ld); device output
    import torch

    import aces_b200

    rng = np.random.default_rng(55)
    n, E = 161, 241
    sup = rand_supports(rng, n, E)
    ptr, flat = ob._csr(sup)
    tables = aces_b200.PauliTables(ctx, n, [0, E, 2 * E], np.concatenate([ptr, ptr[1:] + ptr[-1]]),
                                   np.concatenate([flat, flat]) - 1)
    for rows, ld in [(40_000, 40_000), (40_001, 40_016), (39_999, 39_999), (12_345, 20_000)]:
        host = [np.asfortranarray(rng.integers(0, 256, size=(ld, 21), dtype=np.uint8)) for _ in range(2)]
        dev = [torch.from_numpy(np.ascontiguousarray(h.T)).cuda() for h in host]  # [21, ld] row-major == col-major [ld,21]
        prefixes = [[rows // 7, rows // 2, rows], [1, 2, rows]]
        out = torch.zeros(2 * 3 * E, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        aces_b200._lib.parity_counts(ctx, tables, [0, 1], [d.data_ptr() for d in dev], [rows, rows],
                                     [ld, ld], prefixes, out=out.data_ptr(), asynchronous=True)
        ctx.synchronize()
        got = out.cpu().numpy()
        want = np.concatenate([
            ob.simulate_estimate(host[i][:rows], sup, [0] * E, prefixes[i])[1].reshape(-1) for i in range(2)
        ])
        assert np.array_equal(got, want)


@pytest.mark.parametrize("source", ["real", "synthetic"])
def test_full_size_properties(ctx, source):
    """BASELINE config 3 size (rotated d=9, 1e8 shots, 2.1 GB): size-independent properties."""
    import torch

    import aces_b200

    fd = aces_b200.rotated_design(9, full_covariance=False) if source == "real" else aces_b200.synthetic_design("rotated", 9)
    est = aces_b200.DesignEstimator(ctx, fd)
    shots_set = [10**6, 10**7, 10**8]
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    rows = [int(shots_max[int(est.tuple_of_experiment[e])]) for e in range(n_exp)]
    ld = [(r + 127) // 128 * 128 for r in rows]
    g = torch.Generator(device="cuda").manual_seed(1234)
    dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
    prefix = np.stack([shots_divided[int(est.tuple_of_experiment[e])] for e in range(n_exp)])
    total = int(np.sum(est.experiment_sizes()) * 3)
    out = torch.zeros(total, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def run():
        est.odd_counts([d.data_ptr() for d in dev], rows, ld, prefix, out=out.data_ptr(), asynchronous=True)
        ctx.synchronize()
        return out.cpu().numpy().copy()

    got = run()
    # (1) independent check of a sample of Paulis with plain torch bit arithmetic
    rng = np.random.default_rng(0)
    off = np.concatenate([[0], np.cumsum(est.experiment_sizes() * 3)])
    for e in rng.choice(n_exp, size=6, replace=False):
        E = est.tables.experiment_size(int(e))
        sup = ob.table_supports(est, int(e))
        for p in rng.choice(E, size=5, replace=False):
            acc = torch.zeros(rows[e], dtype=torch.uint8, device="cuda")
            for q in sup[p]:
                acc ^= (dev[e][(q - 1) // 8, :rows[e]] >> ((q - 1) % 8)) & 1
            csum = torch.cumsum(acc.to(torch.int64), 0)
            for k in range(3):
                S = int(prefix[e, k])
                assert got[off[e] + k * E + p] == int(csum[S - 1])
    # (2) idempotence / determinism of the atomics-based reduction
    assert np.array_equal(got, run())
    # (3) complement: flipping every record bit maps odd -> S - odd for odd-weight Paulis and
    #     leaves even-weight Paulis unchanged
    for d in dev:
        d.bitwise_not_()
    torch.cuda.synchronize()  # the library runs on its own stream
    flipped = run()
    for e in range(n_exp):
        E = est.tables.experiment_size(e)
        w = np.array([len(s) for s in ob.table_supports(est, e)])
        a = got[off[e]:off[e + 1]].reshape(3, E)
        b = flipped[off[e]:off[e + 1]].reshape(3, E)
        S = prefix[e][:, None]
        assert np.array_equal(np.where(w % 2 == 1, S - a, a), b)
    # (4) all-zero records: no odd parities at all (noiseless known answer at full size)
    for d in dev:
        d.zero_()
    torch.cuda.synchronize()
    assert not run().any()


def test_d17_shot_shard_properties(ctx):
    """BASELINE config 5 (rotated d=17, 1e9 shots, 73 GB) is sharded by shot ranges over 8 GPUs:
    this runs the 9.1 GB shard of rank 3 (B = 73 byte columns: the wide-record kernel) and checks
    size-independent properties; the shards' counts add up to the whole (test_shard.py)."""
    import torch

    import aces_b200
    from aces_b200 import shard

    fd = aces_b200.synthetic_design("rotated", 17)
    est = aces_b200.DesignEstimator(ctx, fd)
    shots_set = [10**7, 10**8, 10**9]
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    assert B == 73
    rows, prefix = [], []
    for e in range(n_exp):
        t = int(est.tuple_of_experiment[e])
        lo, hi, local = shard.shard_shots(int(shots_max[t]), shots_divided[t], 8, 3)
        rows.append(hi - lo)
        prefix.append(local)
    prefix = np.asarray(prefix, dtype=np.int64)
    ld = [(r + 127) // 128 * 128 for r in rows]
    assert 8.5e9 < sum(r * B for r in rows) < 9.5e9
    g = torch.Generator(device="cuda").manual_seed(17)
    dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
    total = int(np.sum(est.experiment_sizes()) * 3)
    out = torch.zeros(total, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def run():
        est.odd_counts([d.data_ptr() for d in dev], rows, ld, prefix, out=out.data_ptr(), asynchronous=True)
        ctx.synchronize()
        return out.cpu().numpy().copy()

    got = run()
    off = np.concatenate([[0], np.cumsum(est.experiment_sizes() * 3)])
    rng = np.random.default_rng(5)
    for e in rng.choice(n_exp, size=4, replace=False):
        E = est.tables.experiment_size(int(e))
        sup = ob.table_supports(est, int(e))
        for p in rng.choice(E, size=4, replace=False):
            acc = torch.zeros(rows[e], dtype=torch.uint8, device="cuda")
            for q in sup[p]:
                acc ^= (dev[e][(q - 1) // 8, :rows[e]] >> ((q - 1) % 8)) & 1
            csum = torch.cumsum(acc.to(torch.int64), 0)
            for k in range(3):
                S = int(prefix[e, k])
                assert got[off[e] + k * E + p] == (int(csum[S - 1]) if S > 0 else 0)
    for d in dev:
        d.bitwise_not_()
    torch.cuda.synchronize()
    flipped = run()
    for e in range(n_exp):
        E = est.tables.experiment_size(e)
        w = np.array([len(s) for s in ob.table_supports(est, e)])
        a = got[off[e]:off[e + 1]].reshape(3, E)
        b = flipped[off[e]:off[e + 1]].reshape(3, E)
        S = prefix[e][:, None]
        assert np.array_equal(np.where(w % 2 == 1, S - a, a), b)


def test_shot_sharded_counts_add_up(ctx):
    """Two emulated ranks on one GPU: shot-range shards of every sample matrix reduce to counts that
    add up to the unsharded counts (the all_reduce(sum) of a multi-GPU run; exact in integers)."""
    import aces_b200
    from aces_b200 import shard

    fd = aces_b200.synthetic_design("rotated", 5, full_covariance=True)
    est = aces_b200.DesignEstimator(ctx, fd)
    shots_divided, shots_max = aces_b200.shots_divide(fd, [2_000, 33_333, 150_001])
    rng = np.random.default_rng(9)
    mats, prefix = [], []
    for e in range(fd.experiment_number):
        t = int(est.tuple_of_experiment[e])
        mats.append(np.asfortranarray(rng.integers(0, 256, size=(int(shots_max[t]), fd.n_bytes), dtype=np.uint8)))
        prefix.append(shots_divided[t])
    prefix = np.asarray(prefix, dtype=np.int64)
    whole = est.odd_counts(mats, [m.shape[0] for m in mats], [m.shape[0] for m in mats], prefix)
    total = np.zeros_like(whole)
    for rank in range(2):
        sub, rows, pre = [], [], []
        for e, m in enumerate(mats):
            lo, hi, local = shard.shard_shots(m.shape[0], prefix[e], 2, rank)
            sub.append(np.asfortranarray(m[lo:hi]))
            rows.append(hi - lo)
            pre.append(local)
        total += est.odd_counts(sub, rows, [max(r, 1) for r in rows], np.asarray(pre, dtype=np.int64))
    assert np.array_equal(total, whole)


def test_column_groups_bookkeeping(ctx):
    """Narrow experiments are one group over all byte columns; wide records (d = 17: 73 columns) are
    split into groups that together read every column at least once."""
    import aces_b200

    for d, single in ((9, True), (17, False)):
        fd = aces_b200.synthetic_design("rotated", d)
        est = aces_b200.DesignEstimator(ctx, fd)
        cols = [est.tables.columns_read(e) for e in range(fd.experiment_number)]
        if single:
            assert all(c == fd.n_bytes for c in cols)
        else:
            assert all(fd.n_bytes <= c <= 1.25 * fd.n_bytes for c in cols)


def test_wide_record_with_ungroupable_pauli_uses_fallback_kernel(ctx):
    """A Pauli whose support spans more than 25 byte columns of a record wider than 41 bytes cannot
    be put into a column group: the tables stay ungrouped and the CTA-synchronous kernel runs."""
    import aces_b200

    n, rows = 577, 20_011
    rng = np.random.default_rng(77)
    B = (n + 7) // 8
    counts = np.asfortranarray(rng.integers(0, 256, size=(rows, B), dtype=np.uint8))
    sup = rand_supports(rng, n, 60)
    sup[7] = sorted(int(8 * c + rng.integers(0, 8)) + 1 for c in range(0, 70, 2))  # 35 columns
    sup[7] = [q for q in sup[7] if q <= n]
    signs = rng.integers(0, 2, size=len(sup))
    for S in (1, rows // 2, rows):
        got = aces_b200.estimate_experiment(ctx, counts, sup, signs, S)
        want, _ = ob.estimate_experiment(counts, sup, signs, S)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,rows,pad", [(17, 1, 0), (17, 511, 0), (49, 513, 0), (49, 70001, 3), (161, 33333, 0), (161, 4096, 11),
                                        (169, 20000, 0), (577, 5000, 0)])
def test_rowmajor_ingest_vs_oracle(ctx, n, rows, pad):
    """Stim's row-major bit-packed records (shots x bytes, C order) taken as they are --
    bit-exact against the oracle on the transposed (column-major) input; a row pitch larger than the
    record and device-resident records as well."""
    import torch

    import aces_b200

    rng = np.random.default_rng(n + rows)
    B = (n + 7) // 8
    rec = rng.integers(0, 256, size=(rows, B + pad), dtype=np.uint8)       # C order, pitch B + pad
    sup = rand_supports(rng, n, 120)
    ptr, flat = ob._csr(sup)
    tables = aces_b200.PauliTables(ctx, n, [0, 120], ptr, flat - 1)
    prefixes = sorted({1, max(1, rows // 3), rows})
    K = len(prefixes)
    want = np.zeros((K, 120), dtype=np.int64)
    colmajor = np.asfortranarray(rec[:, :B])
    for k, S in enumerate(prefixes):
        _, ps = ob.estimate_experiment(colmajor, sup, [0] * 120, S)
        want[k] = (S - ps) // 2
    got = aces_b200._lib.parity_counts(ctx, tables, [0], [rec], [rows], [B + pad], [prefixes], rowmajor=True)
    assert np.array_equal(got.reshape(K, 120), want)
    dev = torch.from_numpy(rec).cuda()
    got_dev = aces_b200._lib.parity_counts(ctx, tables, [0], [dev.data_ptr()], [rows], [B + pad], [prefixes], rowmajor=True)
    assert np.array_equal(got_dev.reshape(K, 120), want)
    # accumulating two row-major batches == the concatenated records (src/simulate.jl:174-186)
    if rows > 2:
        cut = rows // 2
        acc = aces_b200._lib.parity_counts(ctx, tables, [0], [rec[:cut]], [cut], [B + pad], [[cut]], rowmajor=True)
        acc = aces_b200._lib.parity_counts(ctx, tables, [0], [rec[cut:]], [rows - cut], [B + pad], [[rows - cut]], out=acc,
                                           accumulate=True, rowmajor=True)
        assert np.array_equal(acc, want[-1])


def test_rowmajor_design_batch_equals_colmajor(ctx):
    """A whole repetition of the real rotated d=9 design through the row-major entry point (the ingest
    path of simulate_estimate) equals the column-major path on the same records."""
    import aces_b200

    fd = aces_b200.rotated_design(9, full_covariance=True)
    est = aces_b200.DesignEstimator(ctx, fd)
    rng = np.random.default_rng(2)
    shots_set = [20_000, 300_000, 1_000_003]
    _, shots_max = aces_b200.shots_divide(fd, shots_set)
    recs = [rng.integers(0, 256, size=(int(shots_max[int(est.tuple_of_experiment[e])]), fd.n_bytes), dtype=np.uint8)
            for e in range(fd.experiment_number)]
    eig_r, cov_r = est.simulate_estimate(recs, shots_set)                           # C order -> row-major ingest
    eig_c, cov_c = est.simulate_estimate([np.asfortranarray(r) for r in recs], shots_set)
    assert eig_r == eig_c and cov_r == cov_c


@pytest.mark.parametrize("source", ["synthetic3", "real5"])
def test_estimate_ensemble_thousands_of_small_circuits(ctx, source):
    """estimate_stim_ensemble (src/device.jl:156-244): a randomised design runs thousands of small circuits
    (64 ... 10^4 shots, here 64 ... 1200 with odd counts) with per-randomisation Pauli signs; every job goes
    through ONE launch (n_items = its circuits).  Each circuit is checked against two estimate_experiment
    calls of the oracle, as the reference makes them, and the ensemble nesting against the reference's."""
    import aces_b200

    if source == "synthetic3":
        fd = aces_b200.synthetic_design("rotated", 3, full_covariance=True, stress_weights=True)
    else:
        fd = aces_b200.rotated_design(5, full_covariance=True)
    est = aces_b200.DesignEstimator(ctx, fd)
    rng = np.random.default_rng(23)
    S = 257
    randomisations = 12 if source == "synthetic3" else 3
    has_cov = bool(fd.cov_keys)
    # signs per (tuple, experiment, randomisation)
    sign_ens, cov_sign_ens, circuits = [], [], []
    for i in range(fd.tuple_number):
        sign_ens.append([])
        cov_sign_ens.append([])
        for j in range(int(fd.experiment_numbers[i])):
            ne = len(fd.experiment_ensemble[i][j])
            nc = len(fd.cov_experiment_ensemble[i][j]) if has_cov and len(fd.cov_keys[i]) else 0
            sign_ens[i].append([rng.integers(0, 2, size=ne).astype(bool) for _ in range(randomisations)])
            cov_sign_ens[i].append([rng.integers(0, 2, size=nc).astype(bool) for _ in range(randomisations)])
            circuits += [(i, j, r) for r in range(randomisations)]
    order = rng.permutation(len(circuits))
    circuits = [circuits[k] for k in order]
    n_jobs = 3
    job_indices = [circuits[b::n_jobs] for b in range(n_jobs)]
    assert sum(len(x) for x in job_indices) >= (1000 if source == "synthetic3" else 250)
    jobs = []
    for b, idx in enumerate(job_indices):
        mats = []
        for c in range(len(idx)):
            rows = S + int(rng.integers(0, 3)) * 64 if c % 7 else S        # some circuits carry spare rows
            mats.append(np.asfortranarray(rng.integers(0, 256, size=(rows, fd.n_bytes), dtype=np.uint8)))
        jobs.append(mats)
    launches0 = ctx.kernel_launches()
    eig_ens, cov_ens = est.estimate_ensemble(jobs, job_indices, sign_ens, cov_sign_ens, S)
    assert ctx.kernel_launches() - launches0 <= 3 * n_jobs     # parity kernel + prefix kernel (+ table upload) per job
    # the reference loop with the oracle's estimate_experiment
    exp0 = np.concatenate([[0], np.cumsum(fd.experiment_numbers)]).astype(int)
    want_eig = [[[] for _ in range(int(m))] for m in fd.mapping_lengths]
    want_cov = [[[] for _ in range(len(fd.cov_keys[t]) if has_cov else 0)] for t in range(fd.tuple_number)]
    for mats, idx in zip(jobs, job_indices):
        for m, (i, j, r) in zip(mats, idx):
            sup = ob.table_supports(est, exp0[i] + j)
            ne = len(fd.experiment_ensemble[i][j])
            e1, _ = ob.estimate_experiment(m, sup[:ne], sign_ens[i][j][r], S)
            for k, p in enumerate(fd.experiment_ensemble[i][j]):
                want_eig[i][int(p)].append(float(e1[k]))
            if has_cov and len(fd.cov_keys[i]):
                e2, _ = ob.estimate_experiment(m, sup[ne:], cov_sign_ens[i][j][r], S)
                for k, p in enumerate(fd.cov_experiment_ensemble[i][j]):
                    want_cov[i][int(p)].append(float(e2[k]))
    assert eig_ens == want_eig          # bit-exact, same nesting and order
    assert cov_ens == want_cov
