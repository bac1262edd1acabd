"""CPU tests: the C oracle against the reference's known-answer tests and the golden fixture."""
import numpy as np
import pytest

import oracle_bindings as ob
from golden_util import parity_cases


def pack_records(bits):
    """bits [shots, n] 0/1 -> uint8 [shots, ceil(n/8)] with the reference's addressing: qubit m
    (1-based) is bit (m-1) % 8 of byte fld(m-1, 8) (src/estimate.jl:249, src/utils.jl:477-503)."""
    shots, n = bits.shape
    B = (n + 7) // 8
    out = np.zeros((shots, B), dtype=np.uint8)
    for m in range(1, n + 1):
        out[:, (m - 1) // 8] |= (bits[:, m - 1].astype(np.uint8) << ((m - 1) % 8)).astype(np.uint8)
    return out


def test_golden_fixture_matches_oracle():
    for case in parity_cases():
        for k, S in enumerate(case["prefixes"]):
            est, ps = ob.estimate_experiment(case["counts"], case["supports"], case["signs"], int(S))
            assert np.array_equal(ps, case["pauli_shots"][k])
            assert np.array_equal(est, case["est"][k])  # exact: same IEEE division


def test_bitstring_round_trip_addressing():
    # mirrors test/device_tests.jl:45-54 (101 bits in 13 bytes, bit_num % 8 != 0)
    rng = np.random.default_rng(5)
    for n in (17, 49, 101, 161, 577):
        bits = rng.integers(0, 2, size=(64, n))
        rec = pack_records(bits)
        # garbage in the unused high bits of the last byte must not matter
        pad = (8 - n % 8) % 8
        if pad:
            rec[:, -1] |= (rng.integers(0, 256, size=64).astype(np.uint8) << (8 - pad)).astype(np.uint8)
        for m in rng.choice(n, size=8, replace=False) + 1:
            ps = ob.pauli_estimate(rec, [int(m)], 64)
            assert ps == int(np.sum(1 - 2 * bits[:, m - 1]))


def test_noiseless_known_answer():
    # test/device_tests.jl:144-155: noiseless records => every circuit eigenvalue == 1.0 exactly
    n, S = 49, 777
    rec = np.zeros((S, 7), dtype=np.uint8)
    sup = [[1], [2, 3], [49], [5, 6, 7, 8], []]
    est, ps = ob.estimate_experiment(rec, sup, [0] * 5, S)
    assert np.all(ps == S) and np.all(est == 1.0)
    est, _ = ob.estimate_experiment(rec, sup, [1, 0, 1, 0, 1], S)
    assert np.array_equal(est, np.array([-1.0, 1.0, -1.0, 1.0, -1.0]))


def test_prefix_uses_first_rows_only():
    rng = np.random.default_rng(7)
    rec = rng.integers(0, 256, size=(1000, 3), dtype=np.uint8)
    sup = [[1], [2, 9], [17, 3, 12]]
    est, odd = ob.simulate_estimate(rec, sup, [0, 1, 0], [13, 500, 1000])
    for k, S in enumerate([13, 500, 1000]):
        e2, ps2 = ob.estimate_experiment(rec[:S].copy(), sup, [0, 1, 0], S)
        assert np.array_equal(est[k], e2)
        assert np.array_equal(odd[k], (S - ps2) // 2)


def test_oracle_vs_numpy_bruteforce():
    rng = np.random.default_rng(11)
    n, S = 161, 4099
    bits = rng.integers(0, 2, size=(S, n))
    rec = pack_records(bits)
    sup = [sorted(rng.choice(n, size=w, replace=False) + 1) for w in (1, 2, 2, 3, 6, 8)]
    _, ps = ob.estimate_experiment(rec, sup, [0] * len(sup), S)
    for i, s in enumerate(sup):
        par = np.sum(bits[:, np.asarray(s) - 1], axis=1) % 2
        assert ps[i] == int(np.sum(1 - 2 * par))


@pytest.mark.parametrize("shots,meas,maxs", [(10**6, 17, 10**5), (1041667, 161, 10**8), (255, 3, 100),
                                              (256, 5, 256), (1, 1, 1), (10**7, 49, 10**9 // 7)])
def test_batch_shots_invariants(shots, meas, maxs):
    # src/utils.jl:25-27: all batches positive, sum == shots; all but at most one are multiples of 256
    b = ob.batch_shots(shots, meas, maxs)
    assert np.all(b > 0) and int(b.sum()) == shots
    assert int(np.sum(b % 256 != 0)) <= 1
