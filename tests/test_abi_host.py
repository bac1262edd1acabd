"""The C ABI driven from plain C (tests/abi_host.c: gcc + dlopen), the way the Julia shim drives it:
1-based index arrays, pageable buffers, the shim's call sequences -- a stand-in for running
ACESB200.jl, which needs a julia binary this image does not have."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi_host.c")
LIB = os.path.join(ROOT, "quantumaces.jl_b200", "libaces_b200.so")


def build_host(tmp_path):
    exe = str(tmp_path / "abi_host")
    subprocess.run(["gcc", "-O1", "-Wall", "-o", exe, SRC, "-ldl", "-lm"], check=True)
    return exe


def test_c_host_resolves_every_symbol(built, tmp_path):
    exe = build_host(tmp_path)
    out = subprocess.run([exe, LIB, "--symbols"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "ALL SYMBOLS RESOLVED" in out.stdout


def write_fixture(path, arrays):
    kinds = {np.dtype(np.uint8): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2, np.dtype(np.float64): 3}
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            f.write(name.encode().ljust(32, b"\0"))
            f.write(np.int32(kinds[a.dtype]).tobytes())
            f.write(np.int64(a.size).tobytes())
            f.write(a.tobytes())


@pytest.mark.gpu
def test_c_host_runs_the_shim_sequences(ctx, tmp_path):
    import scipy.sparse as sp

    import aces_b200
    import oracle_bindings as ob
    from fit_util import fo, noisy_ensembles

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import project_oracle as po

    fd = aces_b200.rotated_design(3, full_covariance=True, noise="lognormal", seed=4)
    est = aces_b200.DesignEstimator(ctx, fd)
    rng = np.random.default_rng(12)
    shots_set = [3_000, 20_011, 90_007]
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    recs, rows, prefix = [], [], []
    for e in range(n_exp):
        t = int(est.tuple_of_experiment[e])
        recs.append(rng.integers(0, 256, size=(int(shots_max[t]), B), dtype=np.uint8))
        rows.append(int(shots_max[t]))
        prefix.append(shots_divided[t])
    prefix = np.asarray(prefix, dtype=np.int64)
    odd = ob.oracle_odd_counts(est, [np.asfortranarray(r) for r in recs], prefix)
    rec_off = np.concatenate([[0], np.cumsum([r.size for r in recs])]).astype(np.int64)
    exp_ptr, pauli_ptr, pauli_bits, _ = aces_b200.build_pauli_tables(fd)
    # S1 on experiment 0 with the reference's 1-based qubit indices
    exp0 = fd.experiment_ensemble[0][0]
    sup = [[int(q) for q in fd.meas_indices[0][int(p)]] for p in exp0]
    s1_ptr, s1_flat = ob._csr(sup)
    s1_signs = np.asarray(fd.pauli_signs[0])[exp0].astype(np.uint8)
    s1_est, _ = ob.estimate_experiment(np.asfortranarray(recs[0]), sup, s1_signs, int(prefix[0, -1]))
    # fit inputs and oracle results
    eig0, cov = noisy_ensembles(fd, seed=3)
    multi = np.nonzero(np.diff(sp.csr_matrix(fd.matrix).indptr) >= 2)[0]
    off = np.concatenate([[0], np.cumsum(fd.mapping_lengths)])
    while True:   # clip 11 rows (circuit eigenvalue 0.05 < 0.1) such that the design keeps full column rank
        clip = rng.choice(multi, size=11, replace=False)
        keep_mask = np.ones(fd.M, dtype=bool)
        keep_mask[clip] = False
        if np.linalg.matrix_rank(sp.csr_matrix(fd.matrix)[keep_mask].toarray().astype(np.float64)) == fd.matrix.shape[1]:
            break
    eig = [[list(v) for v in t] for t in eig0]
    for r in clip:
        t = int(np.searchsorted(off, r, side="right") - 1)
        eig[t][int(r - off[t])] = [0.05 for _ in eig[t][int(r - off[t])]]
    core = fo.estimate_gate_noise_core(fd, eig, cov)
    lam = core["est_eigenvalues"]
    lam_cov = np.concatenate(fo.mean_ensemble(cov))
    kept = core["clipped_indices"]
    A = sp.csc_matrix(fd.matrix)
    A.sort_indices()
    Ac = sp.csc_matrix(sp.csr_matrix(fd.matrix)[kept, :]).astype(np.float64)
    Ac.sort_indices()
    dd = aces_b200.DeviceDesign(ctx, fd)
    cov_log = fo.calc_covariance_log(fo.calc_covariance_from_experiments(fd, eig, cov, weight_time=False), lam)
    cml = fo.get_clipped_mapping_lengths(fd.mapping_lengths, kept)
    P = fo.calc_precision_matrix(Ac, core["gls"], fo.sparse_covariance_inv(core["covariance_log"], cml)).toarray()
    blocks = np.concatenate([P[a:b, a:b].ravel() for a, b in zip(fd.gate_ptr[:-1], fd.gate_ptr[1:])])
    g_true = fd.gate_eigenvalues
    true_cl = fo.calc_covariance_log(fo.calc_covariance(fd, fo.get_eigenvalues_ensemble(fd, g_true),
                                                        fo.calc_covariance_eigenvalues(fd, g_true)), fo.get_eigenvalues(fd, g_true))
    keys = np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1, 2) for c in fd.cov_keys])
    cov_ptr = np.concatenate([[0], np.cumsum([len(c) for c in fd.cov_keys])]).astype(np.int64)
    arrays = dict(
        n_qubits=np.array([fd.n_qubits], np.int32), n_exp=np.array([n_exp], np.int32), K=np.array([3], np.int32),
        exp_ptr=exp_ptr.astype(np.int64), pauli_ptr=pauli_ptr.astype(np.int64), pauli_bits=pauli_bits.astype(np.int32),
        rows=np.asarray(rows, np.int64), rec_off=rec_off, records=np.concatenate([r.ravel() for r in recs]),
        prefix=prefix.ravel(), odd=odd.astype(np.int64),
        s1_ptr=s1_ptr.astype(np.int64), s1_qubits=s1_flat.astype(np.int32), s1_signs=s1_signs, s1_est=s1_est.astype(np.float64),
        T=np.array([fd.tuple_number], np.int32), N=np.array([A.shape[1]], np.int64), M=np.array([A.shape[0]], np.int64),
        mapping_lengths=np.asarray(fd.mapping_lengths, np.int64),
        A_colptr=(A.indptr + 1).astype(np.int32), A_rowval=(A.indices + 1).astype(np.int32), A_nzval=A.data.astype(np.int32),
        experiment_numbers=np.asarray(fd.experiment_numbers, np.int64), shot_weights=np.asarray(fd.shot_weights, np.float64),
        tuple_times=np.asarray(fd.tuple_times, np.float64), diag_counts=np.concatenate(fd.diag_counts).astype(np.int64),
        cov_ptr=cov_ptr, cov_p=(keys[:, 0] + 1).astype(np.int64), cov_q=(keys[:, 1] + 1).astype(np.int64),
        cov_counts=np.concatenate(fd.cov_counts).astype(np.int64),
        lam=lam, lam_cov=lam_cov, kept_rows=(kept + 1).astype(np.int64), lam_kept=lam[kept],
        cov_log_nzval=dd.pattern_values(cov_log), true_cov_log_nzval=dd.pattern_values(true_cl),
        ols=core["ols"], wls=core["wls"], gls=core["gls"],
        gate_ptr=np.asarray(fd.gate_ptr, np.int64), gate_idx1=np.arange(1, A.shape[1] + 1, dtype=np.int32),
        precision_blocks=blocks, gls_simple_proj=po.simple_project_gate_eigenvalues(core["gls"], fd.gate_ptr)[0],
        Ac_colptr=(Ac.indptr + 1).astype(np.int64), Ac_rowval=(Ac.indices + 1).astype(np.int64), Ac_nzval=Ac.data,
        gate_eigenvalues=g_true, tr_gls=np.array([np.trace(fo.calc_gls_covariance(fd, true_cl))]),
    )
    dd.close()
    fixture = str(tmp_path / "fixture.bin")
    write_fixture(fixture, arrays)
    exe = build_host(tmp_path)
    out = subprocess.run([exe, LIB, fixture], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASS" in out.stdout and "FAIL" not in out.stdout
