#!/usr/bin/env python
"""bench.py -- ACES estimation hot path on B200: shot-parities/s (K1) at rotated d=9, 1e8 shots.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the K1 parity kernel over one repetition's sample matrices (BASELINE.json
configs[2]: rotated surface code d=9, WLS design, 1e8 shots split over 96 experiments, three
nested budgets 1e6/1e7/1e8): 2.1 GB of bit-packed shot records -> 96 x 241 x 3 odd-parity
counts.  Multi-GPU: one repetition per rank per step (repetitions are independent,
src/simulate.jl:371), the per-repetition count vectors all-gathered with NCCL ("weak").

The JSON line also carries a `fit` object: GLS / WLS fits per second at the rotated d=9 shape
through aces_design_fit (host vectors in, host gate eigenvalues out), the Gram kernel's FP64
rate against the cuBLAS DGEMM rate measured in the same run, and the CPU restatement's time for
one fit.  Multi-GPU: fits are independent per budget and repetition (src/simulate.jl:393), so
every rank runs its own (replicas).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over
Paulis like the reference's @threads :static) -- Julia is not available in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "shot_parities_per_s"
UNIT = "shot-parities/s"
SHOTS_SET = [10**6, 10**7, 10**8]
WORKLOAD = "rotated_d9_wls_1e8shots"

# BASELINE.json configs as workloads (--workload).  The default (the configuration the metric is quoted on)
# is configs[2]; the others emit the same JSON line for their shape.
#   sharding "repetitions": one repetition per rank per step (weak); "shots": every rank holds a contiguous
#   shot range of every experiment and the int64 counts are all-reduced (config 5: 73 GB over 8 GPUs).
WORKLOADS = {
    "rotated_d9_wls_1e8shots": dict(kind="rotated", d=9, full=False, shots=[10**6, 10**7, 10**8], fit_full=True, sharding="repetitions",
                                    baseline="configs[2]: rotated surface code d=9, full_covariance=false WLS, 1e8 shots"),
    "rotated_d5_gls": dict(kind="rotated", d=5, full=True, shots=[10**6, 10**7, 10**8], fit_full=True, sharding="repetitions",
                           baseline="configs[1]: rotated surface code d=5, log-normal Pauli noise, full_covariance=true GLS"),
    "rotated_d9_gls": dict(kind="rotated", d=9, full=True, shots=[10**6, 10**7, 10**8], fit_full=True, sharding="repetitions",
                           baseline="configs[2]'s circuit with full_covariance=true: the covariance Paulis are reduced in the same pass (449 Paulis per experiment)"),
    "unrotated_d7_gls": dict(kind="unrotated", d=7, full=True, shots=[10**6, 10**7, 10**8], fit_full=True, sharding="repetitions",
                             baseline="configs[3]: unrotated surface code d=7, log-normal noise, GLS estimate + calc_covariance merit"),
    "rotated_d9_wls_experiments": dict(kind="rotated", d=9, full=False, shots=[10**6, 10**7, 10**8], fit_full=False, sharding="experiments",
                                       baseline="configs[2] with ONE repetition split over the GPUs by experiment (SURVEY 8e, second axis): "
                                                "strong scaling, int64 all-reduce of the counts"),
    "rotated_d17_1e9": dict(kind="rotated", d=17, full=False, shots=[10**7, 10**8, 10**9], fit_full=False, sharding="shots",
                            baseline="configs[4]: rotated surface code d=17, 1e9-shot synthetic budget, shots sharded over 8 GPUs"),
}
WL = WORKLOADS[WORKLOAD]


FIT_CONTEXTS = 4         # concurrent fit contexts of the throughput line (--fit-contexts)
DESIGN_SOURCE = "real"   # "real": circuit_design.rotated_design(9) (the reference's own tables); "synthetic": look-alike


def workload_config(n_gpus, fd=None, sizes=None, bytes_per_gpu=None):
    n_exp = int(fd.experiment_number) if fd is not None else 96
    mean_paulis = float(sum(int(x) for x in sizes)) / n_exp if sizes is not None else None
    shots_sharded = WL["sharding"] == "shots"
    return {
        "workload": WORKLOAD,
        "baseline_config": WL["baseline"],
        "design": (f"generate_design of the {WL['kind']} planar d={WL['d']} circuit restated in numpy (inputs/circuit_design.py): "
                   "real Pauli tables, experiment packing and shot weights" if DESIGN_SOURCE == "real" else "synthetic look-alike tables"),
        "n_qubits": int(fd.n_qubits) if fd is not None else 161, "bytes_per_shot": int(fd.n_bytes) if fd is not None else 21,
        "experiments": n_exp, "paulis_per_experiment_mean": mean_paulis,
        "budgets": [int(x) for x in shots_for(n_gpus)],
        "repetitions_per_step": 1 if (shots_sharded or WL["sharding"] == "experiments") else n_gpus,
        "l2": (f"inputs ({bytes_per_gpu / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush needed" if bytes_per_gpu and bytes_per_gpu > 2.5e8
               else "inputs fit the 126 MB L2: a 256 MB buffer is overwritten between timed steps"),
        "parallelism": ((f"shot ranges of every experiment x{n_gpus}, int64 all-reduce of the counts (the full 1e9-shot budget "
                         "needs 8 GPUs: per-GPU work is fixed, 1/8 of the budget per rank)") if shots_sharded
                        else (f"the experiments of one repetition dealt to {n_gpus} GPUs by bytes (greedy), int64 all-reduce of the counts"
                              if WL["sharding"] == "experiments" and n_gpus > 1
                              else (f"repetitions x{n_gpus}" if n_gpus > 1 else "single GPU"))),
    }


def shots_for(n_gpus):
    """Budgets of the workload: the shot-sharded config keeps 1/8 of the 1e9-shot budget per rank (weak)."""
    if WL["sharding"] == "shots":
        return [s * n_gpus // 8 for s in WL["shots"]]
    return list(WL["shots"])


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (NVML in-process: forking
    nvidia-smi from a CUDA process stalls the launching thread; nvidia-smi is the fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples = []  # (sm_mhz, sm_max_mhz, set of reasons)
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            masks = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            self.samples.append((sm, self.sm_max, {k for k, m in masks.items() if bits & m}))
            return
        out = subprocess.run(
            ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
            capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.samples.append((float(f[0]), float(f[1]),
                                 {n for n, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")}))

    def run(self):
        while not self.stop.is_set():
            try:
                self.sample()
            except Exception:
                pass
            self.stop.wait(self.period if self.nvml is not None else 0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[1] for s in self.samples), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_design():
    import aces_b200

    if DESIGN_SOURCE == "real":
        gen = aces_b200.rotated_design if WL["kind"] == "rotated" else aces_b200.unrotated_design
        return gen(WL["d"], full_covariance=WL["full"])
    return aces_b200.synthetic_design(WL["kind"], WL["d"], full_covariance=WL["full"])


def cpu_sample(fd, est_sizes, seconds_target=12.0):
    """Times the CPU restatement of the reference on a bounded sample of the workload: one
    experiment at a time (its measured Paulis, 3 budget passes as in src/simulate.jl:206-213) over `rows` shots."""
    import numpy as np
    import oracle_bindings as ob

    import aces_b200

    rng = np.random.default_rng(0)
    ob.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the reference uses every core
    threads = ob.num_threads()
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_for(1))
    rows = int(shots_max.max())
    m = np.asfortranarray(rng.integers(0, 256, size=(rows, fd.n_bytes), dtype=np.uint8))
    ob.estimate_experiment(m[:1000], [[1]], [0], 1000)  # start the OpenMP team outside the timing
    done, units, dt = 0, 0, 0.0
    for t in range(fd.tuple_number):
        for j in range(int(fd.experiment_numbers[t])):
            exp = fd.experiment_ensemble[t][j]
            sup = [[int(q) for q in fd.meas_indices[t][int(p)]] for p in exp]
            signs = np.asarray(fd.pauli_signs[t])[exp].astype(np.uint8)
            S = int(shots_max[t])
            t0 = time.perf_counter()
            ob.simulate_estimate(m[:S], sup, signs, [int(s) for s in shots_divided[t]])
            dt += time.perf_counter() - t0
            done += 1
            units += len(sup) * S  # shot-parities at the largest budget (same unit as the GPU metric)
            if dt >= seconds_target:
                break
        if dt >= seconds_target:
            break
    value = units / dt
    return {
        "value": value, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{done} of {fd.experiment_number} experiments of one repetition ({units} shot-parities, "
                  f"3 nested budgets = one pass per budget as src/simulate.jl:206-213), {dt:.2f} s, "
                  f"OpenMP static over Paulis",
        "note": "CPU restatement of the reference algorithm (oracle/parity_oracle.c), not Julia: "
                "no julia binary in this image",
    }


def dgemm_peak_tflops(torch, n=8192, reps=5):
    """cuBLAS DGEMM n^3 on this GPU (the FP64 roofline denominator)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def fit_inputs():
    """Rotated d=9 design with a full covariance pattern and estimates whose covariance blocks are
    positive definite (the common case; the not-PD fallback is covered by the tests)."""
    import numpy as np

    import aces_b200

    if DESIGN_SOURCE == "real":
        gen = aces_b200.rotated_design if WL["kind"] == "rotated" else aces_b200.unrotated_design
        fd = gen(WL["d"], full_covariance=True, noise="lognormal", seed=9)
    else:
        fd = aces_b200.synthetic_design(WL["kind"], WL["d"], full_covariance=True, with_matrix=True)
    eig, cov = aces_b200.synthetic_estimates(fd, shots=10**8, seed=11)
    lam = np.concatenate(aces_b200.mean_ensemble(eig))
    lc = np.concatenate(aces_b200.mean_ensemble(cov))
    return fd, eig, cov, lam, lc


def fit_leg(ctx, torch, dist, world, rank, steps, warmup, with_cpu):
    """GLS / WLS fits per second at the rotated d=9 shape, through the C ABI with host buffers."""
    import numpy as np

    import aces_b200

    fd, eig, cov, lam, lc = fit_inputs()
    dd = aces_b200.DeviceDesign(ctx, fd)
    out = {"workload": f"{WL['kind']}_d{WL['d']} GLS (full covariance pattern)", "design": DESIGN_SOURCE, "M": int(fd.M), "N": int(fd.matrix.shape[1]),
           "tuples": int(fd.tuple_number), "covariance_keys": int(dd.C),
           "api": "aces_design_fit: host estimate vectors in, host gate eigenvalues out (H2D + D2H inside)"}
    res = {}
    for name in ("gls", "wls"):
        for _ in range(warmup):
            g = dd.fit(name, lam, lc)
        assert dd.fallback_count() == 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            g = dd.fit(name, lam, lc)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = (world * steps / float(t.item()), g)
        out[f"{name}_fits_per_s"] = res[name][0]
        out[f"{name}_ms_per_fit"] = 1e3 * float(t.item()) / steps
    out["max_abs_error_vs_truth"] = float(np.max(np.abs(res["gls"][1] - fd.gate_eigenvalues)))
    # Throughput with several fit contexts on this GPU: fits are independent per budget and repetition
    # (src/simulate.jl:393) and one fit leaves most SMs idle (its Cholesky is a latency-bound chain),
    # so concurrent contexts (own stream and workspaces each, one host thread each) overlap.
    n_conc = FIT_CONTEXTS
    ctxs = [aces_b200.Context(ctx.device) for _ in range(n_conc - 1)]
    dds = [dd] + [aces_b200.DeviceDesign(c, fd) for c in ctxs]
    for name in ("gls", "wls"):
        outs = [None] * n_conc

        def work(i, n, name=name, outs=outs):
            for _ in range(n):
                outs[i] = dds[i].fit(name, lam, lc)

        for n in (warmup, steps):
            ths = [threading.Thread(target=work, args=(i, n)) for i in range(n_conc)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            [t.start() for t in ths]
            [t.join() for t in ths]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert all(np.array_equal(o, res[name][1]) for o in outs), "concurrent fits differ from the single-context fit"
        out[f"{name}_fits_per_s_concurrent{n_conc}"] = world * n_conc * steps / float(t.item())
    for x in dds[1:]:
        x.close()
    for c in ctxs:
        c.close()
    # Gram kernel of the GLS fit on the FP64 tensor pipe (lower tiles of At' At)
    ctx.set_timing(True)
    ks, phases = [], {}
    for name in ("wls", "gls", "gls", "gls"):
        dd.fit(name, lam, lc)
        phases[name] = dd.phase_ms()
        if name == "gls":
            ks.append(ctx.last_kernel_ms())
    ctx.set_timing(False)
    k_ms = sorted(ks)[1]
    out["phases_ms"] = phases  # CUDA events inside the library, one fit each
    n = fd.matrix.shape[1]
    n_pad = -(-n // 128) * 128
    flops = dd.gram_flops()
    peak = dgemm_peak_tflops(torch)
    potrf_ms = phases["gls"].get("potrf", float("nan"))
    chol_tf = n_pad ** 3 / 3.0 / (potrf_ms * 1e-3) / 1e12
    # The dominant kernel of a fit is the Cholesky factorisation of the Gram matrix: it is the roofline entry.
    out["roofline"] = {"bound": "tensor",
                       "kernel": ("dense::potrf_dag_kernel (persistent task-DAG Cholesky: DMMA m8n8k4 f64 update / panel tiles, "
                                  "128-block factor + inverse on the chain CTA)" if n_pad // 128 <= 80 else
                                  "dense::potrf (multi-stream blocked Cholesky, DMMA m8n8k4 f64 tiles)"),
                       "achieved": chol_tf, "peak": peak, "unit": "TFLOP/s", "frac": chol_tf / peak, "kernel_ms": potrf_ms,
                       "flops_per_launch": n_pad ** 3 / 3.0,
                       "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (burst, best of 5)",
                       "share_of_fit": {"gls": potrf_ms / out["gls_ms_per_fit"],
                                        "wls": phases["wls"].get("potrf", float("nan")) / out["wls_ms_per_fit"]},
                       "traffic": None,
                       "note": "latency-bound chain of 128 x 128 diagonal blocks, not a bandwidth- or pipe-bound kernel: "
                               "see profiles/fit_r02_potrf_variants.txt for the per-task trace"}
    out["gram"] = {"kernel": "dense::gemm_batched_kernel<TN> (Gram over the column windows of the tuple blocks, DMMA m8n8k4 f64)",
                   "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                   "frac": flops / (k_ms * 1e-3) / 1e12 / peak, "kernel_ms": k_ms, "executed_flops_per_launch": flops,
                   "share_of_fit": k_ms / out["gls_ms_per_fit"], "traffic": None}
    gpath = os.path.join(ROOT, "profiles", "fit_gram_traffic.json")
    if DESIGN_SOURCE == "real" and WORKLOAD == "rotated_d9_wls_1e8shots" and os.path.exists(gpath):
        try:  # one ncu --set full capture of this kernel (per launch), see profiles/fit_r01_gram_ncu_summary.txt
            gj = json.load(open(gpath))
            out["gram"]["traffic"] = gj.get("dram_bytes_per_launch")
            out["gram"]["ncu_tensor_pipe_dmma_active_pct"] = gj.get("tensor_pipe_dmma_active_pct")
        except Exception:
            pass
    # K6 (calc_gls / wls / ols_covariance + nrmse_moments, src/merit.jl:556-629, 672-682): BASELINE configs[3]'s
    # "calc_covariance merit", trace moments only (no N x N copy to the host)
    lam_ens, cov_ens = aces_b200.mean_ensemble(eig), aces_b200.mean_ensemble(cov)
    cl = dd.calc_covariance(lam_ens, cov_ens, weight_time=False, log_form=True)
    k6 = {}
    for name in ("gls", "wls", "ols"):
        dd.calc_ls_covariance(name, cl, want_matrix=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(steps // 2, 2)):
            _, tr, trsq = dd.calc_ls_covariance(name, cl, want_matrix=False)
        torch.cuda.synchronize()
        k6[name] = {"ms": 1e3 * (time.perf_counter() - t0) / max(steps // 2, 2),
                    "nrmse_expectation": float(aces_b200.nrmse_moments(tr, trsq, n)[0])}
    out["ls_covariance"] = dict(k6, api="aces_design_ls_covariance (trace moments; pattern values from the host)")
    # the whole numeric path of estimate_gate_noise with split projections (src/estimate.jl:703-831)
    if fd.gate_ptr is not None:
        eig_p, cov_p = aces_b200.synthetic_estimates(fd, shots=10**6, seed=5)   # few shots: projections are active
        aces_b200.estimate_gate_noise_projected(dd, eig_p, cov_p, clip_warning=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = aces_b200.estimate_gate_noise_projected(dd, eig_p, cov_p, clip_warning=False)
        torch.cuda.synchronize()
        out["estimate_gate_noise_projected"] = {
            "ms": 1e3 * (time.perf_counter() - t0), "precision_project": bool(r["precision_project"]),
            "gates": int(len(fd.gate_ptr) - 1),
            "what": "means, clipping, OLS + WLS + GLS fits, Euclidean projection of OLS, per-gate Mahalanobis projection of "
                    "WLS / GLS (precision slices from the device-resident Gram matrix, batched active-set QP), host vectors in / out"}
        # the simultaneous projection of all gates (split = false, the reference's default below 5000 gate eigenvalues)
        aces_b200.estimate_gate_noise_projected(dd, eig_p, cov_p, clip_warning=False, split=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rf = aces_b200.estimate_gate_noise_projected(dd, eig_p, cov_p, clip_warning=False, split=False)
        torch.cuda.synchronize()
        out["estimate_gate_noise_projected"]["split_false_ms"] = 1e3 * (time.perf_counter() - t0)
        out["estimate_gate_noise_projected"]["split_false_pivoting_steps_last"] = int(aces_b200.project.full_project_gate_eigenvalues.last_iterations)
        out["estimate_gate_noise_projected"]["split_false_max_diff_vs_split"] = float(np.max(np.abs(rf["gls"] - r["gls"])))
    # calc_merit (src/merit.jl:925-1006): three estimator covariances, marginal / relative
    # transforms and nine spectra in one call; and the eigenvalue solver alone on the GLS covariance
    if fd.gate_types is not None and fd.gate_ptr is not None:
        cl_true = dd.calc_covariance_log_of_gates(fd.gate_eigenvalues)
        tf = (aces_b200.get_transform(fd, "marginal"), aces_b200.get_transform(fd, "relative"))
        aces_b200.calc_merit(dd, covariance_log=cl_true, transforms=tf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mer = aces_b200.calc_merit(dd, covariance_log=cl_true, transforms=tf)
        torch.cuda.synchronize()
        merit_ms = 1e3 * (time.perf_counter() - t0)
        sigma, tr_s, _ = dd.calc_ls_covariance("gls", cl_true, want_matrix=True)
        aces_b200.eigvals(ctx, sigma)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = aces_b200.eigvals(ctx, sigma)
        torch.cuda.synchronize()
        eig_ms = 1e3 * (time.perf_counter() - t0)
        h2d_ms = sigma.nbytes / 50e9 * 1e3  # the N x N matrix crosses PCIe inside this call (~50 GB/s pageable-pinned mix)
        out["merit"] = {
            "api": "aces_design_merit: gate eigenvalues + covariance values in, nine spectra + Gram extremes out",
            "ms": merit_ms, "N": n, "N_marginal": int(mer["N_marginal"]), "N_relative": int(mer["N_relative"]),
            "gls_expectation": float(mer["gls_expectation"]), "wls_expectation": float(mer["wls_expectation"]),
            "ols_expectation": float(mer["ols_expectation"]), "cond_num": float(mer["cond_num"]),
            "trace_check_rel": float(abs(mer["gls_cov_eigenvalues"].sum() - tr_s) / tr_s),
            "eigvals": {"api": "aces_symmetric_eigenvalues (host N x N matrix in, N eigenvalues out)", "ms": eig_ms,
                        "kernel": "eigen::sytrd_panel_kernel (persistent, one launch per 64 reflectors) + DMMA trailing "
                                  "update + Sturm multisection",
                        "bound": "hbm", "algorithmic_bytes": 4.0 * n ** 3 / 3.0,
                        "achieved_gbs_incl_copy": 4.0 * n ** 3 / 3.0 / (eig_ms * 1e-3) / 1e9,
                        "note": "lower-triangle tiles of the trailing matrix read once per column (4 n^3 / 3 bytes); the call "
                                f"includes the H2D copy of the matrix (~{h2d_ms:.0f} ms)"}}
        if with_cpu and rank == 0:
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=os.cpu_count() or 1):
                t0 = time.perf_counter()
                ev_cpu = np.linalg.eigvalsh(sigma)
                out["merit"]["eigvals"]["cpu_lapack_ms"] = 1e3 * (time.perf_counter() - t0)
            out["merit"]["eigvals"]["cpu_cores"] = os.cpu_count()
            out["merit"]["eigvals"]["max_abs_diff_over_norm"] = float(np.max(np.abs(ev - ev_cpu)) / np.abs(ev_cpu).max())
            assert out["merit"]["eigvals"]["max_abs_diff_over_norm"] <= 1e-12, "eigenvalue parity lost"
    if with_cpu and rank == 0:
        # parity at the benchmarked shape: the GPU estimates of the SAME inputs against the CPU
        # restatement (north star: 1e-9 relative in fp64)
        base, cpu_vecs = cpu_fit_sample(fd, eig, cov, want_vectors=True)
        gpu_vecs = {"gls": res["gls"][1], "wls": res["wls"][1], "ols": dd.fit("ols", lam, lc)}
        out["parity_rel_err"] = {k: float(np.linalg.norm(gpu_vecs[k] - cpu_vecs[k]) / np.linalg.norm(cpu_vecs[k]))
                                 for k in ("ols", "wls", "gls")}
        out["parity_tolerance"] = 1e-9
        assert all(v <= 1e-9 for v in out["parity_rel_err"].values()), f"fit parity lost: {out['parity_rel_err']}"
        out["cpu_baseline"] = base
    dd.close()
    return out


def cpu_fit_sample(fd, eig, cov, want_vectors=False):
    """One GLS fit of the same inputs by the CPU restatement (numpy / LAPACK QR as the SPQR stand-in)."""
    import fit_util  # noqa: F401  (puts oracle/ on the path)
    import fit_oracle as fo
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=os.cpu_count() or 1):  # BLAS threads = every host core, stated in `cores`
        t0 = time.perf_counter()
        r = fo.estimate_gate_noise_core(fd, eig, cov)
        dt = time.perf_counter() - t0
    base = {"value": 1.0 / dt, "unit": "estimate_gate_noise cores/s (OLS + WLS + GLS fits of one estimate set)",
            "cores": os.cpu_count(), "kind": "port", "seconds": dt,
            "algorithm": "dense LAPACK QR (gelsy) stand-in for the reference's sparse SPQR solve: NOT the reference's "
                         "cost for OLS / WLS (3.5 nonzeros per row) -- a checker timing, not a speed-up denominator",
            "sample": "one estimate_gate_noise numeric core (covariance, block factor, OLS + WLS + GLS solves) at "
                      "rotated d=9", "gls_fits_per_s_equivalent": 3.0 / dt,
            "checksum": float(r["gls"].sum())}
    if want_vectors:
        return base, {k: r[k] for k in ("ols", "wls", "gls")}
    return base


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import aces_b200  # noqa: F401  (design generator only; no GPU context is created)

    fd = build_design()
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_sample(fd, None, seconds_target=4.0 if i >= args.warmup else 1.0)
        if i >= args.warmup:
            vals.append(base["value"])
    value = sum(vals) / len(vals)
    _, shots_max = aces_b200.shots_divide(fd, shots_for(args.gpus))
    sizes = [len(e) for t in range(fd.tuple_number) for e in fd.experiment_ensemble[t]]
    step_units = sum(len(e) * int(shots_max[t]) for t in range(fd.tuple_number) for e in fd.experiment_ensemble[t])
    base["value"] = value
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_units / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args.gpus, fd, sizes), "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_fit:
        f_fd, f_eig, f_cov, _, _ = fit_inputs()
        line["fit"] = {"workload": "rotated_d9 GLS (full covariance pattern)", "cpu_baseline": cpu_fit_sample(f_fd, f_eig, f_cov)}
    print(json.dumps(line))


def main():
    global DESIGN_SOURCE, FIT_CONTEXTS, WORKLOAD, WL
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--fit-steps", type=int, default=5)
    ap.add_argument("--design", default="real", choices=["real", "synthetic"])
    ap.add_argument("--fit-contexts", type=int, default=4)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="count exchange of the multi-GPU run: push / pull kernels over peer memory (NVLink), or NCCL")
    ap.add_argument("--workload", default="rotated_d9_wls_1e8shots", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    DESIGN_SOURCE = args.design
    FIT_CONTEXTS = args.fit_contexts
    WORKLOAD, WL = args.workload, WORKLOADS[args.workload]
    if not WL["fit_full"]:
        args.no_fit = True   # the d = 17 fit is a 25k x 25k factorisation per estimator: tools/fit_sharded.py times it
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # Libraries print banners on stdout (NCCL: "NCCL version ..."): stdout carries exactly one JSON
    # line, so everything else goes to stderr until the line is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import aces_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # bind the process to the CPUs next to its GPU before any pinned allocation (first touch places the
    # pages): the end-to-end leg is a host-to-device copy
    numa = "unchanged"
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        numa = f"nvmlDeviceSetCpuAffinity: {len(os.sched_getaffinity(0))} CPUs"
    except Exception as e:  # noqa: BLE001
        numa = f"not bound ({type(e).__name__})"

    fd = build_design()
    ctx = aces_b200.Context(local)
    est = aces_b200.DesignEstimator(ctx, fd)
    # A real (non-default) stream: handle 0 would select the library's own stream and the events
    # below would not bracket its kernels.
    stream = torch.cuda.Stream()
    comm = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    shots_set = shots_for(world)
    K = len(shots_set)
    shots_divided, shots_max = aces_b200.shots_divide(fd, shots_set)
    n_exp, B = fd.experiment_number, fd.n_bytes
    rows_full = [int(shots_max[int(est.tuple_of_experiment[e])]) for e in range(n_exp)]
    prefix_full = np.stack([shots_divided[int(est.tuple_of_experiment[e])] for e in range(n_exp)])
    shot_sharded = WL["sharding"] == "shots"
    if shot_sharded:
        # every rank: a contiguous shot range of every experiment, prefixes clamped into it (shard.shard_shots)
        parts = [aces_b200.shard.shard_shots(rows_full[e], prefix_full[e], world, rank) for e in range(n_exp)]
        rows = [int(hi - lo) for lo, hi, _ in parts]
        prefix = np.stack([p for _, _, p in parts])
    elif WL["sharding"] == "experiments" and world > 1:
        # one repetition over all ranks: every rank reduces the experiments dealt to it (shard.shard_experiments,
        # greedy by bytes); the others are empty items, so the count vector keeps the full layout and the
        # all-reduce (sum, int64) hands every rank the complete, exact counts
        mine = set(aces_b200.shard.shard_experiments([rows_full[e] * B for e in range(n_exp)], world)[rank])
        rows = [rows_full[e] if e in mine else 0 for e in range(n_exp)]
        prefix = np.stack([prefix_full[e] if e in mine else np.zeros_like(prefix_full[e]) for e in range(n_exp)])
    else:
        rows, prefix = rows_full, prefix_full
    exp_sharded = WL["sharding"] == "experiments" and world > 1
    reduce_counts = shot_sharded or exp_sharded
    ld = [max((r + 127) // 128 * 128, 128) for r in rows]
    sizes = est.experiment_sizes()
    g = torch.Generator(device="cuda").manual_seed(0x5EED0000 + rank)
    dev = [torch.randint(0, 256, (B, ld[e]), dtype=torch.uint8, device="cuda", generator=g) for e in range(n_exp)]
    ptrs = [d.data_ptr() for d in dev]
    n_counts = int(np.sum(sizes) * K)
    outs = [torch.zeros(n_counts, dtype=torch.int64, device="cuda") for _ in range(2)]   # double buffered: the collective
    gathered = [torch.zeros(world * n_counts, dtype=torch.int64, device="cuda") for _ in range(2)] if world > 1 and not reduce_counts else None
    alg_bytes = int(sum(rows[e] * B for e in range(n_exp)))          # read once, this rank
    shot_parities = int(sum(rows[e] * int(sizes[e]) for e in range(n_exp)))
    shots = int(sum(rows))
    small = alg_bytes < 2.5e8   # the inputs would stay in the 126 MB L2 between steps: flush it
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    # multi-GPU count exchange: kernels over peer memory (csrc/exchange.cu) unless NCCL is asked for or IPC fails
    xch, xch_note = None, None
    if world > 1 and args.exchange == "peer":
        try:
            xch = aces_b200.shard.PeerExchange(ctx, n_counts)
        except Exception as e:  # noqa: BLE001
            xch_note = f"peer exchange unavailable ({type(e).__name__}: {e}); NCCL used"
        ok = torch.tensor([1 if xch is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and xch is not None:
            xch.close()
            xch = None
    if xch is not None:
        xch.attach()
    reduced = [torch.zeros(n_counts, dtype=torch.int64, device="cuda") for _ in range(2)] if xch is not None and reduce_counts else None
    if xch is not None and gathered is None and not reduce_counts:
        gathered = [torch.zeros(world * n_counts, dtype=torch.int64, device="cuda") for _ in range(2)]
    batches = [est.prepare(ptrs, rows, ld, prefix, o.data_ptr(), asynchronous=True) for o in outs]
    done_ev = [torch.cuda.Event(), torch.cuda.Event()]      # kernel of buffer b finished
    free_ev = [torch.cuda.Event(), torch.cuda.Event()]      # collective of buffer b finished
    step_no = [0]

    def step():
        """One repetition (or one shot shard of it) on this rank: memset + parity kernel + prefix kernel on the
        compute stream; the NCCL exchange of its small count vector runs on a second stream, overlapped with
        the next step's kernel (two output buffers)."""
        b = step_no[0] & 1
        step_no[0] += 1
        if flush is not None:
            flush.fill_(1)
        if world > 1:
            stream.wait_event(free_ev[b])        # the collective that last read this buffer is done
        if world > 1 and xch is not None:
            stream.wait_event(free_ev[b])   # this rank's pull of step n - 2 (four slot sets rotate: see csrc/exchange.cuh)
        batches[b].launch()
        if world > 1 and xch is not None:
            # the push is fused into K1's prefix kernel (exchange attached to the context: the kernel that forms
            # the counts stores them into every rank's buffer over NVLink and releases the flags); the pull runs
            # on the second stream (waits for every rank's flag) and may finish behind the next step's kernel
            done_ev[b].record(stream)
            comm.wait_event(done_ev[b])
            xch.pull((reduced[b] if reduce_counts else gathered[b]).data_ptr(), reduce_counts, comm.cuda_stream)
            free_ev[b].record(comm)
        elif world > 1:
            done_ev[b].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(done_ev[b])
                if reduce_counts:
                    dist.all_reduce(outs[b], op=dist.ReduceOp.SUM)
                else:
                    dist.all_gather_into_tensor(gathered[b], outs[b])
                free_ev[b].record(comm)

    def fence():
        if world > 1:
            comm.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    fence()
    launches0 = ctx.kernel_launches()
    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        fence()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        if world > 1:
            stream.wait_stream(comm)   # the timed region ends when the last exchange has finished
        ev1.record(stream)
        fence()
        ms_total = ev0.elapsed_time(ev1)
        launches = ctx.kernel_launches() - launches0
        ctx.set_timing(True)
        # per-launch duration of the parity kernel itself (events inside the library, same stream)
        for _ in range(min(args.steps, 10)):
            step()
            kernel_ms.append(ctx.last_kernel_ms())
        fence()
    ctx.set_timing(False)
    flush_ms = 0.0
    if flush is not None:   # the L2 flush sits inside the timed region: time it alone and take it out
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe0.record(stream)
        for _ in range(args.steps):
            flush.fill_(1)
        fe1.record(stream)
        torch.cuda.synchronize()
        flush_ms = fe0.elapsed_time(fe1)
    t = torch.tensor([ms_total - flush_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    tot = torch.tensor([shot_parities, shots, alg_bytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    value = float(tot[0].item()) / (ms_per_step * 1e-3)
    out = outs[(step_no[0] - 1) & 1]
    if xch is not None:
        xch.status()      # raises if a pull timed out
        xch.attach(False)
        if reduce_counts:
            out = reduced[(step_no[0] - 1) & 1]
    counts_equal = None
    if reduce_counts and world > 1:   # after the all-reduce every rank holds the same counts: compare a checksum
        chk = torch.stack([out.sum(), (out * torch.arange(1, n_counts + 1, device="cuda")).sum()]).to(torch.int64)
        allchk = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        counts_equal = bool(all(torch.equal(c, allchk[0]) for c in allchk))
        assert counts_equal, "shot-sharded counts differ between ranks after the all-reduce"

    # ---- end to end through the public API with host buffers ---------------------------------
    host = [torch.empty((B, ld[e]), dtype=torch.uint8).pin_memory() for e in range(n_exp)]
    for e in range(n_exp):
        host[e].copy_(dev[e])
    torch.cuda.synchronize()
    hptrs = [h.data_ptr() for h in host]
    host_out = np.zeros(n_counts, dtype=np.int64)
    ref_out = torch.zeros(n_counts, dtype=torch.int64, device="cuda")
    est.prepare(ptrs, rows, ld, prefix, ref_out.data_ptr(), asynchronous=True).launch()   # this rank's own counts
    torch.cuda.synchronize()

    e2e_batch = est.prepare(hptrs, rows, ld, prefix, host_out)

    def e2e_step():
        # host matrices in, host counts out: H2D of every sample matrix and D2H of the counts
        e2e_batch.launch()

    e2e_step()
    ref_counts = ref_out.cpu().numpy()
    assert np.array_equal(host_out, ref_counts), "e2e counts differ from the device-resident run"

    def timed(fn, n):
        fence()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        fence()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    e2e_s = timed(e2e_step, args.e2e_steps)
    e2e_value = float(tot[0].item()) * args.e2e_steps / e2e_s
    # what bounds it: the same bytes as plain pinned host-to-device copies (no kernel), all ranks at once
    def plain_copies():
        for e in range(n_exp):
            dev[e].copy_(host[e], non_blocking=True)

    plain_s = timed(plain_copies, args.e2e_steps)
    # ingest path: the same records ROW-major, as Stim's sampler returns them
    rm_host = [torch.empty((max(rows[e], 1), B), dtype=torch.uint8).pin_memory() for e in range(n_exp)]
    for e in range(n_exp):
        rm_host[e][:rows[e]].copy_(dev[e][:, :rows[e]].t())
    torch.cuda.synchronize()
    rm_out = np.zeros(n_counts, dtype=np.int64)
    rm_batch = est.prepare([h.data_ptr() for h in rm_host], rows, [B] * n_exp, prefix, rm_out, rowmajor=True)
    rm_batch.launch()
    assert np.array_equal(rm_out, ref_counts), "row-major ingest counts differ from the column-major run"
    rm_s = timed(rm_batch.launch, args.e2e_steps)
    h2d_bytes = int(sum(rows[e] * B for e in range(n_exp)))

    fit = None
    if not args.no_fit:
        ctx.set_stream(None)
        fit = fit_leg(ctx, torch, dist, world, rank, args.fit_steps, 2, not args.no_cpu and world == 1)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # Launch duration of the parity kernel for the roofline figure: the median of CUDA events that
        # bracket a single launch inside the library (aces_set_timing), on the stream the kernel runs on.
        # The per-step time of the timed region (memset + parity kernel + prefix kernel back to back, and
        # at N > 1 the exposed part of the exchange) is reported beside it.
        k_iso = sorted(kernel_ms)[len(kernel_ms) // 2]
        k_ms = k_iso
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if WORKLOAD == "rotated_d9_wls_1e8shots" and os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        e2e_ms = 1e3 * e2e_s / args.e2e_steps
        plain_ms = 1e3 * plain_s / args.e2e_steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if WL["sharding"] == "experiments" else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world, fd, sizes, alg_bytes),
            "shots_per_s": float(tot[1].item()) / (ms_per_step * 1e-3),
            "clocks": clocks.summary(),
            "exchange": (None if world == 1 else
                         {"collective": (("K1's prefix kernel stores the counts into every rank's buffer over peer memory (CUDA IPC, NVLink stores + "
                                          "release / acquire flags), a pull kernel on a second stream collects them: "
                                          + ("sum of the counts" if reduce_counts else "counts of all repetitions side by side"))
                                         if xch is not None else
                                         ("ncclAllReduce(sum, int64) of the counts" if reduce_counts else "ncclAllGather of the per-repetition counts")),
                          "note": xch_note,
                          "bytes_per_rank": n_counts * 8, "stream": "second stream, overlapped with the next step's kernel (two count buffers)",
                          "counts_equal_on_all_ranks": counts_equal}),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": n_counts * 8, "steps": args.e2e_steps,
                    "ms_per_step": e2e_ms,
                    "api": "aces_parity_counts with pinned host matrices and a host count vector",
                    "h2d_gbs_per_rank": h2d_bytes / (e2e_ms * 1e-3) / 1e9,
                    "plain_h2d_ms_per_step": plain_ms, "plain_h2d_gbs_per_rank": h2d_bytes / (plain_ms * 1e-3) / 1e9,
                    "limiter": ("pcie_h2d: the same bytes as plain pinned cudaMemcpyAsync copies, all ranks at once, take "
                                f"{plain_ms:.1f} ms of the {e2e_ms:.1f} ms" if plain_ms > 0.8 * e2e_ms else "not the copies alone"),
                    "cpu_binding": numa,
                    "rowmajor_ingest": {"api": "aces_parity_counts_rowmajor: Stim-native row-major records from pinned memory "
                                               "(contiguous copy + device transposition + the same kernel)",
                                        "ms_per_step": 1e3 * rm_s / args.e2e_steps,
                                        "value": float(tot[0].item()) * args.e2e_steps / rm_s}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k1::parity_kernel_v2", "kernel_ms": k_ms,
                         "kernel_ms_source": "median of CUDA events bracketing single launches of the parity kernel inside "
                                             "the library (its own stream), 10 launches after the timed region",
                         "step_ms_timed_region": ms_per_step,
                         "frac_of_step": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg_bytes},
        }
        if not args.no_cpu and world == 1:  # the CPU baseline is reported at N=1 only
            line["cpu_baseline"] = cpu_sample(fd, sizes)
        if fit is not None:
            line["fit"] = fit
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
