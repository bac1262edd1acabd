# ACESB200.jl -- Julia shim that binds libaces_b200.so (include/aces_b200.h) with `ccall` and plugs it
# into QuantumACES.jl's estimation hot path without changing user code.
#
# STATUS: written against the C ABI and the reference's Julia signatures, NOT executed: there is
# no `julia` binary in the build image or on the GPU box (DESIGN.md section 1).  What IS checked
# without Julia: every `ccall` type tuple in this file against the prototypes of
# include/aces_b200.h (tools/check_ccall.py, part of the CPU test-suite), and the same call
# sequences with Julia-style 1-based index arrays and pageable host buffers from plain C
# (tests/abi_host.c).  The Python/ctypes twin (quantumaces.jl_b200/_lib.py, merit.py, estimate.py,
# project.py) binds the same entry points with the same argument meaning and is what the
# test-suite exercises on the GPU.
#
# Usage
#     using QuantumACES
#     include("ACESB200.jl"); using .ACESB200
#     ACESB200.init!(; device = 0, lib = "/path/to/libaces_b200.so")
#     ACESB200.install!()        # overwrite the hot-path methods inside QuantumACES
#     aces_data = simulate_aces(d, budget_set; repetitions = 20)      # unchanged user code
#
# What install!() replaces (file:line of the reference):
#   simulate_stim_estimate(d, shots_set; ...)                src/simulate.jl:84-252   sampling loop kept, the
#        estimation inner loop (:195-234) replaced by ONE aces_parity_counts_rowmajor call per sampled
#        batch: Stim's row-major buffer goes through pinned staging, all budgets and both Pauli sets
#        in one pass, batches accumulated instead of vcat'ed
#   estimate_gate_noise(d, ens, cov_ens, shot_budget; ...)    src/estimate.jl:691-941  numeric core on the
#        device-resident Design (covariance, clipping, block factors, OLS / WLS / GLS fits, precision
#        slices, split projection); NoiseEstimate bookkeeping through the reference's own helpers
#   calc_covariance(d, eigenvalues_ensemble, covariance_ensemble; ...)   src/merit.jl:206-291   (S3)
#   calc_gls_covariance / calc_wls_covariance / calc_ols_covariance(d, covariance_log)  src/merit.jl:556-629 (S4)
#   gls_ / wls_ / ols_optimise_weights(d, covariance_log; options)   src/optimise_weights.jl:160-790   the
#        descent loop restated once, its gradient (calc_*_merit_grad_log) on the GPU        (f3)
#   calc_merit(d; warning)                                      src/merit.jl:925-1043    estimator covariances,
#        marginal / relative transforms and the nine eigvals on the GPU in one call (aces_design_merit)  (f2)
#   estimate_stim_ensemble(d_rand, experiment_shots; ...)       src/device.jl:156-244    the job loop with ONE
#        aces_parity_counts launch per job (thousands of 64 ... 10^4-shot circuits as items of one batch)
#   S1  estimate_experiment(::Matrix{UInt8}, ...)             src/estimate.jl:284-301
#   S2  estimate_gate_eigenvalues(A, F, est; constrain)        src/estimate.jl:492-524 (and the OLS form)
# sparse_covariance_inv_factor(covariance_log, lengths) (src/merit.jl:378-427) has no Design argument
# and is NOT overwritten; its callers on the path are (estimate_gate_noise, calc_gls_covariance), and
# a new method sparse_covariance_inv_factor(d::Design, covariance_log; clipped_indices) is added.
# Everything else (Stim circuit strings, NoiseEstimate / Merit bookkeeping, the NRMSE pdf, I/O) stays in
# QuantumACES.
module ACESB200

using SparseArrays, LinearAlgebra, Random, Statistics
import QuantumACES
import QuantumACES: Design

const LIB = Ref{String}("libaces_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

const ACES_E_NOT_PD = Int32(-4)

struct AcesError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::AcesError) = print(io, "libaces_b200 error $(e.code): $(e.msg)")

function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:aces_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[]))
    throw(AcesError(rc, msg))
end

"""
    init!(; device = 0, lib = "libaces_b200.so")

Creates the per-process context on GPU `device` (aces_init).  There is no CPU fallback: this throws
when no sm_100 device is present.
"""
function init!(; device::Integer = 0, lib::String = LIB[])
    LIB[] = lib
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:aces_init, LIB[]), Int32, (Int32, Ptr{Ptr{Cvoid}}), Int32(device), ref)
    if rc != 0
        msg = unsafe_string(ccall((:aces_last_error, LIB[]), Cstring, (Ptr{Cvoid},), C_NULL))
        throw(AcesError(rc, msg))
    end
    CTX[] = ref[]
    return nothing
end

function shutdown!()
    CTX[] == C_NULL && return nothing
    for t in values(TABLES_CACHE)
        ccall((:aces_tables_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], t.handle)
        t.handle = C_NULL
    end
    for dd in values(DESIGN_CACHE)
        ccall((:aces_design_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle)
        dd.handle = C_NULL
    end
    empty!(TABLES_CACHE); empty!(DESIGN_CACHE)
    if PINNED[].ptr != C_NULL
        ccall((:aces_host_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], PINNED[].ptr)
        PINNED[] = PinnedBuffer(C_NULL, 0)
    end
    ccall((:aces_destroy, LIB[]), Int32, (Ptr{Cvoid},), CTX[])
    CTX[] = C_NULL
    return nothing
end

# Pinned host staging for the sampled records (aces_host_alloc): Stim's buffer is copied into it with
# one contiguous memcpy, so the host-to-device copy runs at PCIe line rate instead of through the
# driver's pageable bounce buffers.
struct PinnedBuffer
    ptr::Ptr{Cvoid}
    bytes::Int
end
const PINNED = Ref{PinnedBuffer}(PinnedBuffer(C_NULL, 0))

function pinned_buffer(bytes::Integer)
    if PINNED[].bytes < bytes
        if PINNED[].ptr != C_NULL
            check(ccall((:aces_host_free, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], PINNED[].ptr))
        end
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        want = Int(bytes) + Int(bytes) >> 3
        check(ccall((:aces_host_alloc, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), CTX[], Int64(want), ref))
        PINNED[] = PinnedBuffer(ref[], want)
    end
    return Ptr{UInt8}(PINNED[].ptr)
end

# One PauliTables / DeviceDesign per Design, built on first use (keyed by identity of the Design's
# design matrix, which is not copied when a Design is passed around).
const TABLES_CACHE = IdDict{Any, Any}()
const DESIGN_CACHE = IdDict{Any, Any}()
pauli_tables(d::Design) = get!(() -> PauliTables(d), TABLES_CACHE, d.matrix)
device_design(d::Design) = get!(() -> DeviceDesign(d), DESIGN_CACHE, d.matrix)

# ---------------------------------------------------------------------------------------------
# S1: estimate_experiment (src/estimate.jl:284-301)
# ---------------------------------------------------------------------------------------------

function csr(lists::Vector{Vector{Int}})
    ptr = zeros(Int64, length(lists) + 1)
    for (i, l) in pairs(lists)
        ptr[i + 1] = ptr[i] + length(l)
    end
    flat = Vector{Int32}(undef, max(ptr[end], 1))
    k = 0
    for l in lists, x in l
        flat[k += 1] = Int32(x)
    end
    return ptr, flat
end

"""
Drop-in for `estimate_experiment(counts::Matrix{UInt8}, experiment_meas_indices,
experiment_pauli_signs, shots_value)`.  `counts` is passed in place (column-major, ld = size(counts, 1)).
"""
function estimate_experiment(
    counts::Matrix{UInt8},
    experiment_meas_indices::Vector{Vector{Int}},
    experiment_pauli_signs::Vector{Bool},
    shots_value::Integer,
)
    E = length(experiment_meas_indices)
    @assert E == length(experiment_pauli_signs) "The number of signs does not match the number of circuit eigenvalues."
    ptr, flat = csr(experiment_meas_indices)
    signs = UInt8.(experiment_pauli_signs)
    est = Vector{Float64}(undef, E)
    GC.@preserve counts ptr flat signs est begin
        check(ccall((:aces_estimate_experiment, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{UInt8}, Int64, Ptr{Float64}),
            CTX[], counts, size(counts, 1), max(size(counts, 1), 1), Int32(size(counts, 2)), Int32(E),
            ptr, flat, signs, Int64(shots_value), est))
    end
    return est::Vector{Float64}
end

# ---------------------------------------------------------------------------------------------
# K1 batched: all experiments, all budgets, both Pauli sets in ONE launch
# (the inner loop of simulate_stim_estimate, src/simulate.jl:195-234)
# ---------------------------------------------------------------------------------------------

mutable struct PauliTables
    handle::Ptr{Cvoid}
    exp_ptr::Vector{Int64}      # Paulis of experiment e (tuple-major order)
    n_eig::Vector{Int}          # circuit-eigenvalue Paulis of experiment e (the rest are covariance Paulis)
    sign_factor::Vector{Int}    # (-1)^sign per table entry
    tuple_of::Vector{Int}
    experiment::Vector{Vector{Int}}             # pauli_idx scatter targets (src/simulate.jl:214-219)
    covariance_experiment::Vector{Vector{Int}}  # (src/simulate.jl:227-232)
end

"""
    PauliTables(d::Design)

Flattens get_experiment_data / get_covariance_experiment_data (src/estimate.jl:308-428) into the CSR
tables of aces_tables_create, once per Design.
"""
function PauliTables(d::Design)
    n = d.c.qubit_num
    (experiment_ensemble, meas_indices_ensemble, pauli_sign_ensemble) = QuantumACES.get_experiment_data(d)
    (cov_experiment_ensemble, cov_meas_indices_ensemble, cov_sign_ensemble) =
        QuantumACES.get_covariance_experiment_data(d)
    exp_ptr = Int64[0]; pauli_ptr = Int64[0]; bits = Int32[]
    n_eig = Int[]; sign_factor = Int[]; tuple_of = Int[]
    experiment = Vector{Int}[]; covariance_experiment = Vector{Int}[]
    for i in eachindex(d.tuple_set), j in 1:d.experiment_numbers[i]
        exp = experiment_ensemble[i][j]
        for p in exp
            append!(bits, Int32.(meas_indices_ensemble[i][p] .- 1))   # record bit = qubit - 1
            push!(pauli_ptr, length(bits))
            push!(sign_factor, pauli_sign_ensemble[i][p] ? -1 : 1)
        end
        cexp = length(cov_meas_indices_ensemble[i]) > 0 ? cov_experiment_ensemble[i][j] : Int[]
        for c in cexp
            append!(bits, Int32.(cov_meas_indices_ensemble[i][c] .- 1))
            push!(pauli_ptr, length(bits))
            push!(sign_factor, cov_sign_ensemble[i][c] ? -1 : 1)
        end
        push!(exp_ptr, exp_ptr[end] + length(exp) + length(cexp))
        push!(n_eig, length(exp)); push!(tuple_of, i)
        push!(experiment, exp); push!(covariance_experiment, cexp)
    end
    isempty(bits) && push!(bits, Int32(0))
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve exp_ptr pauli_ptr bits begin
        check(ccall((:aces_tables_create, LIB[]), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
            CTX[], Int32(n), Int32(length(exp_ptr) - 1), exp_ptr, pauli_ptr, bits, ref))
    end
    t = PauliTables(ref[], exp_ptr, n_eig, sign_factor, tuple_of, experiment, covariance_experiment)
    finalizer(x -> (CTX[] == C_NULL || ccall((:aces_tables_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.handle)), t)
    return t
end

"""
    estimate_counts!(eig_set, cov_set, tables, stim_counts, shots_divided)

The estimation half of simulate_stim_estimate for ALL experiments at once: `stim_counts[e]` is the
`Matrix{UInt8}` of table experiment `e`, `shots_divided[i, s]` as in src/simulate.jl:103.  Pushes
`((-1)^sign * pauli_shots) / shots_value` (src/estimate.jl:297-298, formed here in Float64 from the
exact integer counts) into `eig_set[s][i][pauli_idx]` / `cov_set[s][i][pauli_idx]`.
"""
function estimate_counts!(eig_set, cov_set, t::PauliTables, stim_counts::Vector{Matrix{UInt8}},
                          shots_divided::Matrix{Int})
    n_exp = length(stim_counts)
    K = size(shots_divided, 2)
    order = sortperm(shots_divided[1, :])            # the kernel wants nested prefixes
    ids = Int32.(0:(n_exp - 1))
    ptrs = [pointer(m) for m in stim_counts]
    rows = Int64[size(m, 1) for m in stim_counts]
    ld = max.(rows, 1)
    prefix = Matrix{Int64}(undef, K, n_exp)          # item-major, K fastest
    for e in 1:n_exp
        prefix[:, e] = shots_divided[t.tuple_of[e], order]
    end
    odd = zeros(Int64, K * t.exp_ptr[end])
    GC.@preserve stim_counts ids ptrs rows ld prefix odd begin
        check(ccall((:aces_parity_counts, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Ptr{UInt8}}, Ptr{Int64}, Ptr{Int64}, Int32,
             Ptr{Int64}, Ptr{Int64}, Int32),
            CTX[], t.handle, Int32(n_exp), ids, ptrs, rows, ld, Int32(K), prefix, odd, Int32(0)))
    end
    off = 0
    for e in 1:n_exp
        E = Int(t.exp_ptr[e + 1] - t.exp_ptr[e]); ne = t.n_eig[e]; i = t.tuple_of[e]
        sf = view(t.sign_factor, (t.exp_ptr[e] + 1):t.exp_ptr[e + 1])
        for kk in 1:K
            s = order[kk]; S = prefix[kk, e]
            for idx in 1:E
                pauli_shots = S - 2 * odd[off + (kk - 1) * E + idx]
                est = (sf[idx] * pauli_shots) / S
                if idx <= ne
                    push!(eig_set[s][i][t.experiment[e][idx]], est)
                else
                    push!(cov_set[s][i][t.covariance_experiment[e][idx - ne]], est)
                end
            end
        end
        off += E * K
    end
    return nothing
end

"""
    estimate_job!(eig_ens, cov_ens, tables, job_counts, circuit_indices, pauli_sign_ensemble,
                  covariance_pauli_sign_ensemble, experiment_shots, exp_offset)

One job of a randomised design (src/device.jl:190-229): circuit `c` ran experiment `j` of tuple `i` under
randomisation `r` (`circuit_indices[c] = (i, j, r)`).  Where the reference calls estimate_experiment
twice per circuit, ALL circuits of the job go through ONE aces_parity_counts launch (n_items = circuits,
K = 1; both Pauli sets of an experiment in the same pass; the library packs the small host matrices
into one pinned upload).  Estimates are pushed in the reference's order.
"""
function estimate_job!(eig_ens, cov_ens, t::PauliTables, job_counts::Vector{Matrix{UInt8}},
                       circuit_indices, pauli_sign_ensemble, covariance_pauli_sign_ensemble,
                       experiment_shots::Integer, exp_offset::Vector{Int})
    n_items = length(job_counts)
    n_items == 0 && return nothing
    eids = [exp_offset[i] + j for (i, j, r) in circuit_indices]          # 1-based table experiments
    ids = Int32.(eids .- 1)
    ptrs = [pointer(m) for m in job_counts]
    rows = Int64[size(m, 1) for m in job_counts]
    ld = max.(rows, 1)
    prefix = fill(Int64(experiment_shots), n_items)
    total = sum(Int(t.exp_ptr[e + 1] - t.exp_ptr[e]) for e in eids)
    odd = zeros(Int64, total)
    GC.@preserve job_counts ids ptrs rows ld prefix odd begin
        check(ccall((:aces_parity_counts, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Ptr{UInt8}}, Ptr{Int64}, Ptr{Int64}, Int32,
             Ptr{Int64}, Ptr{Int64}, Int32),
            CTX[], t.handle, Int32(n_items), ids, ptrs, rows, ld, Int32(1), prefix, odd, Int32(0)))
    end
    S = Int(experiment_shots)
    off = 0
    for (c, (i, j, r)) in pairs(circuit_indices)
        e = eids[c]
        E = Int(t.exp_ptr[e + 1] - t.exp_ptr[e]); ne = t.n_eig[e]
        signs = pauli_sign_ensemble[i][j][r]
        @assert length(signs) == ne "The number of signs does not match the number of circuit eigenvalues."
        for idx in 1:ne
            pauli_shots = S - 2 * odd[off + idx]
            push!(eig_ens[i][t.experiment[e][idx]], ((-1)^signs[idx] * pauli_shots) / S)
        end
        if E > ne
            csigns = covariance_pauli_sign_ensemble[i][j][r]
            for idx in 1:(E - ne)
                pauli_shots = S - 2 * odd[off + ne + idx]
                push!(cov_ens[i][t.covariance_experiment[e][idx]], ((-1)^csigns[idx] * pauli_shots) / S)
            end
        end
        off += E
    end
    return nothing
end

"""
Drop-in for `estimate_stim_ensemble(d_rand, experiment_shots; fail_jobs, simulation_seed, min_eigenvalue,
split, precision)` (src/device.jl:156-244): the reference's job loop with one launch per job
(`estimate_job!`), then the reference's estimate_gate_noise(d_rand, ...) (whose numeric core install!()
routes to the GPU as well).
"""
function estimate_stim_ensemble(d_rand, experiment_shots::Integer; fail_jobs::Vector{Int} = Int[],
                                simulation_seed::Union{UInt64, Nothing} = nothing, min_eigenvalue::Real = 0.1,
                                split::Bool = false, precision::Real = 1e-8)
    d = d_rand.d
    tuple_number = length(d.tuple_set)
    backend = simulation_seed !== nothing ? "stim_$(simulation_seed)" : "stim"
    t = get!(() -> PauliTables(d), TABLES_CACHE, d.matrix)
    mapping_lengths = length.(d.mapping_ensemble)
    covariance_mapping_lengths = length.(QuantumACES.get_covariance_experiment_data(d)[2])   # src/device.jl:181-184
    exp_offset = [0; cumsum(d.experiment_numbers)[1:(end - 1)]]
    job_indices = d_rand.job_indices
    job_number = length(job_indices)
    @assert all(fail_jobs .> 1) && all(fail_jobs .<= job_number) "The failed job indices are not consistent with the number of jobs."
    success_jobs = setdiff(1:job_number, fail_jobs)
    eig_ens = [[Float64[] for j in 1:mapping_lengths[i]] for i in 1:tuple_number]
    cov_ens = [[Float64[] for j in 1:covariance_mapping_lengths[i]] for i in 1:tuple_number]
    for job_idx in success_jobs
        job_counts = QuantumACES.load_rand_design_job(d_rand, backend, job_idx)
        @assert length(job_counts) == length(job_indices[job_idx]) "The number of circuits in the job does not match the number of counts."
        estimate_job!(eig_ens, cov_ens, t, Vector{Matrix{UInt8}}(job_counts), job_indices[job_idx],
                      d_rand.pauli_sign_ensemble, d_rand.covariance_pauli_sign_ensemble, experiment_shots, exp_offset)
    end
    return QuantumACES.estimate_gate_noise(d_rand, eig_ens, cov_ens; min_eigenvalue = min_eigenvalue, split = split,
                                           precision = precision)
end

"""
    odd_counts_rowmajor!(odd, tables, e, records, rows, prefixes; accumulate)

Odd-parity counts of ONE sampled batch of table experiment `e` (1-based): `records` points at `rows`
row-major records of `B` bytes (Stim's `sample(bit_packed = true)` buffer, here already in pinned
memory), `prefixes[k]` = how many of these rows belong to budget `k` (nondecreasing).  The counts of
all budgets and of both Pauli sets come from one pass; `accumulate = true` adds to `odd`
(batched sampling, src/simulate.jl:174-186, without the vcat).
"""
function odd_counts_rowmajor!(odd::Vector{Int64}, t::PauliTables, e::Int, records::Ptr{UInt8}, rows::Int,
                              B::Int, prefixes::Vector{Int64}; accumulate::Bool = false)
    ids = Int32[e - 1]; ptrs = Ptr{UInt8}[records]; nrows = Int64[rows]; pitch = Int64[B]
    GC.@preserve ids ptrs nrows pitch prefixes odd begin
        check(ccall((:aces_parity_counts_rowmajor, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Ptr{UInt8}}, Ptr{Int64}, Ptr{Int64}, Int32,
             Ptr{Int64}, Ptr{Int64}, Int32),
            CTX[], t.handle, Int32(1), ids, ptrs, nrows, pitch, Int32(length(prefixes)), prefixes, odd,
            Int32(accumulate)))
    end
    return odd
end

"""
Drop-in for `simulate_stim_estimate(d, shots_set; seed, max_samples, detailed_diagnostics)`
(src/simulate.jl:84-252).  The sampling half is the reference's (same shot division :98-104, same seeds
:105-121, same circuit strings :146-170, same batches :172-186, Stim on thread 1); the estimation half
(:188-234) is one `aces_parity_counts_rowmajor` call per sampled batch: the row-major numpy buffer is
copied once into pinned memory (no pyconvert to a column-major Matrix, no vcat of the batches), every
budget prefix and both Pauli sets are reduced in the same pass, and the estimates
`((-1)^sign * pauli_shots) / shots_value` (src/estimate.jl:297-298) are formed here in Float64 from the
exact integer counts, then pushed in the reference's order (:214-233).
"""
function simulate_stim_estimate(d::Design, shots_set::Vector{Int}; seed::Union{UInt64, Nothing} = nothing,
                                max_samples::Integer = 10^9, detailed_diagnostics::Bool = false)
    Q = QuantumACES
    start_time = time()
    tuple_number = length(d.tuple_set)
    experiment_numbers = d.experiment_numbers
    n = d.c.qubit_num
    B = cld(n, 8)
    budget_count = length(shots_set)
    @assert sum(d.shot_weights) ≈ 1.0 "The shot weights are not appropriately normalised."
    @assert all(d.shot_weights .> 0.0) "The shot weights are not all positive."
    shots_divided_float = [shots_set[s] * d.shot_weights[t] / experiment_numbers[t] for t in 1:tuple_number, s in 1:budget_count]
    @assert vec(sum(shots_divided_float .* experiment_numbers; dims = 1)) ≈ shots_set "The shots have not been divided correctly."
    shots_divided = ceil.(Int, shots_divided_float)
    shots_maximum = vec(maximum(shots_divided; dims = 2))
    # seeds exactly as the reference draws them
    seed !== nothing && Q.Random.seed!(seed)
    batched_shots = any(shots_maximum * n .> max_samples)
    stim_seeds = batched_shots ?
        [[rand(UInt64, length(Q.batch_shots(shots_maximum[i], n, max_samples))) for j in 1:experiment_numbers[i]] for i in 1:tuple_number] :
        [rand(UInt64, experiment_numbers[i]) for i in 1:tuple_number]
    seed !== nothing && Q.Random.seed!()
    t = pauli_tables(d)
    mapping_lengths = length.(d.mapping_ensemble)
    covariance_mapping_lengths = [length(ks) for ks in Q.get_covariance_mapping_data(d)[1]]
    eig_set = [[[Float64[] for j in 1:mapping_lengths[i]] for i in 1:tuple_number] for s in 1:budget_count]
    cov_set = [[[Float64[] for j in 1:covariance_mapping_lengths[i]] for i in 1:tuple_number] for s in 1:budget_count]
    order = sortperm(shots_set)                  # the kernel wants nested prefixes; budgets may be unsorted
    e = 0
    for i in 1:tuple_number
        tuple_circuit_string = Q.get_stim_circuit(d.c.circuit[d.tuple_set[i]], d.c.gate_probabilities, d.c.noisy_prep, d.c.noisy_meas)
        for j in 1:experiment_numbers[i]
            e += 1
            prep = Q.get_stim_circuit([d.prep_ensemble[i][j]], d.c.gate_probabilities, d.c.noisy_prep, d.c.noisy_meas;
                extra_fields = d.c.extra_fields)
            meas = Q.get_stim_circuit([d.meas_ensemble[i][j]], d.c.gate_probabilities, d.c.noisy_prep, d.c.noisy_meas)
            circuit_string = prep * tuple_circuit_string * meas
            shot_batches = batched_shots ? Q.batch_shots(shots_maximum[i], n, max_samples) : [shots_maximum[i]]
            E = Int(t.exp_ptr[e + 1] - t.exp_ptr[e])
            odd = zeros(Int64, budget_count * E)
            done = 0
            for (b, batch) in pairs(shot_batches)
                @assert Threads.threadid() == 1 "Stim sampling cannot be multithreaded."
                sampler = Q.stim.Circuit(circuit_string).compile_sampler(; seed = batched_shots ? stim_seeds[i][j][b] : stim_seeds[i][j])
                records = Q.PythonCall.PyArray(sampler.sample(; shots = batch, bit_packed = true))   # row-major, no copy
                @assert size(records) == (batch, B) && strides(records) == (B, 1) "Unexpected layout of the Stim sample buffer."
                staged = pinned_buffer(batch * B)
                GC.@preserve records unsafe_copyto!(staged, Ptr{UInt8}(pointer(records)), batch * B)
                prefixes = Int64[clamp(shots_divided[i, order[k]] - done, 0, batch) for k in 1:budget_count]
                odd_counts_rowmajor!(odd, t, e, staged, batch, B, prefixes; accumulate = b > 1)
                done += batch
            end
            ne = t.n_eig[e]
            sf = view(t.sign_factor, (t.exp_ptr[e] + 1):t.exp_ptr[e + 1])
            for kk in 1:budget_count
                s = order[kk]; S = shots_divided[i, s]
                for idx in 1:E
                    est = (sf[idx] * (S - 2 * odd[(kk - 1) * E + idx])) / S
                    if idx <= ne
                        push!(eig_set[s][i][t.experiment[e][idx]], est)
                    else
                        push!(cov_set[s][i][t.covariance_experiment[e][idx - ne]], est)
                    end
                end
            end
            if detailed_diagnostics
                println("Simulated sampling experiment $(j) of $(experiment_numbers[i]) for tuple $(i) of $(tuple_number). The time elapsed since simulation started is $(round(time() - start_time, digits = 3)) s.")
            end
        end
    end
    return (eig_set::Vector{Vector{Vector{Vector{Float64}}}}, cov_set::Vector{Vector{Vector{Vector{Float64}}}})
end

# ---------------------------------------------------------------------------------------------
# S2: estimate_gate_eigenvalues (src/estimate.jl:492-524)
# ---------------------------------------------------------------------------------------------

function estimate_gate_eigenvalues(
    design_matrix::SparseMatrixCSC{Float64, Int},
    est_covariance_log_inv_factor::Union{SparseMatrixCSC{Float64, Int}, Nothing},
    est_eigenvalues::Vector{Float64};
    constrain::Bool = false,
)
    (M, N) = size(design_matrix)
    A = design_matrix; F = est_covariance_log_inv_factor
    g = Vector{Float64}(undef, N)
    fc = F === nothing ? Ptr{Int64}(C_NULL) : pointer(F.colptr)
    fr = F === nothing ? Ptr{Int64}(C_NULL) : pointer(F.rowval)
    fz = F === nothing ? Ptr{Float64}(C_NULL) : pointer(F.nzval)
    GC.@preserve A F est_eigenvalues g begin
        check(ccall((:aces_estimate_gate_eigenvalues, LIB[]), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
             Int32, Ptr{Float64}, Int32, Ptr{Float64}),
            CTX[], M, N, A.colptr, A.rowval, A.nzval, fc, fr, fz, Int32(1), est_eigenvalues,
            Int32(constrain), g))
    end
    return g::Vector{Float64}
end
estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64, Int}, est_eigenvalues::Vector{Float64};
                          constrain::Bool = false) =
    estimate_gate_eigenvalues(design_matrix, nothing, est_eigenvalues; constrain = constrain)

# ---------------------------------------------------------------------------------------------
# Design-resident path
# ---------------------------------------------------------------------------------------------

mutable struct DeviceDesign
    handle::Ptr{Cvoid}
    d::Design
    M::Int
    N::Int
    C::Int
    mapping_lengths::Vector{Int}
    cov_colptr::Vector{Int64}   # 0-based pattern of every covariance matrix of this design
    cov_rowval::Vector{Int64}
end

"""
    DeviceDesign(d::Design)

Uploads what the fit path reads from `d` (aces_design_create): d.matrix as stored
(SparseMatrixCSC{Int32,Int32}), experiment numbers, shot weights, tuple times, and the
covariance dictionary counts in the key order of get_covariance_mapping_data (src/merit.jl:144-163).
"""
function DeviceDesign(d::Design)
    T = length(d.tuple_set)
    mapping_lengths = length.(d.mapping_ensemble)
    (keys_ensemble, _) = QuantumACES.get_covariance_mapping_data(d)
    diag_counts = Int64[]
    for i in 1:T, p in 1:mapping_lengths[i]
        push!(diag_counts, d.covariance_dict_ensemble[i][CartesianIndex(p, p)][2])
    end
    cov_ptr = Int64[0]; cov_p = Int64[]; cov_q = Int64[]; cov_c = Int64[]
    for i in 1:T
        for key in keys_ensemble[i]
            push!(cov_p, key[1]); push!(cov_q, key[2])
            push!(cov_c, d.covariance_dict_ensemble[i][key][2])
        end
        push!(cov_ptr, length(cov_p))
    end
    A = d.matrix
    ml = Int64.(mapping_lengths); X = Int64.(d.experiment_numbers)
    w = Float64.(d.shot_weights); tau = Float64.(d.tuple_times)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    has_cov = length(cov_p) > 0
    GC.@preserve A ml X w tau diag_counts cov_ptr cov_p cov_q cov_c begin
        check(ccall((:aces_design_create, LIB[]), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Cvoid}}),
            CTX[], Int32(T), ml, size(A, 2), A.colptr, A.rowval, A.nzval, Int32(1), X, w, tau, diag_counts,
            has_cov ? pointer(cov_ptr) : Ptr{Int64}(C_NULL), cov_p, cov_q, cov_c, ref))
    end
    nnz_cov = ccall((:aces_design_covariance_nnz, LIB[]), Int64, (Ptr{Cvoid},), ref[])
    colptr = Vector{Int64}(undef, size(A, 1) + 1); rowval = Vector{Int64}(undef, nnz_cov)
    check(ccall((:aces_design_covariance_pattern, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}),
        ref[], colptr, rowval))
    dd = DeviceDesign(ref[], d, size(A, 1), size(A, 2), length(cov_p), mapping_lengths, colptr, rowval)
    finalizer(x -> (CTX[] == C_NULL || ccall((:aces_design_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.handle)), dd)
    return dd
end

pattern_matrix(dd::DeviceDesign, nzval::Vector{Float64}) =
    SparseMatrixCSC{Float64, Int}(dd.M, dd.M, Int.(dd.cov_colptr .+ 1), Int.(dd.cov_rowval .+ 1), nzval)

"""
calc_covariance(d, eigenvalues_ensemble, covariance_ensemble; epsilon, weight_time)
(src/merit.jl:206-291); `log_form = true` also applies calc_covariance_log (src/merit.jl:356-365).
"""
function calc_covariance(dd::DeviceDesign, eigenvalues_ensemble::Vector{Vector{Float64}},
                         covariance_ensemble::Vector{Vector{Float64}};
                         epsilon::Real = 1e-10, weight_time::Bool = true, log_form::Bool = false)
    lam = vcat(eigenvalues_ensemble...); lc = vcat(covariance_ensemble...)
    nz = Vector{Float64}(undef, length(dd.cov_rowval))
    GC.@preserve lam lc nz begin
        check(ccall((:aces_calc_covariance, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Int32, Int32, Ptr{Float64}),
            CTX[], dd.handle, lam, dd.C > 0 ? pointer(lc) : Ptr{Float64}(C_NULL), Float64(epsilon),
            Int32(weight_time), Int32(log_form), nz))
    end
    return pattern_matrix(dd, nz)
end

"""
    fit(dd, ls_type, est_eigenvalues, est_covariance_eigenvalues; clipped_indices, constrain)

Unprojected gate eigenvalues of one estimator (:ols, :wls, :gls) exactly as estimate_gate_noise forms
them (src/estimate.jl:711-743, 820-831), in one call: covariance of the estimates, log-covariance,
clipping, inverse factor (with the reference's not-PD fallback), whitened least squares.
"""
function fit(dd::DeviceDesign, ls_type::Symbol, est_eigenvalues::Vector{Float64},
             est_covariance_eigenvalues::Vector{Float64};
             clipped_indices::Union{Vector{Int}, Nothing} = nothing, epsilon::Real = 1e-10, constrain::Bool = false)
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : ls_type == :gls ? 2 :
           throw(error("The least squares type $(ls_type) must be either :gls, :wls, or :ols."))
    g = Vector{Float64}(undef, dd.N)
    kept = clipped_indices === nothing ? Int64[] : Int64.(clipped_indices)
    GC.@preserve est_eigenvalues est_covariance_eigenvalues kept g begin
        check(ccall((:aces_design_fit, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Float64, Int32,
             Ptr{Float64}),
            CTX[], dd.handle, Int32(kind), est_eigenvalues,
            dd.C > 0 ? pointer(est_covariance_eigenvalues) : Ptr{Float64}(C_NULL),
            clipped_indices === nothing ? Ptr{Int64}(C_NULL) : pointer(kept), length(kept), Int32(1),
            Float64(epsilon), Int32(constrain), g))
    end
    return g
end

"""
    fit_tuple_sharded(dd, ls_type, est_eigenvalues, est_covariance_eigenvalues, tuple_on, allreduce!; ...)

One fit with the tuple blocks split over several GPUs (one process per GPU, each with its own
`DeviceDesign` of the same `Design`): this process works on the tuples with `tuple_on[t] == true`.
`allreduce!(ptr::Ptr{Float64}, n::Int)` must sum `n` doubles at the DEVICE pointer `ptr` in place
over the processes (NCCL.jl `Allreduce!` on an unsafe_wrap'd CuArray, or a CUDA-aware MPI.jl
`Allreduce!`).  The Gram matrix and the right-hand sides are summed at the reduction points of
include/aces_b200.h (`aces_design_fit_begin` .. `aces_design_fit_end`); every process returns the
same gate eigenvalues.
"""
function fit_tuple_sharded(dd::DeviceDesign, ls_type::Symbol, est_eigenvalues::Vector{Float64},
                           est_covariance_eigenvalues::Vector{Float64}, tuple_on::Vector{Bool}, allreduce!::Function;
                           clipped_indices::Union{Vector{Int}, Nothing} = nothing, epsilon::Real = 1e-10,
                           constrain::Bool = false, distributed::Union{Nothing, Tuple{Int, Int, Function}} = nothing)
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : ls_type == :gls ? 2 :
           throw(error("The least squares type $(ls_type) must be either :gls, :wls, or :ols."))
    on = UInt8.(tuple_on)
    g = Vector{Float64}(undef, dd.N)
    kept = clipped_indices === nothing ? Int64[] : Int64.(clipped_indices)
    gram = Ref{Ptr{Float64}}(C_NULL); rhs = Ref{Ptr{Float64}}(C_NULL)
    n_gram = Ref{Int64}(0); n_rhs = Ref{Int64}(0)
    GC.@preserve on est_eigenvalues est_covariance_eigenvalues kept g begin
        check(ccall((:aces_design_set_tuple_shard, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}), CTX[], dd.handle, on))
        try
            check(ccall((:aces_design_fit_begin, LIB[]), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Float64),
                CTX[], dd.handle, Int32(kind), est_eigenvalues,
                dd.C > 0 ? pointer(est_covariance_eigenvalues) : Ptr{Float64}(C_NULL),
                clipped_indices === nothing ? Ptr{Int64}(C_NULL) : pointer(kept), length(kept), Int32(1), Float64(epsilon)))
            check(ccall((:aces_design_fit_buffers, LIB[]), Int32,
                (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Int64}, Ptr{Ptr{Float64}}, Ptr{Int64}), dd.handle, gram, n_gram, rhs, n_rhs))
            check(ccall((:aces_synchronize, LIB[]), Int32, (Ptr{Cvoid},), CTX[]))
            allreduce!(gram[], Int(n_gram[]))
            allreduce!(rhs[], Int(n_rhs[]))
            if distributed === nothing
                check(ccall((:aces_design_fit_solve, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle))
            else
                (rank, world, broadcast!) = distributed
                factor_distributed!(dd, gram[], rank, world, broadcast!)
                check(ccall((:aces_design_fit_dist_solve, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle))
            end
            for it in 1:2
                check(ccall((:aces_design_fit_refine_begin, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle))
                check(ccall((:aces_synchronize, LIB[]), Int32, (Ptr{Cvoid},), CTX[]))
                allreduce!(rhs[], Int(n_rhs[]))
                xm = Ref{Float64}(0.0); dm = Ref{Float64}(0.0)
                check(ccall((:aces_design_fit_refine_end, LIB[]), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), CTX[], dd.handle, xm, dm))
                (it == 1 && dm[] <= 1e-7 * xm[]) && break
            end
            check(ccall((:aces_design_fit_end, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}),
                CTX[], dd.handle, Int32(constrain), g))
        finally
            ccall((:aces_design_set_tuple_shard, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}), CTX[], dd.handle, C_NULL)
        end
    end
    return g
end

"""
    factor_distributed!(dd, gram, rank, world, broadcast!; panel = 512)

The Cholesky factorisation of the summed Gram matrix distributed over `world` processes (0-based `rank`):
block columns of `panel` columns owned cyclically; the owner factorises its block column in place, then
`broadcast!(ptr, n, src)` -- `n` doubles at a DEVICE pointer from process `src` (NCCL.jl Broadcast! /
CUDA-aware MPI.Bcast!) -- distributes its full columns and the inverses of their 128-blocks, and every
process applies the panel to the later block columns it owns.  In-order version of
DeviceDesign.factor_distributed of the Python host (which adds look-ahead on a second stream).
"""
function factor_distributed!(dd::DeviceDesign, gram::Ptr{Float64}, rank::Int, world::Int, broadcast!::Function;
                             panel::Int = 512)
    dinv = Ref{Ptr{Float64}}(C_NULL); n_dinv = Ref{Int64}(0); dead = Ref{Ptr{Int32}}(C_NULL)
    check(ccall((:aces_design_fit_dist_buffers, LIB[]), Int32,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Int64}, Ptr{Ptr{Int32}}), dd.handle, dinv, n_dinv, dead))
    Np = Int(n_dinv[]) ÷ 128
    panels = [(c, min(panel, Np - c)) for c in 0:panel:(Np - 1)]
    check(ccall((:aces_design_fit_dist_begin, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle))
    for (p, (col0, ncols)) in enumerate(panels)
        owner = (p - 1) % world
        if rank == owner
            check(ccall((:aces_design_fit_dist_factor, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64),
                CTX[], dd.handle, Int64(col0), Int64(ncols)))
        end
        if world > 1
            check(ccall((:aces_synchronize, LIB[]), Int32, (Ptr{Cvoid},), CTX[]))
            broadcast!(gram + 8 * col0 * Np, ncols * Np, owner)
            broadcast!(dinv[] + 8 * col0 * 128, ncols * 128, owner)
        end
        mine = [(c, n) for (q, (c, n)) in enumerate(panels) if q > p && (q - 1) % world == rank]
        if !isempty(mine)
            t0 = Int64[c for (c, n) in mine]; tn = Int64[n for (c, n) in mine]
            GC.@preserve t0 tn begin
                check(ccall((:aces_design_fit_dist_update, LIB[]), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int64}, Ptr{Int64}),
                    CTX[], dd.handle, Int64(col0), Int64(ncols), Int32(length(mine)), t0, tn))
            end
        end
    end
    return nothing
end

"""
S3: sparse_covariance_inv_factor(covariance_log, mapping_lengths; epsilon) (src/merit.jl:378-427) for
a covariance_log in the design's pattern, optionally restricted to `clipped_indices`.
"""
function sparse_covariance_inv_factor(dd::DeviceDesign, covariance_log::SparseMatrixCSC{Float64, Int};
                                      clipped_indices::Union{Vector{Int}, Nothing} = nothing,
                                      epsilon::Real = 1e-12)
    nz = [covariance_log[dd.cov_rowval[e] + 1, c] for c in 1:dd.M for e in (dd.cov_colptr[c] + 1):dd.cov_colptr[c + 1]]
    lens = clipped_indices === nothing ? dd.mapping_lengths :
           QuantumACES.get_clipped_mapping_lengths(dd.mapping_lengths, clipped_indices)
    blocks = Vector{Float64}(undef, sum(lens .^ 2)); got = Vector{Int64}(undef, length(lens))
    kept = clipped_indices === nothing ? Int64[] : Int64.(clipped_indices)
    GC.@preserve nz kept blocks got begin
        check(ccall((:aces_design_inv_factor, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Float64, Ptr{Float64}, Ptr{Int64}),
            CTX[], dd.handle, nz, clipped_indices === nothing ? Ptr{Int64}(C_NULL) : pointer(kept), length(kept),
            Int32(1), Float64(epsilon), blocks, got))
    end
    mats = SparseMatrixCSC{Float64, Int}[]; off = 0
    for k in lens
        push!(mats, sparse(reshape(view(blocks, (off + 1):(off + k * k)), k, k))); off += k * k
    end
    return blockdiag(mats...)::SparseMatrixCSC{Float64, Int}
end

"""
calc_gls_covariance / calc_wls_covariance / calc_ols_covariance(d, covariance_log) (src/merit.jl:556-629)
and the trace moments nrmse_moments needs (src/merit.jl:672-682).  `want_matrix = false` skips the
N x N device-to-host copy.
"""
function calc_ls_covariance(dd::DeviceDesign, covariance_log::SparseMatrixCSC{Float64, Int};
                            ls_type::Symbol = :gls, want_matrix::Bool = true)
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : 2
    nz = [covariance_log[dd.cov_rowval[e] + 1, c] for c in 1:dd.M for e in (dd.cov_colptr[c] + 1):dd.cov_colptr[c + 1]]
    g = QuantumACES.get_gate_eigenvalues(dd.d)
    sigma = want_matrix ? Matrix{Float64}(undef, dd.N, dd.N) : Matrix{Float64}(undef, 0, 0)
    tr = Ref{Float64}(0.0); trsq = Ref{Float64}(0.0)
    GC.@preserve nz g sigma begin
        check(ccall((:aces_design_ls_covariance, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], dd.handle, Int32(kind), g, nz, want_matrix ? pointer(sigma) : Ptr{Float64}(C_NULL), tr, trsq))
    end
    return (want_matrix ? Symmetric(sigma) : nothing, tr[], trsq[])
end

# ---------------------------------------------------------------------------------------------
# Simplex projection (src/estimate.jl:580-663, src/utils.jl:36-97) -- SURVEY.md section 8f, row f1
# ---------------------------------------------------------------------------------------------

"Gate index groups of `gate_data` as the CSR arrays of include/aces_b200.h (1-based indices kept)."
function gate_groups(gate_data)
    ptr = Int64[0]; idx = Int32[]
    for gate_idx in gate_data.gate_indices
        append!(idx, Int32.(gate_idx.indices))
        push!(ptr, length(idx))
    end
    return ptr, idx
end

"""
Drop-in for `simple_project_gate_eigenvalues(est_unproj_gate_eigenvalues, gate_data)`
(src/estimate.jl:635-663): all gates in one launch, same order of floating-point operations.
Requires the padded indices of the gates to follow each other (`pad_indices` of gate g start at
`ptr[g] + g`, as get_gate_data builds them for a design that is not `combined`).
"""
function simple_project_gate_eigenvalues(est_unproj_gate_eigenvalues::Vector{Float64}, gate_data)
    ptr, idx = gate_groups(gate_data)
    N = length(est_unproj_gate_eigenvalues); G = length(ptr) - 1
    out = Vector{Float64}(undef, N)
    unproj_probs = Vector{Float64}(undef, length(idx) + G); probs = similar(unproj_probs)
    GC.@preserve ptr idx est_unproj_gate_eigenvalues out unproj_probs probs begin
        check(ccall((:aces_project_simplex, LIB[]), Int32,
            (Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], Int64(N), Int32(G), ptr, idx, Int32(1), est_unproj_gate_eigenvalues, out, unproj_probs, probs))
    end
    return (out, unproj_probs, probs)
end

"Symplectically ordered Walsh-Hadamard block of a gate with `k` eigenvalues (src/noise.jl:285-312)."
function wht_block(k::Int)
    k == 1 && return [1.0 1.0; 1.0 -1.0]
    n = k == 3 ? 1 : k == 15 ? 2 : throw(error("Unsupported gate size $(k)."))
    P = k + 1; mask = (1 << n) - 1
    return [isodd(count_ones((a & mask) & (b >> n)) + count_ones((a >> n) & (b & mask))) ? -1.0 : 1.0 for a in 0:(P - 1), b in 0:(P - 1)]
end

"""
`split_project_gate_eigenvalues` (src/estimate.jl:580-628) for the estimator of the LAST `fit(dd, ls_type, ...)`:
the per-gate slices of the precision matrix come from the device-resident Gram matrix
(aces_design_precision_blocks), the per-gate quadratic programs are solved in one launch
(aces_project_nonnegative_batched; exact where the reference iterates SCS to `precision`).
"""
function split_project_gate_eigenvalues(dd::DeviceDesign, ls_type::Symbol, est_unproj_gate_eigenvalues::Vector{Float64},
                                        gate_data; precision::Real = 1e-8)
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : 2
    ptr, idx = gate_groups(gate_data)
    G = length(ptr) - 1
    sizes = diff(ptr)
    blocks = Vector{Float64}(undef, sum(sizes .^ 2))
    GC.@preserve ptr idx est_unproj_gate_eigenvalues blocks begin
        check(ccall((:aces_design_precision_blocks, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}),
            CTX[], dd.handle, Int32(kind), Int32(G), ptr, idx, Int32(1), est_unproj_gate_eigenvalues, blocks))
    end
    # per gate: unprojected probabilities, T = W[2:end, 2:end] .- W[2:end, 1], Q = T * P_gg * T'
    v = Vector{Float64}(undef, length(idx)); Q = Vector{Float64}(undef, length(blocks))
    qs = Vector{Vector{Float64}}(undef, G); Ws = Vector{Matrix{Float64}}(undef, G)
    boff = 0
    for g in 1:G
        k = Int(sizes[g]); rng = (ptr[g] + 1):ptr[g + 1]
        W = wht_block(k)
        q = (W ./ (k + 1)) * vcat(1.0, est_unproj_gate_eigenvalues[idx[rng]])
        T = W[2:end, 2:end] .- W[2:end, 1]
        Pgg = reshape(view(blocks, (boff + 1):(boff + k * k)), k, k)
        Q[(boff + 1):(boff + k * k)] = vec(T * Pgg * T')
        v[rng] = q[2:end]
        qs[g] = q; Ws[g] = W
        boff += k * k
    end
    y = Vector{Float64}(undef, length(idx)); failed = Ref{Int32}(0)
    GC.@preserve ptr Q v y begin
        check(ccall((:aces_project_nonnegative_batched, LIB[]), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Int32}),
            CTX[], Int32(G), ptr, Q, v, Float64(precision), y, failed))
    end
    est_gate_eigenvalues = copy(est_unproj_gate_eigenvalues)
    unproj_probs = Float64[]; probs = Float64[]
    for g in 1:G
        rng = (ptr[g] + 1):ptr[g + 1]
        pr = vcat(1.0 - sum(y[rng]), y[rng])
        est_gate_eigenvalues[idx[rng]] = (Ws[g] * pr)[2:end]
        append!(unproj_probs, qs[g]); append!(probs, pr)
    end
    return (est_gate_eigenvalues, unproj_probs, probs)
end

"""
`full_project_gate_eigenvalues` (src/estimate.jl:532-573), the simultaneous projection of all gates (the
reference's default below 5000 gate eigenvalues), for the estimator of the LAST `fit(dd, ls_type, ...)`: the
precision matrix is rebuilt on the device from that fit's Gram matrix, transformed to the space of unpadded gate
probabilities and the N-variable non-negative quadratic program solved exactly by block principal pivoting
(aces_design_project_full) where the reference iterates SCS to `precision`.
"""
function full_project_gate_eigenvalues(dd::DeviceDesign, ls_type::Symbol, est_unproj_gate_eigenvalues::Vector{Float64},
                                       gate_data; precision::Real = 1e-8)
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : 2
    ptr, idx = gate_groups(gate_data)
    G = length(ptr) - 1
    sizes = diff(ptr)
    v = Vector{Float64}(undef, length(idx)); T = Vector{Float64}(undef, sum(sizes .^ 2))
    qs = Vector{Vector{Float64}}(undef, G); Ws = Vector{Matrix{Float64}}(undef, G)
    boff = 0
    for g in 1:G
        k = Int(sizes[g]); rng = (ptr[g] + 1):ptr[g + 1]
        W = wht_block(k)
        q = (W ./ (k + 1)) * vcat(1.0, est_unproj_gate_eigenvalues[idx[rng]])
        Tg = W[2:end, 2:end] .- W[2:end, 1]
        T[(boff + 1):(boff + k * k)] = vec(permutedims(Tg))        # row-major block
        v[idx[rng]] = q[2:end]
        qs[g] = q; Ws[g] = W
        boff += k * k
    end
    y = Vector{Float64}(undef, length(idx)); its = Ref{Int32}(0)
    GC.@preserve ptr idx est_unproj_gate_eigenvalues T v y begin
        check(ccall((:aces_design_project_full, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Float64, Ptr{Float64}, Ptr{Int32}),
            CTX[], dd.handle, Int32(kind), Int32(G), ptr, idx, Int32(1), est_unproj_gate_eigenvalues, T, v,
            Float64(precision), y, its))
    end
    est_gate_eigenvalues = copy(est_unproj_gate_eigenvalues)
    unproj_probs = Float64[]; probs = Float64[]
    for g in 1:G
        rng = (ptr[g] + 1):ptr[g + 1]
        yg = y[idx[rng]]
        pr = vcat(1.0 - sum(yg), yg)
        est_gate_eigenvalues[idx[rng]] = (Ws[g] * pr)[2:end]
        append!(unproj_probs, qs[g]); append!(probs, pr)
    end
    return (est_gate_eigenvalues, unproj_probs, probs)
end

# ---------------------------------------------------------------------------------------------
# estimate_gate_noise (src/estimate.jl:691-941)
# ---------------------------------------------------------------------------------------------

"The ten per-estimator fields of NoiseEstimate (src/estimate.jl:42-76), in the struct's order."
function estimator_fields(unproj::Vector{Float64}, proj::Vector{Float64}, unproj_probs_vec::Vector{Float64},
                          probs_vec::Vector{Float64}, gate_data)
    Q = QuantumACES
    unproj_probs = Q.get_gate_probabilities_dict(unproj_probs_vec, gate_data)
    probs = Q.get_gate_probabilities_dict(probs_vec, gate_data)
    return (unproj, proj, unproj_probs, probs,
        Q.get_marginal_gate_eigenvalues(unproj, gate_data), Q.get_marginal_gate_eigenvalues(proj, gate_data),
        Q.get_marginal_gate_probabilities(unproj_probs), Q.get_marginal_gate_probabilities(probs),
        Q.get_relative_gate_eigenvalues(unproj, gate_data), Q.get_relative_gate_eigenvalues(proj, gate_data))
end

"""
Drop-in for `estimate_gate_noise(d, est_eigenvalues_experiment_ensemble, est_covariance_experiment_ensemble,
shot_budget; ...)` (src/estimate.jl:691-941).  Same decisions in the same order -- OLS and WLS first, the
Euclidean projection of OLS decides whether WLS / GLS are projected in the Mahalanobis distance (:744),
GLS only for a full-covariance design (:820) -- with the numeric work on the GPU: one
`aces_design_fit` per estimator on the cached device-resident Design (covariance, log-covariance,
clipping, block factors with the not-PD fallback, whitened least squares), the projections through
aces_project_simplex / aces_design_precision_blocks / aces_project_nonnegative_batched.  With
`split = false` the simultaneous projection of all gates runs on the GPU as well (full_project_gate_eigenvalues:
aces_design_project_full on the precision matrix of the fit's own Gram matrix).
"""
function estimate_gate_noise(d::Design, est_eigenvalues_experiment_ensemble::Vector{Vector{Vector{Float64}}},
                             est_covariance_experiment_ensemble::Vector{Vector{Vector{Float64}}}, shot_budget::Int;
                             min_eigenvalue::Float64 = 0.1, clip_warning::Bool = true, N_warn_split::Integer = 5 * 10^3,
                             split::Bool = (d.c.gate_data.N < N_warn_split ? false : true), split_warning::Bool = true,
                             precision::Real = 1e-8)
    Q = QuantumACES
    gate_data = d.c.gate_data
    if split_warning && gate_data.N >= N_warn_split && ~split
        @warn "This circuit has a large number of gate eigenvalues: splitting the gate eigenvalue projection across gates by setting `split` to be `true` is advised."
    end
    dd = device_design(d)
    tuple_number = length(d.tuple_set)
    est_eigenvalues = vcat([mean.(est_eigenvalues_experiment_ensemble[i]) for i in 1:tuple_number]...)
    est_cov_eigenvalues = vcat([mean.(est_covariance_experiment_ensemble[i]) for i in 1:tuple_number]...)
    clipped_indices = Q.get_clipped_indices(est_eigenvalues; min_eigenvalue = min_eigenvalue, warning = clip_warning)
    kept = length(clipped_indices) == dd.M ? nothing : clipped_indices
    project(ls_type, unproj, precision_project) = begin
        if ~precision_project
            simple_project_gate_eigenvalues(unproj, gate_data)
        elseif split
            split_project_gate_eigenvalues(dd, ls_type, unproj, gate_data; precision = precision)
        else
            full_project_gate_eigenvalues(dd, ls_type, unproj, gate_data; precision = precision)
        end
    end
    ols_unproj = fit(dd, :ols, est_eigenvalues, est_cov_eigenvalues; clipped_indices = kept)
    ols = simple_project_gate_eigenvalues(ols_unproj, gate_data)
    precision_project = ~(ols[1] ≈ ols_unproj)
    ols_fields = estimator_fields(ols_unproj, ols[1], ols[2], ols[3], gate_data)
    wls_unproj = fit(dd, :wls, est_eigenvalues, est_cov_eigenvalues; clipped_indices = kept)
    wls = project(:wls, wls_unproj, precision_project)      # directly after the WLS fit: the slices are its Gram matrix
    wls_fields = estimator_fields(wls_unproj, wls[1], wls[2], wls[3], gate_data)
    if d.full_covariance
        gls_unproj = fit(dd, :gls, est_eigenvalues, est_cov_eigenvalues; clipped_indices = kept)
        gls = project(:gls, gls_unproj, precision_project)
        gls_fields = estimator_fields(gls_unproj, gls[1], gls[2], gls[3], gate_data)
    else
        gls_fields = wls_fields
    end
    return Q.NoiseEstimate(est_eigenvalues_experiment_ensemble, est_covariance_experiment_ensemble, shot_budget,
        gls_fields..., wls_fields..., ols_fields...)
end

# ---------------------------------------------------------------------------------------------
# S3 / S4 with the reference's signatures
# ---------------------------------------------------------------------------------------------

"calc_covariance(d, eigenvalues_ensemble, covariance_ensemble; epsilon, weight_time, warning) (src/merit.jl:206-291)."
function calc_covariance_design(d::Design, eigenvalues_ensemble::Vector{Vector{Float64}},
                                covariance_ensemble::Vector{Vector{Float64}}; epsilon::Real = 1e-10,
                                weight_time::Bool = true, warning::Bool = true)
    if warning && ~d.full_covariance
        @warn "The design does not include the full covariance matrix; only the diagonal will be calculated."
    end
    return calc_covariance(device_design(d), eigenvalues_ensemble, covariance_ensemble; epsilon = epsilon, weight_time = weight_time)
end

"calc_gls / wls / ols_covariance(d, covariance_log) (src/merit.jl:556-629): Symmetric N x N on the host."
ls_covariance_design(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}, ls_type::Symbol) =
    calc_ls_covariance(device_design(d), covariance_log; ls_type = ls_type, want_matrix = true)[1]

# ---------------------------------------------------------------------------------------------
# Shot-weight optimisation (src/optimise_weights.jl) -- SURVEY.md section 8f, row f3
# ---------------------------------------------------------------------------------------------

"""
    merit_grad_log(dd, ls_type, shot_weights, covariance_log_unweighted, gate_transform_matrix)

calc_gls_merit_grad_log / calc_wls_merit_grad_log / calc_ols_merit_grad_log
(src/optimise_weights.jl:103-152, 327-391, 563-606) -> (merit_grad_log, merit) through
aces_design_merit_grad.  Every estimator takes covariance_log_unweighted itself (the GLS method of the
reference takes its block inverse, the OLS method three dense matrices built from it).
"""
function merit_grad_log(dd::DeviceDesign, ls_type::Symbol, shot_weights::Vector{Float64},
                        covariance_log_unweighted::SparseMatrixCSC{Float64, Int},
                        gate_transform_matrix::SparseMatrixCSC{Float64, Int})
    kind = ls_type == :ols ? 0 : ls_type == :wls ? 1 : 2
    nz = [covariance_log_unweighted[dd.cov_rowval[e] + 1, c] for c in 1:dd.M for e in (dd.cov_colptr[c] + 1):dd.cov_colptr[c + 1]]
    tt = SparseMatrixCSC{Float64, Int}(gate_transform_matrix' * gate_transform_matrix)
    tt_colptr = Int64.(tt.colptr); tt_rowval = Int64.(tt.rowval); tt_nzval = Float64.(tt.nzval)
    tau = Float64.(dd.d.tuple_times)
    grad = Vector{Float64}(undef, length(shot_weights))
    merit = Ref{Float64}(0.0); tr = Ref{Float64}(0.0); trsq = Ref{Float64}(0.0)
    GC.@preserve nz shot_weights tau tt_colptr tt_rowval tt_nzval grad begin
        check(ccall((:aces_design_merit_grad, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64},
             Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            CTX[], dd.handle, Int32(kind), nz, shot_weights, tau, Int64(size(gate_transform_matrix, 1)), tt_colptr,
            tt_rowval, tt_nzval, Int32(1), merit, grad, tr, trsq))
    end
    return (grad, merit[])
end

"""
    ls_optimise_weights(d, covariance_log; options)

gls_ / wls_ / ols_optimise_weights (src/optimise_weights.jl:160-321, 398-555, 613-790): the reference's
Nesterov descent on the logarithms of the shot weights, restated once for the three estimators, with the
gradient of every step on the GPU.
"""
function ls_optimise_weights(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}, ls_type::Symbol;
                             options = QuantumACES.OptimOptions())
    @assert options.ls_type == ls_type "Inappropriate least squares optimisation type $(options.ls_type) supplied."
    est_type = options.est_type
    est_weight = options.est_weight
    momentum = options.momentum
    convergence_steps = options.convergence_steps
    convergence_threshold = options.convergence_threshold
    dd = device_design(d)
    tuple_number = length(d.tuple_set)
    shot_weights = QuantumACES.project_simplex(d.shot_weights)
    mapping_lengths = length.(d.mapping_ensemble)
    covariance_log_unweighted = SparseMatrixCSC{Float64, Int}(covariance_log *
        QuantumACES.get_shot_weights_factor_inv(d.shot_weights, d.tuple_times, mapping_lengths))
    gate_eigenvalues_diag = sparse(Diagonal(QuantumACES.get_gate_eigenvalues(d)))
    transforms = (est_type == :sum || est_type == :prod) ?
        [QuantumACES.get_transform(d, :ordinary) * gate_eigenvalues_diag, QuantumACES.get_transform(d, :relative) * gate_eigenvalues_diag] :
        [QuantumACES.get_transform(d, est_type) * gate_eigenvalues_diag]
    function grad_and_merit(w::Vector{Float64})
        res = [merit_grad_log(dd, ls_type, w, covariance_log_unweighted, SparseMatrixCSC{Float64, Int}(tm)) for tm in transforms]
        if est_type == :sum
            return (est_weight * res[1][1] + (1 - est_weight) * res[2][1], est_weight * res[1][2] + (1 - est_weight) * res[2][2])
        elseif est_type == :prod
            merit = res[1][2]^(est_weight) * res[2][2]^(1 - est_weight)
            return (merit * (est_weight * res[1][1] / res[1][2] + (1 - est_weight) * res[2][1] / res[2][2]), merit)
        else
            return res[1]
        end
    end
    stepping = true
    step = 1
    recently_pruned = 0
    recently_zeroed = 0
    scaled_learning_rate = options.learning_rate
    old_shot_weights = deepcopy(shot_weights)
    velocity = zeros(tuple_number)
    merit_descent = Vector{Float64}()
    while stepping
        nesterov_log_shot_weights = -log.(shot_weights) + momentum * velocity
        nesterov_shot_weights = exp.(-nesterov_log_shot_weights) / sum(exp.(-nesterov_log_shot_weights))
        (merit_grad, merit) = grad_and_merit(nesterov_shot_weights)
        push!(merit_descent, merit)
        velocity = momentum * velocity - scaled_learning_rate * merit_grad
        log_shot_weights_update = -log.(shot_weights) + velocity
        shot_weights_update = exp.(-log_shot_weights_update) / sum(exp.(-log_shot_weights_update))
        if (length(merit_descent) >= 2) && (merit_descent[end] > merit_descent[end - 1])
            shot_weights = old_shot_weights
            velocity = zeros(tuple_number)
            if recently_zeroed > 0
                scaled_learning_rate /= options.learning_rate_scale_factor
            end
            recently_pruned += 1
            recently_zeroed = convergence_steps
        else
            old_shot_weights = deepcopy(nesterov_shot_weights)
            shot_weights = shot_weights_update
        end
        if step >= convergence_steps
            merit_converged = all([abs(1 - merit_descent[idx_2] / merit_descent[idx_1]) < convergence_steps * convergence_threshold
                                   for idx_1 in (step - convergence_steps + 1):step for idx_2 in (step - convergence_steps + 1):step]) &&
                              (abs(1 - merit_descent[end - 1] / merit_descent[end]) < convergence_threshold)
        else
            merit_converged = false
        end
        if (merit_converged || step >= options.max_steps) && recently_pruned == 0
            stepping = false
        else
            step += 1
            recently_pruned > 0 && (recently_pruned -= 1)
            recently_zeroed > 0 && (recently_zeroed -= 1)
        end
    end
    shot_weights_factor = QuantumACES.get_shot_weights_factor(shot_weights, d.tuple_times, mapping_lengths)
    covariance_log_new = SparseMatrixCSC{Float64, Int}(covariance_log_unweighted * shot_weights_factor)
    d_new = QuantumACES.Accessors.@set d.shot_weights = QuantumACES.project_simplex(shot_weights)
    d_new = QuantumACES.Accessors.@set d_new.ls_type = ls_type
    return (d_new::Design, covariance_log_new::SparseMatrixCSC{Float64, Int}, merit_descent::Vector{Float64})
end

# ---------------------------------------------------------------------------------------------
# install!: method overwrite inside QuantumACES (SURVEY.md section 8b, option ii)
# ---------------------------------------------------------------------------------------------

# ---------------------------------------------------------------------------------------------
# calc_merit tail (src/merit.jl:925-1006) -- SURVEY.md section 8f, row f2
# ---------------------------------------------------------------------------------------------

"""
    symmetric_eigenvalues(A::Symmetric{Float64, Matrix{Float64}})

LinearAlgebra.eigvals(A) as calc_merit calls it (src/merit.jl:954-986) through
aces_symmetric_eigenvalues: all eigenvalues, ascending.  The library reads the upper triangle of the
buffer it is given (Julia's default uplo); a Symmetric stored in its lower triangle is copied out first.
"""
function symmetric_eigenvalues(A::Symmetric{Float64, Matrix{Float64}})
    n = size(A, 1)
    data = A.uplo == 'U' ? parent(A) : Matrix{Float64}(A)
    w = Vector{Float64}(undef, n)
    GC.@preserve data w begin
        check(ccall((:aces_symmetric_eigenvalues, LIB[]), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
            CTX[], Int64(n), data, Int64(stride(data, 2)), w))
    end
    return w
end

"""
    merit_spectra(dd, covariance_log) -> (spectra, gram_extremes)

The nine spectra of calc_merit in ONE call of aces_design_merit: `spectra[(ls, est)]` for
ls in (:gls, :wls, :ols), est in (:ordinary, :marginal, :relative); the estimator covariances, their
transforms and the tridiagonalisations never leave the device.  gram_extremes = (largest, smallest)
eigenvalue of d.matrix' * d.matrix.
"""
function merit_spectra(dd::DeviceDesign, covariance_log::SparseMatrixCSC{Float64, Int})
    d = dd.d
    nz = [covariance_log[dd.cov_rowval[e] + 1, c] for c in 1:dd.M for e in (dd.cov_colptr[c] + 1):dd.cov_colptr[c + 1]]
    g = QuantumACES.get_gate_eigenvalues(d)
    tm = QuantumACES.get_transform(d, :marginal)
    tr = QuantumACES.get_transform(d, :relative)
    nm = size(tm, 1); nr = size(tr, 1)
    mc = Int64.(tm.colptr); mr = Int64.(tm.rowval); mv = Float64.(tm.nzval)
    rc = Int64.(tr.colptr); rr = Int64.(tr.rowval); rv = Float64.(tr.nzval)
    per = dd.N + nm + nr
    ev = Vector{Float64}(undef, 3 * per)
    ext = Vector{Float64}(undef, 2)
    GC.@preserve nz g mc mr mv rc rr rv ev ext begin
        check(ccall((:aces_design_merit, LIB[]), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64,
             Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
            CTX[], dd.handle, g, nz, Int64(nm), mc, mr, mv, Int64(nr), rc, rr, rv, Int32(1), ev, ext))
    end
    spectra = Dict{Tuple{Symbol, Symbol}, Vector{Float64}}()
    for (k, ls) in enumerate((:gls, :wls, :ols))
        base = (k - 1) * per
        spectra[(ls, :ordinary)] = ev[(base + 1):(base + dd.N)]
        spectra[(ls, :marginal)] = ev[(base + dd.N + 1):(base + dd.N + nm)]
        spectra[(ls, :relative)] = ev[(base + dd.N + nm + 1):(base + per)]
    end
    return (spectra, (ext[1], ext[2]))
end

"""
    calc_merit(d::Design; warning = true)

calc_merit(d) (src/merit.jl:925-1043) with its numeric core on the GPU; the `Merit` object is assembled
exactly as the reference does (same field order, nrmse_moments of the spectra on the host).  A design
without full covariance data goes through get_full_design like the reference, and its GLS fields are
the WLS ones (src/merit.jl:934-938).
"""
function calc_merit(d::Design; warning::Bool = true)
    d_cov = d
    if ~d.full_covariance
        if warning
            println("The design does not have full covariance matrix data. Generating this data.\nBE CAREFUL: this may consume extreme amounts of memory and be extremely slow.")
        end
        d_cov = QuantumACES.get_full_design(d)
    end
    covariance_log = QuantumACES.calc_covariance_log(d_cov)
    (spectra, (largest, smallest)) = merit_spectra(device_design(d_cov), covariance_log)
    if ~d.full_covariance
        for est in (:ordinary, :marginal, :relative)
            spectra[(:gls, est)] = deepcopy(spectra[(:wls, est)])
        end
    end
    fields = Any[]
    for ls in (:gls, :wls, :ols), est in (:ordinary, :marginal, :relative)
        ev = spectra[(ls, est)]
        (expectation, variance) = QuantumACES.nrmse_moments(ev)
        push!(fields, expectation, variance, ev)
    end
    largest_singular_value = sqrt(largest)
    smallest_singular_value = sqrt(smallest)
    design_matrix_cond_num = largest_singular_value / smallest_singular_value
    design_matrix_pinv_norm = 1 / smallest_singular_value
    return QuantumACES.Merit(
        d.c.circuit_param, d.c.noise_param, d.tuple_set_data, d.tuple_times, d.shot_weights, d.experiment_numbers,
        d.experiment_number, d.c.gate_data.N, d.c.gate_data.N_marginal, d.c.gate_data.N_relative,
        length(d.c.total_gates), fields..., design_matrix_cond_num, design_matrix_pinv_norm)
end

"""
    install!()

Overwrites, inside QuantumACES, the functions of the estimation hot path listed at the top of this file;
every caller (simulate_aces, estimate_stim_ensemble, calc_merit, the optimisers, the tests, user
scripts) then runs through libaces_b200 unchanged.  The Dict{String,Int} method of
estimate_experiment (Qiskit counts, src/estimate.jl:255-277) and
sparse_covariance_inv_factor(covariance_log, mapping_lengths) are left alone.
"""
function install!()
    CTX[] == C_NULL && init!()
    @eval QuantumACES begin
        # K1: sampling loop with the estimation half on the GPU (src/simulate.jl:84-252)
        simulate_stim_estimate(d::Design, shots_set::Vector{Int}; seed::Union{UInt64, Nothing} = nothing,
            max_samples::Integer = 10^9, detailed_diagnostics::Bool = false) =
            $(simulate_stim_estimate)(d, shots_set; seed = seed, max_samples = max_samples,
                detailed_diagnostics = detailed_diagnostics)
        # second caller of S1: one launch per job instead of two estimate_experiment calls per circuit
        estimate_stim_ensemble(d_rand::RandDesign, experiment_shots::Integer; fail_jobs::Vector{Int} = Int[],
            simulation_seed::Union{UInt64, Nothing} = nothing, min_eigenvalue::Real = 0.1, split::Bool = false,
            precision::Real = 1e-8) =
            $(estimate_stim_ensemble)(d_rand, experiment_shots; fail_jobs = fail_jobs, simulation_seed = simulation_seed,
                min_eigenvalue = min_eigenvalue, split = split, precision = precision)
        # S1 (remaining caller: estimate_qiskit_ensemble on parsed counts, src/device.jl:295-313)
        estimate_experiment(counts::Matrix{UInt8}, experiment_meas_indices::Vector{Vector{Int}},
            experiment_pauli_signs::Vector{Bool}, shots_value::Integer) =
            $(estimate_experiment)(counts, experiment_meas_indices, experiment_pauli_signs, shots_value)
        # numeric core of the noise estimate (src/estimate.jl:691-941)
        estimate_gate_noise(d::Design, est_eigenvalues_experiment_ensemble::Vector{Vector{Vector{Float64}}},
            est_covariance_experiment_ensemble::Vector{Vector{Vector{Float64}}}, shot_budget::Int;
            min_eigenvalue::Float64 = 0.1, clip_warning::Bool = true, N_warn_split::Integer = 5 * 10^3,
            split::Bool = (d.c.gate_data.N < N_warn_split ? false : true), split_warning::Bool = true,
            precision::Real = 1e-8) =
            $(estimate_gate_noise)(d, est_eigenvalues_experiment_ensemble, est_covariance_experiment_ensemble, shot_budget;
                min_eigenvalue = min_eigenvalue, clip_warning = clip_warning, N_warn_split = N_warn_split, split = split,
                split_warning = split_warning, precision = precision)
        # S2
        estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64, Int},
            est_covariance_log_inv_factor::SparseMatrixCSC{Float64, Int}, est_eigenvalues::Vector{Float64};
            constrain::Bool = false) =
            $(estimate_gate_eigenvalues)(design_matrix, est_covariance_log_inv_factor, est_eigenvalues;
                constrain = constrain)
        estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64, Int}, est_eigenvalues::Vector{Float64};
            constrain::Bool = false) =
            $(estimate_gate_eigenvalues)(design_matrix, nothing, est_eigenvalues; constrain = constrain)
        # S3
        calc_covariance(d::Design, eigenvalues_ensemble::Vector{Vector{Float64}},
            covariance_ensemble::Vector{Vector{Float64}}; epsilon::Real = 1e-10, weight_time::Bool = true,
            warning::Bool = true) =
            $(calc_covariance_design)(d, eigenvalues_ensemble, covariance_ensemble; epsilon = epsilon,
                weight_time = weight_time, warning = warning)
        sparse_covariance_inv_factor(d::Design, covariance_log::SparseMatrixCSC{Float64, Int};
            clipped_indices::Union{Vector{Int}, Nothing} = nothing, epsilon::Real = 1e-12) =
            $(sparse_covariance_inv_factor)($(device_design)(d), covariance_log; clipped_indices = clipped_indices,
                epsilon = epsilon)
        # S4
        calc_gls_covariance(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}) = $(ls_covariance_design)(d, covariance_log, :gls)
        calc_wls_covariance(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}) = $(ls_covariance_design)(d, covariance_log, :wls)
        calc_ols_covariance(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}) = $(ls_covariance_design)(d, covariance_log, :ols)
        # f2: calc_merit with the covariances, transforms and nine spectra on the GPU (src/merit.jl:925-1043)
        calc_merit(d::Design; warning::Bool = true) = $(calc_merit)(d; warning = warning)
        # f3: the descent loops of the shot-weight optimiser (src/optimise_weights.jl:160-321, 398-555, 613-790)
        gls_optimise_weights(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}; options::OptimOptions = OptimOptions()) =
            $(ls_optimise_weights)(d, covariance_log, :gls; options = options)
        wls_optimise_weights(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}; options::OptimOptions = OptimOptions()) =
            $(ls_optimise_weights)(d, covariance_log, :wls; options = options)
        ols_optimise_weights(d::Design, covariance_log::SparseMatrixCSC{Float64, Int}; options::OptimOptions = OptimOptions()) =
            $(ls_optimise_weights)(d, covariance_log, :ols; options = options)
    end
    return nothing
end

# Only names QuantumACES does not export itself: a user script has `using QuantumACES` in scope, and two
# modules exporting the same name would make every unqualified use ambiguous.  The GPU versions of
# QuantumACES functions are reached through install!() or qualified (ACESB200.fit, ...).
export init!, shutdown!, install!, PauliTables, DeviceDesign, estimate_counts!

end # module
