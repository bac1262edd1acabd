# ACESB200.jl -- Julia shim that binds libaces_b200.so (include/aces_b200.h) with `ccall` and plugs it
# into QuantumACES.jl's estimation hot path without changing user code.
#
# STATUS: written against the C ABI and the reference's Julia signatures, NOT executed: there is
# no `julia` binary in the build image or on the GPU box (DESIGN.md section 1).  The Python/ctypes
# twin (quantumaces.jl_b200/_lib.py, merit.py, estimate.py) binds the same entry points with the
# same argument meaning and is what the test-suite exercises.
#
# Usage
#     using QuantumACES
#     include("ACESB200.jl"); using .ACESB200
#     ACESB200.init!(; device = 0, lib = "/path/to/libaces_b200.so")
#     ACESB200.install!()        # overwrite the hot-path methods inside QuantumACES (seams S1-S3)
#     aces_data = simulate_aces(d, budget_set; repetitions = 20)      # unchanged user code
#
# What is replaced (file:line of the reference):
#   S1  estimate_experiment(::Matrix{UInt8}, ...)          src/estimate.jl:284-301
#   S2  estimate_gate_eigenvalues(A, F, est; constrain)     src/estimate.jl:492-524 (and the OLS form)
#   S3  sparse_covariance_inv_factor(cov_log, lengths)      src/merit.jl:378-427
# plus a design-resident fast path (`DeviceDesign`, `fit`, `calc_ls_covariance_moments`) for callers
# that fit the same Design many times (simulate_aces repetitions, calc_merit, scaling sweeps).
# Everything else (Stim sampling, projections, NoiseEstimate bookkeeping, I/O) stays in QuantumACES.
module ACESB200

using SparseArrays, LinearAlgebra
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
    ccall((:aces_destroy, LIB[]), Int32, (Ptr{Cvoid},), CTX[])
    CTX[] = C_NULL
    return nothing
end

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
    finalizer(x -> ccall((:aces_tables_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.handle), t)
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
    finalizer(x -> ccall((:aces_design_destroy, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.handle), dd)
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
                           constrain::Bool = false)
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
            check(ccall((:aces_design_fit_solve, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], dd.handle))
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
# install!: method overwrite of the seams inside QuantumACES (SURVEY.md section 8b, option ii)
# ---------------------------------------------------------------------------------------------

"""
    install!()

Overwrites, inside QuantumACES, the three pure functions on the hot path with the GPU versions; every
caller (simulate_aces, estimate_gate_noise, estimate_stim_ensemble, calc_merit, the tests) then runs
through libaces_b200 unchanged.  The Dict{String,Int} method of estimate_experiment (Qiskit counts,
src/estimate.jl:255-277) is left alone.
"""
function install!()
    CTX[] == C_NULL && init!()
    @eval QuantumACES begin
        estimate_experiment(counts::Matrix{UInt8}, experiment_meas_indices::Vector{Vector{Int}},
            experiment_pauli_signs::Vector{Bool}, shots_value::Integer) =
            $(estimate_experiment)(counts, experiment_meas_indices, experiment_pauli_signs, shots_value)
        estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64, Int},
            est_covariance_log_inv_factor::SparseMatrixCSC{Float64, Int}, est_eigenvalues::Vector{Float64};
            constrain::Bool = false) =
            $(estimate_gate_eigenvalues)(design_matrix, est_covariance_log_inv_factor, est_eigenvalues;
                constrain = constrain)
        estimate_gate_eigenvalues(design_matrix::SparseMatrixCSC{Float64, Int}, est_eigenvalues::Vector{Float64};
            constrain::Bool = false) =
            $(estimate_gate_eigenvalues)(design_matrix, nothing, est_eigenvalues; constrain = constrain)
    end
    return nothing
end

export init!, shutdown!, install!, PauliTables, estimate_counts!, DeviceDesign, fit, calc_covariance,
    sparse_covariance_inv_factor, calc_ls_covariance

end # module
