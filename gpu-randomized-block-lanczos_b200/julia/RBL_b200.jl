# RBL_b200.jl - drop-in for the reference's RBL_gpu(A,k,b) / RBL(A,k,b) on top of librbl_b200.so (ccall).
#
#   include("RBL_b200.jl")            # instead of include("RBL_gpu.jl")   (Julia/benchmark.jl:8, images.jl)
#   d, v = RBL_gpu(A, 100, 16)        # same positional call, same return values (RBL_gpu.jl:205-221)
#   d, v = RBL_gpu(A, 100, 16; ngpus=8, precision=:mixed)     # row-sharded over the GPUs of this process
#
# The reference's host code is Julia, so this wrapper is the host side of the boundary; everything below
# `ccall` is the C ABI of include/rbl_b200.h.  No Julia toolchain exists in the build image of this repository:
# the SAME calls with the SAME arrays (Int64 1-based colptr/rowval with index_base = 1, Float64 nzval,
# column-major Float64 blocks) are executed by the Python ctypes mirror (binding.py / rbl.py) in
# tests/test_gpu_parity_gates.py::test_julia_abi_*; tests/test_host.py checks that the struct field lists below equal
# the header's, field for field.
using SparseArrays
using LinearAlgebra

const LIBRBL = get(ENV, "RBL_B200_LIB", joinpath(@__DIR__, "..", "lib", "librbl_b200.so"))

# mirrors `rbl_options` (include/rbl_b200.h), field for field
mutable struct RblOptions
    max_kryl_sz::Int64
    tol::Float64
    reorth_period::Int32
    check_period::Int32
    precision::Int32
    op::Int32
    sigma::Float64
    device::Int32
    async_check::Int32
    host_threads::Int32
    v_fp32::Int32
    verbose::Int32
    reorth_impl::Int32
    seed::Int32
    ngpus::Int32
    filter_degree::Int32
    restart::Int32
    spill::Int32
    probe_steps::Int32
    mem_limit_mb::Int32
end
RblOptions() = RblOptions(0, 0.0, 0, 0, 0, 0, 0.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)

# mirrors `rbl_stats`, field for field
mutable struct RblStats
    iterations::Int64
    kryl_sz::Int64
    iterations_run::Int64
    converged::Int32
    checks::Int32
    full_checks::Int32
    deflated::Int32
    t_total::Float64
    t_spmm::Float64        # "AQ"            RBL_gpu.jl:152,176
    t_3term::Float64       # "3-term"        :153-154,177-179
    t_qr::Float64          # "qr"            :155,180
    t_part_reorth::Float64 # "part reorth"   :165
    t_loc_reorth::Float64  # "loc reorth"    :167
    t_eig::Float64         # "eig"           :187
    t_ritz::Float64        # "Ritz vectors"  :219
    t_h2d::Float64
    t_d2h::Float64
    t_eig_wait::Float64
    bytes_part_reorth::Float64
    bytes_spmm::Float64
    kernel_launches::Int64
    t_reorth_gram::Float64
    t_reorth_update::Float64
    bytes_reorth_gram::Float64
    bytes_reorth_update::Float64
    launches_reorth_gram::Int64
    launches_reorth_update::Int64
    launches_spmm::Int64
    t_ritz_kernel::Float64
    bytes_ritz::Float64
    flops_ritz::Float64
    host_factorizations::Int64
    restarts::Int64
    locked::Int64
    spilled_blocks::Int64
    buffer_blocks::Int64
    max_residual::Float64
    t_host_blocked::Float64
    filter_cut::Float64
    filter_degree::Int32
    filter_two_sided::Int32
end
RblStats() = RblStats(0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, 0.0,
                      0, 0, 0, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0, 0)

function rbl_default_options()
    o = RblOptions()
    so = Ref{Int64}(0); ss = Ref{Int64}(0)
    ccall((:rbl_struct_sizes, LIBRBL), Cint, (Ref{Int64}, Ref{Int64}), so, ss)
    (so[] == sizeof(RblOptions) && ss[] == sizeof(RblStats)) ||
        error("RBL_b200.jl is out of date with librbl_b200.so: struct sizes $(sizeof(RblOptions))/$(sizeof(RblStats)) vs $(so[])/$(ss[])")
    ccall((:rbl_options_default, LIBRBL), Cint, (Ref{RblOptions},), o)
    return o
end

rbl_error() = unsafe_string(ccall((:rbl_last_error, LIBRBL), Cstring, ()))

# The reference accumulates its phase times in a global TimerOutput `to` (RBL_gpu.jl:152-187,219; test.jl:7 creates
# it).  The device phases are measured with CUDA events inside the library; they are added to `to` under the
# reference's own labels so that `show(to)` after a benchmark prints the familiar table.
function _record_phase!(to, label::String, seconds::Float64, ncalls::Int64)
    t = get!(() -> typeof(to)(label), to.inner_timers, label)
    t.accumulated_data.ncalls += ncalls
    t.accumulated_data.time += round(Int64, seconds * 1e9)
    return nothing
end

"""
    RBL_gpu(A, k, b; Ω=nothing, max_kryl_sz=1200, tol=1e-7, reorth_period=2, check_period=4,
            precision=:fp64, shift=nothing, device=-1, ngpus=1, restart=false, filter_degree=0,
            spill=false, seed=0, FLOAT=Float64)

k eigenvalues of largest magnitude (descending |λ|, RBL.jl:116) and the n×k Ritz vectors (`Matrix{FLOAT}`,
RBL_gpu.jl:109,219).  Keywords are the constants the reference hard-codes, with its values as defaults
(RBL_gpu.jl:211,189,164,186; common.jl:5-6).  `shift=σ` solves for σI − A (lowest eigenpairs of A as the largest of
the shifted operator).  `ngpus` row-shards A and every Krylov block over that many GPUs of this process.
`restart=true` locks converged pairs and restarts at the Krylov cap (restarted.jl); `filter_degree=d` iterates with
the degree-d Chebyshev-filtered operator; `spill=true` keeps Krylov blocks that do not fit HBM in pinned host memory
(hybrid_part_reorth!, RBL_gpu.jl:59-81).
"""
function RBL_gpu(A::Union{SparseMatrixCSC{Float64},Matrix{Float64}}, k::Int64, b::Int64;
                 Ω::Union{Nothing,Matrix{Float64}}=nothing, max_kryl_sz::Int64=1200, tol::Float64=1e-7,
                 reorth_period::Int=2, check_period::Int=4, precision::Symbol=:fp64,
                 shift::Union{Nothing,Float64}=nothing, device::Int=-1, verbose::Int=0, ngpus::Int=1,
                 restart::Bool=false, filter_degree::Int=0, spill::Bool=false, seed::Int=0,
                 FLOAT::Type=Float64, stats::Union{Nothing,Ref{RblStats}}=nothing)
    n = size(A, 2)
    o = rbl_default_options()
    o.max_kryl_sz = max_kryl_sz; o.tol = tol
    o.reorth_period = reorth_period; o.check_period = check_period
    o.precision = precision == :mixed ? 1 : 0
    o.op = shift === nothing ? 0 : 1
    o.sigma = shift === nothing ? 0.0 : shift
    o.device = device; o.verbose = verbose
    o.ngpus = ngpus; o.restart = restart ? 1 : 0; o.filter_degree = filter_degree; o.spill = spill ? 1 : 0
    o.seed = seed
    o.v_fp32 = FLOAT == Float32 ? 1 : 0
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if A isa SparseMatrixCSC
        # symmetric A: its CSC arrays are the CSR arrays; Julia's 1-based Int64 indices go in unchanged
        rc = ccall((:rbl_create, LIBRBL), Cint,
                   (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ref{RblOptions}, Ref{Ptr{Cvoid}}),
                   n, nnz(A), A.colptr, A.rowval, A.nzval, 1, o, h)
    else
        rc = ccall((:rbl_create_dense, LIBRBL), Cint, (Int64, Ptr{Float64}, Ref{RblOptions}, Ref{Ptr{Cvoid}}),
                   n, A, o, h)
    end
    rc == 0 || error("rbl_create: status $rc: $(rbl_error())")
    D = zeros(Float64, k)
    V = zeros(FLOAT, n, k)
    st = RblStats()
    try
        rc = ccall((:rbl_solve, LIBRBL), Cint,
                   (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}, Ref{RblStats}),
                   h[], k, b, Ω === nothing ? C_NULL : Ω, D, V, st)
    finally
        ccall((:rbl_destroy, LIBRBL), Cint, (Ptr{Cvoid},), h[])
    end
    if rc == 1
        @warn "RBL_gpu: Krylov cap reached before convergence; returning the Ritz pairs of the last check"
    elseif rc != 0
        error("rbl_solve: status $rc: $(rbl_error())")
    end
    println("Iterations: $(st.iterations) and kryl_sz: $(st.kryl_sz)")   # RBL_gpu.jl:195
    if @isdefined(to)   # the reference needs a global TimerOutput `to` (RBL_gpu.jl:152); optional here
        its = st.iterations_run
        _record_phase!(to, "AQ", st.t_spmm, st.launches_spmm)                       # :152,176
        _record_phase!(to, "3-term", st.t_3term, its)                               # :153-154,177-179
        _record_phase!(to, "qr", st.t_qr, its)                                      # :155,180
        _record_phase!(to, "part reorth", st.t_part_reorth, st.launches_reorth_gram) # :165
        _record_phase!(to, "loc reorth", st.t_loc_reorth, its)                      # :167
        _record_phase!(to, "eig", st.t_eig, Int64(st.checks))                       # :187
        _record_phase!(to, "Ritz vectors", st.t_ritz, Int64(1))                     # :219
    end
    stats === nothing || (stats[] = st)
    return D, V
end

# RBL(A,k,b) of RBL.jl:119 - same contract, CPU cap of 1400 columns (RBL.jl:133)
RBL(A, k::Int64, b::Int64; kwargs...) = RBL_gpu(A, k, b; max_kryl_sz=1400, kwargs...)

# RBL_gpu_restarted(A,k) / RBL_restarted(A,k) of restarted.jl:98-146,196-246: b = 1, short cycles, locking.  Unlike the
# reference (whose V is never filled, restarted.jl:100,145) the Ritz vectors are returned.
RBL_gpu_restarted(A, k::Int64; max_kryl_sz::Int64=100, kwargs...) =
    RBL_gpu(A, k, 1; max_kryl_sz=max_kryl_sz, restart=true, kwargs...)
RBL_restarted(A, k::Int64; max_kryl_sz::Int64=80, kwargs...) =
    RBL_gpu(A, k, 1; max_kryl_sz=max_kryl_sz, restart=true, kwargs...)
