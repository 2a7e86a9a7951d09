# RBL_b200.jl - drop-in for the reference's RBL_gpu(A,k,b) / RBL(A,k,b) on top of librbl_b200.so (ccall).
#
#   include("RBL_b200.jl")            # instead of include("RBL_gpu.jl")   (Julia/benchmark.jl:8, images.jl)
#   d, v = RBL_gpu(A, 100, 16)        # same positional call, same return values (RBL_gpu.jl:205-221)
#
# The reference's host code is Julia, so this wrapper is the host side of the boundary; everything below
# `ccall` is the C ABI of include/rbl_b200.h.  (No Julia toolchain exists in the build image of this
# repository: this file is exercised through the byte-identical Python ctypes mirror in binding.py / rbl.py,
# which passes the same arrays - Int64 colptr/rowval, Float64 nzval, column-major Float64 blocks.)
using SparseArrays
using LinearAlgebra

const LIBRBL = get(ENV, "RBL_B200_LIB", joinpath(@__DIR__, "..", "lib", "librbl_b200.so"))

# mirrors `rbl_options` (include/rbl_b200.h)
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
    reserved::NTuple{7,Int32}
end

# mirrors `rbl_stats`; only the leading fields are named, the rest is padding up to sizeof(rbl_stats)
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
    rest::NTuple{32,Float64}
end
RblStats() = RblStats(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, ntuple(_ -> 0.0, 32))

function rbl_default_options()
    o = RblOptions(0, 0.0, 0, 0, 0, 0, 0.0, 0, 0, 0, 0, 0, 0, ntuple(_ -> Int32(0), 7))
    ccall((:rbl_options_default, LIBRBL), Cint, (Ref{RblOptions},), o)
    return o
end

rbl_error() = unsafe_string(ccall((:rbl_last_error, LIBRBL), Cstring, ()))

"""
    RBL_gpu(A, k, b; Ω=nothing, max_kryl_sz=1200, tol=1e-7, reorth_period=2, check_period=4,
            precision=:fp64, shift=nothing, device=-1)

k eigenvalues of largest magnitude (descending |λ|, RBL.jl:116) and the n×k Ritz vectors.  Keywords are the
constants the reference hard-codes, with its values as defaults (RBL_gpu.jl:211,189,164,186; common.jl:5-6).
`shift=σ` solves for σI − A (lowest eigenpairs of A as the largest of the shifted operator).
"""
function RBL_gpu(A::Union{SparseMatrixCSC{Float64},Matrix{Float64}}, k::Int64, b::Int64;
                 Ω::Union{Nothing,Matrix{Float64}}=nothing, max_kryl_sz::Int64=1200, tol::Float64=1e-7,
                 reorth_period::Int=2, check_period::Int=4, precision::Symbol=:fp64,
                 shift::Union{Nothing,Float64}=nothing, device::Int=-1, verbose::Int=0)
    n = size(A, 2)
    o = rbl_default_options()
    o.max_kryl_sz = max_kryl_sz; o.tol = tol
    o.reorth_period = reorth_period; o.check_period = check_period
    o.precision = precision == :mixed ? 1 : 0
    o.op = shift === nothing ? 0 : 1
    o.sigma = shift === nothing ? 0.0 : shift
    o.device = device; o.verbose = verbose
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
    V = zeros(Float64, n, k)
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
        # phase seconds are in st.t_spmm ("AQ"), st.t_3term, st.t_qr, st.t_part_reorth, st.t_loc_reorth, st.t_eig, st.t_ritz
    end
    return D, V
end

# RBL(A,k,b) of RBL.jl:119 - same contract, CPU cap of 1400 columns (RBL.jl:133)
RBL(A, k::Int64, b::Int64; kwargs...) = RBL_gpu(A, k, b; max_kryl_sz=1400, kwargs...)
