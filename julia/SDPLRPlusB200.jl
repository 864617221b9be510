# SDPLRPlusB200.jl -- thin `ccall` shim that plugs libsdplrp_b200.so into SDPLRPlus.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither the build container nor the GPU box has a Julia
# toolchain (DESIGN.md section 1).  The same control flow, through the same C entry points, is
# exercised by sdplrplus.jl_b200/solver.py (ctypes) in tests/ and bench.py.
#
# What it does: `_sdplr` (src/sdplr.jl:140-449) is duck-typed over (data, var, aux) and reaches the
# hot path through the functions listed in SURVEY.md 8b -- the same seam src/lowrankopt.jl:47-135
# overloads for LowRankOpt models.  This file adds a device-backed `aux` (B200Auxiliary), a device
# matrix wrapper usable as the `TR` type parameter of SolverVars / LBFGSVector (src/structs.jl:194,
# src/lbfgs.jl:4), and methods of the seam functions for them.  The user-facing API is untouched:
#
#     using SDPLRPlus, SDPLRPlusB200
#     res = SDPLRPlusB200.sdplr(C, As, b, r; ptol = 1e-2, objtol = 1e-2, prior_trace_bound = n)
#
# (identical signature and result Dict to SDPLRPlus.sdplr, src/sdplr.jl:91-138, 426-448).
module SDPLRPlusB200

using LinearAlgebra, SparseArrays
import SDPLRPlus
import SDPLRPlus: SDPData, SolverVars, SolverStats, BurerMonteiroConfig, SymLowRankMatrix,
    f!, g!, fg!, 𝒜!, 𝒜t!, 𝒜t_preprocess!, linesearch!, linesearch_armijo!, lbfgs_dir!, lbfgs_update!,
    lbfgs_clear!, lbfgs_init, dual_obj, approx_mineigval_lanczos, side_dimension, b_vector, C_matrix,
    LBFGSHistory, _sdplr

const LIB = get(ENV, "SDPLRP_B200_LIB", joinpath(@__DIR__, "..", "sdplrplus.jl_b200", "libsdplrp_b200.so"))

# ids of include/sdplrp_b200.h
const MAT_R, MAT_G, MAT_D = Cint(0), Cint(1), Cint(2)
const VEC_LAMBDA, VEC_LAMBDA_UB, VEC_B, VEC_PVIO_RAW, VEC_Y, VEC_PVIO_LB, VEC_A_RD, VEC_A_DD = Cint.(0:7)

struct B200Error <: Exception
    code::Int32
    msg::String
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:sdplrp_create, LIB), Int32, (Int32, Int32, Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, 0, 1, C_NULL, out)
        rc == 0 || throw(B200Error(rc, unsafe_string(ccall((:sdplrp_error_string, LIB), Cstring, (Int32,), rc))))
        h = new(out[])
        finalizer(x -> ccall((:sdplrp_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

@inline function check(h::Handle, rc::Int32)
    rc == 0 && return nothing
    throw(B200Error(rc, unsafe_string(ccall((:sdplrp_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))))
end

"""
Device-backed replacement of `SolverAuxiliary` (src/structs.jl:274-361): the aggregated pattern and every
index map of `preprocess_sparsecons` are built on the GPU by `sdplrp_preprocess`.
"""
struct B200Auxiliary
    h::Handle
    n::Int
    m::Int
end
side_dimension(aux::B200Auxiliary) = aux.n   # src/structs.jl:363

"""
Concatenate the sparse matrices in `findnz` order exactly as `SolverAuxiliary` walks them
(src/structs.jl:303-332): sparse / Diagonal `A_i` in order of appearance, then `C` if sparse;
`SymLowRankMatrix` constraints are registered separately.
"""
function B200Auxiliary(data::SDPData{Ti,Tv}; device=0) where {Ti,Tv}
    h = Handle(device)
    I, J, V, off, gids = Int64[], Int64[], Float64[], Int64[0], Int64[]
    lowrank = Tuple{Int,SymLowRankMatrix{Tv}}[]
    add!(A, gid) = begin
        A isa Diagonal && (A = sparse(A))
        if A isa SymLowRankMatrix
            push!(lowrank, (gid, A))
        else
            i, j, v = findnz(A)      # CSC: column-major; COO: stored order
            append!(I, i); append!(J, j); append!(V, v)
            push!(off, length(I)); push!(gids, gid)
        end
    end
    for (i, A) in enumerate(data.As); add!(A, i); end
    add!(data.C, data.m + 1)
    GC.@preserve off I J V gids check(h, ccall((:sdplrp_preprocess, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}),
        h.ptr, data.n, data.m, length(gids), off, I, J, V, gids))
    for (gid, A) in lowrank
        B = Matrix(A.B); D = Vector(A.D.diag)
        GC.@preserve B D check(h, ccall((:sdplrp_add_symlowrank, LIB), Int32,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}), h.ptr, gid, size(B, 2), B, D))
    end
    ineq = UInt8.(data.constraint_types)
    GC.@preserve ineq check(h, ccall((:sdplrp_set_problem, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}),
        h.ptr, data.b, data.has_inequalities ? pointer(ineq) : Ptr{UInt8}(C_NULL)))
    return B200Auxiliary(h, data.n, data.m)
end

"""
`DevMat` stands for an r x n matrix that lives on the device (`id` = SDPLRP_MAT_*).  It is the `TR` of
`SolverVars{Ti,Tv,TR}` and the `Ts` of `LBFGSVector{T,Ts}`; the BLAS-1 calls `_sdplr` makes on such arrays
(`dot`, `axpy!`, `norm`, `BLAS.scal!`, `copyto!`) never move data: the fused entry points below do the work.
"""
struct DevMat <: AbstractMatrix{Float64}
    h::Handle
    id::Cint
    r::Int
    n::Int
end
Base.size(A::DevMat) = (A.r, A.n)
Base.getindex(A::DevMat, i::Int, j::Int) = Array(A)[i, j]     # debugging only: downloads the matrix
function Base.Array(A::DevMat)
    out = Matrix{Float64}(undef, A.r, A.n)
    check(A.h, ccall((:sdplrp_download_mat, LIB), Int32, (Ptr{Cvoid}, Cint, Ptr{Float64}), A.h.ptr, A.id, out))
    return out
end
upload!(A::DevMat, X::Matrix{Float64}) =
    check(A.h, ccall((:sdplrp_upload_mat, LIB), Int32, (Ptr{Cvoid}, Cint, Ptr{Float64}), A.h.ptr, A.id, X))

"""
SolverVars whose Rt/Gt are device matrices; the m-vectors stay host `Vector`s that mirror device state only
when Julia needs them (λ for the result Dict, y for `best_λ`).
"""
function device_vars(data::SDPData, aux::B200Auxiliary, r, config::BurerMonteiroConfig)
    host = SolverVars(data, r, config)                       # draws Rt0 / applies init_func exactly as the reference
    h = aux.h
    check(h, ccall((:sdplrp_set_rank, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), h.ptr, r, config.numlbfgsvecs))
    Rt, Gt = DevMat(h, MAT_R, r, data.n), DevMat(h, MAT_G, r, data.n)
    upload!(Rt, Matrix(host.Rt))
    check(h, ccall((:sdplrp_upload_vec, LIB), Int32, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), h.ptr, VEC_LAMBDA, host.λ, data.m))
    check(h, ccall((:sdplrp_set_sigma, LIB), Int32, (Ptr{Cvoid}, Float64), h.ptr, config.σ_0))
    var = SolverVars(Rt, Gt, host.λ, host.λ_ub, host.r, host.σ, host.obj, host.y, host.primal_vio_raw,
        host.primal_vio_lb, host.primal_vio, host.A_RD, host.A_DD)
    return var, host.Rt
end

# ---- seam functions (SURVEY.md 8b) for the device types -------------------------------------------------
sync_sigma(var, aux) = check(aux.h, ccall((:sdplrp_set_sigma, LIB), Int32, (Ptr{Cvoid}, Float64), aux.h.ptr, var.σ[]))

function fg!(data, var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary, normC, normb) where {Ti,Tv}   # src/coreop.jl:323-349
    sync_sigma(var, aux)
    out = zeros(4)
    check(aux.h, ccall((:sdplrp_fg, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), aux.h.ptr, out))
    var.obj[] = out[2]
    return out[1], sqrt(out[3]) / normC, sqrt(out[4]) / normb
end

function g!(var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary) where {Ti,Tv}                        # src/coreop.jl:305-317
    gn2, pn2 = Ref(0.0), Ref(0.0)
    check(aux.h, ccall((:sdplrp_g, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), aux.h.ptr, gn2, pn2))
    return gn2[], pn2[]
end
# `norm(var.Gt)` and the capped-residual norm of src/sdplr.jl:224-234 are the two values g! returned:
LinearAlgebra.norm(G::DevMat, p::Real=2) = sqrt(last_gnorm2[])
const last_gnorm2, last_pnorm2 = Ref(0.0), Ref(0.0)

function lbfgs_dir!(dirt::DevMat, his, Gt::DevMat; negate::Bool=true)                               # src/lbfgs.jl:77-124
    d = Ref(0.0)
    check(dirt.h, ccall((:sdplrp_lbfgs_dir, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}), dirt.h.ptr, d))
    last_descent[] = d[]
    return nothing
end
const last_descent = Ref(0.0)
LinearAlgebra.dot(dirt::DevMat, Gt::DevMat) = last_descent[]                                        # src/sdplr.jl:201
# the non-descent fallback `Gt .*= -1; copyto!(dirt, Gt)` (src/sdplr.jl:202-205):
use_gradient_direction!(aux) = check(aux.h, ccall((:sdplrp_use_gradient_direction, LIB), Int32, (Ptr{Cvoid},), aux.h.ptr))

function linesearch!(var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary, dirt::DevMat; α_max=1.0) where {Ti,Tv}  # src/linesearch.jl:4-127
    biquadratic = zeros(5)
    check(aux.h, ccall((:sdplrp_linesearch_coeffs, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), aux.h.ptr, biquadratic))
    α, 𝓛 = SDPLRPlus.cubic_linesearch_from_coeffs(biquadratic, α_max)   # root selection stays in Julia (src/linesearch.jl:58-112)
    obj = Ref(0.0)
    # primal_vio_raw += α(α A_DD + A_RD), obj, and Rt += α*dirt (src/linesearch.jl:118-124, src/sdplr.jl:219)
    check(aux.h, ccall((:sdplrp_step, LIB), Int32, (Ptr{Cvoid}, Float64, Ref{Float64}), aux.h.ptr, α, obj))
    var.obj[] = obj[]
    return α, 𝓛
end
LinearAlgebra.axpy!(α, dirt::DevMat, Rt::DevMat) = Rt      # already applied by sdplrp_step

lbfgs_update!(dirt::DevMat, his, Gt::DevMat, α) =                                                   # src/lbfgs.jl:129-149
    check(dirt.h, ccall((:sdplrp_lbfgs_update, LIB), Int32, (Ptr{Cvoid}, Float64), dirt.h.ptr, α))
lbfgs_clear!(his::LBFGSHistory{<:Any,<:Any,DevMat}) =                                               # src/lbfgs.jl:52-59
    check(his.vecs[1].s.h, ccall((:sdplrp_lbfgs_clear, LIB), Int32, (Ptr{Cvoid},), his.vecs[1].s.h.ptr))

function dual_obj(data, var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary, trace_bound, iter; highprecision=false) where {Ti,Tv}  # src/coreop.jl:376-415
    v0 = randn(data.n)                                        # src/coreop.jl:473: the start vector stays Julia's
    dual, lam, steps = Ref(0.0), Ref(0.0), Ref{Int64}(0)
    GC.@preserve v0 check(aux.h, ccall((:sdplrp_dual_obj, LIB), Int32,
        (Ptr{Cvoid}, Float64, Int64, Ptr{Float64}, UInt64, Ref{Float64}, Ref{Float64}, Ref{Int64}),
        aux.h.ptr, trace_bound, iter, v0, 0, dual, lam, steps))
    return dual[]
end

# λ_i <- min(ub_i, λ_i − σ v_i) (src/sdplr.jl:358-362) happens on the device; the host copy is refreshed for `best_λ`
dual_update!(var, aux) = begin
    check(aux.h, ccall((:sdplrp_dual_update, LIB), Int32, (Ptr{Cvoid},), aux.h.ptr))
    check(aux.h, ccall((:sdplrp_download_vec, LIB), Int32, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), aux.h.ptr, VEC_LAMBDA, var.λ, length(var.λ)))
end

"""
    sdplr(C, As, b, r; kwargs...)

Same contract as `SDPLRPlus.sdplr` (src/sdplr.jl:91-138); the hot path runs on the B200.
"""
function sdplr(C, As, b, r; device=0, kwargs...)
    config = BurerMonteiroConfig()
    for (k, v) in kwargs
        hasfield(BurerMonteiroConfig, k) ? setfield!(config, k, v) : @error "Unrecognized keyword argument $k"
    end
    data = SDPData(C, As, b)
    aux = B200Auxiliary(data; device)
    var, Rt0 = device_vars(data, aux, r, config)
    ans = _sdplr(data, var, aux, SolverStats{Float64}(), config)
    ans["Rt"] = Array(var.Rt); ans["Rt0"] = Rt0
    return ans
end

end # module
