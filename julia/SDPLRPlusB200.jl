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
    if highprecision   # SDP_S_eigval (GenericArpack symeigs, src/coreop.jl:386-400) -> thick-restart Lanczos on the device
        GC.@preserve v0 check(aux.h, ccall((:sdplrp_dual_obj_highprecision, LIB), Int32,
            (Ptr{Cvoid}, Float64, Ptr{Float64}, UInt64, Ref{Float64}, Ref{Float64}, Ref{Int64}),
            aux.h.ptr, trace_bound, v0, 0, dual, lam, steps))
    else
        GC.@preserve v0 check(aux.h, ccall((:sdplrp_dual_obj, LIB), Int32,
            (Ptr{Cvoid}, Float64, Int64, Ptr{Float64}, UInt64, Ref{Float64}, Ref{Float64}, Ref{Int64}),
            aux.h.ptr, trace_bound, iter, v0, 0, dual, lam, steps))
    end
    return dual[], lam[]
end

"""
`SDP_S_eigval(var, aux, nevs, true; which=:SA, ncv, tol, maxiter)` (src/coreop.jl:351-374) on the S last assembled.
"""
function SDPLRPlus.SDP_S_eigval(var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary, nevs, preprocessed::Bool=false;
                                ncv=min(100, aux.n), tol=0.0, maxiter=1000000, kwargs...) where {Ti,Tv}
    preprocessed || check(aux.h, ccall((:sdplrp_At_preprocess, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), aux.h.ptr, var.y))
    ev = zeros(nevs); v0 = randn(aux.n)
    dt = @elapsed GC.@preserve ev v0 check(aux.h, ccall((:sdplrp_S_eigval, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Float64, Int64, Ptr{Float64}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
        aux.h.ptr, nevs, ncv, tol, maxiter, v0, 0, ev, C_NULL, C_NULL, C_NULL))
    return ev, dt
end

"""
`DIMACS_errors(data, var, aux)` (src/coreop.jl:426-453) in one call.
"""
function SDPLRPlus.DIMACS_errors(data, var::SolverVars{Ti,Tv,DevMat}, aux::B200Auxiliary) where {Ti,Tv}
    errs = zeros(6); v0 = randn(data.n)
    GC.@preserve errs v0 check(aux.h, ccall((:sdplrp_dimacs_errors, LIB), Int32,
        (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, UInt64, Ptr{Float64}),
        aux.h.ptr, norm(data.b, 2), norm(data.C, 2), v0, 0, errs))
    return errs
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

# ---- the whole solve as ONE call (SURVEY.md 8f/f1): the native driver of csrc/driver.cu runs the loop of
# src/sdplr.jl:140-449 inside the library.  Field order = sdplrp_config / sdplrp_result of include/sdplrp_b200.h.
struct NativeConfig
    ptol::Float64; gtol::Float64; objtol::Float64; sigma_0::Float64; sigmafac::Float64; maxtime::Float64; printfreq::Float64
    fprec::Float64; prior_trace_bound::Float64; alpha_max::Float64
    maxmajoriter::Int64; maxiter::Int64; numlbfgsvecs::Int64; rankupd_tol::Int64; printlevel::Int64
    gtol_relative::Int64; ptol_relative::Int64; objtol_relative::Int64; eval_DIMACS_errs::Int64; eigval_highprecision::Int64
    seed::UInt64
end
struct NativeResult
    sigma::Float64; grad_norm::Float64; primal_vio::Float64; obj::Float64; L::Float64; max_dual_value::Float64
    min_duality_gap::Float64; totaltime::Float64; dual_time::Float64; primaltime::Float64; DIMACS_time::Float64
    DIMACS_errs::NTuple{6,Float64}
    iter::Int64; majoriter::Int64; lanczos_steps::Int64; r::Int64; status::Int64
end
NativeConfig(c::BurerMonteiroConfig; seed=0) = NativeConfig(c.ptol, c.gtol, c.objtol, c.σ_0, c.σfac, c.maxtime, c.printfreq, c.fprec,
    c.prior_trace_bound, 1.0, c.maxmajoriter, c.maxiter, c.numlbfgsvecs, c.rankupd_tol, c.printlevel,
    c.gtol_mode == "relative", c.ptol_mode == "relative", c.objtol_mode == "relative", c.eval_DIMACS_errs, c.eigval_highprecision, seed)

"""
    sdplr_native(C, As, b, r; kwargs...)

`sdplr` with the outer loop inside the library (`sdplrp_solve`).  Rt0 / λ0 come from `SolverVars(data, r, config)` as
in the reference; the eigenvalue start vectors and the random point of a rank update come from the device generator.
"""
function sdplr_native(C, As, b, r; device=0, seed=0, kwargs...)
    config = BurerMonteiroConfig()
    for (k, v) in kwargs
        hasfield(BurerMonteiroConfig, k) ? setfield!(config, k, v) : @error "Unrecognized keyword argument $k"
    end
    data = SDPData(C, As, b)
    aux = B200Auxiliary(data; device)
    host = SolverVars(data, r, config)
    Rt0 = Matrix(host.Rt); λ0 = Vector(host.λ)
    cfg = Ref(NativeConfig(config; seed)); res = Ref{NativeResult}(); best = zeros(data.m + 1)
    GC.@preserve Rt0 λ0 best check(aux.h, ccall((:sdplrp_solve, LIB), Int32,
        (Ptr{Cvoid}, Ref{NativeConfig}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ref{NativeResult}, Ptr{Float64}),
        aux.h.ptr, cfg, r, Rt0, λ0, norm(b, 2), norm(C, 2), res, best))
    R = res[]
    Rt = Array(DevMat(aux.h, MAT_R, Int(R.r), data.n))
    return Dict("Rt" => Rt, "lambda" => best, "Rt0" => Rt0, "lambda0" => λ0, "sigma" => R.sigma, "grad_norm" => R.grad_norm,
        "primal_vio" => R.primal_vio, "obj" => R.obj, "max_dual_value" => R.max_dual_value, "min_duality_gap" => R.min_duality_gap,
        "totaltime" => R.totaltime, "dual_time" => R.dual_time, "primaltime" => R.primaltime, "iter" => R.iter,
        "majoriter" => R.majoriter, "DIMACS_errs" => collect(R.DIMACS_errs), "ptol" => config.ptol, "objtol" => config.objtol,
        "fprec" => config.fprec, "rankupd_tol" => config.rankupd_tol, "r" => Int(R.r))
end

end # module
