# GaPLACB200.jl — Julia binding of libgaplac_b200.so (include/gaplac_b200.h) for GaPLAC.
#
# Source only: there is no Julia toolchain in the build image, so this file has not been executed.  The same C ABI
# is exercised through Python ctypes (gaplac_b200/_lib.py), which follows the rules `ccall` imposes (column-major
# Float64, Cint sizes, caller-owned buffers kept alive with GC.@preserve).
#
# What it replaces in GaPLAC (paths relative to the reference tree):
#   kernel(formula; hyperparams)          src/abstractgp_translations.jl:45-71   -> flatten(formula) :: Vector{GplOp}
#   logpdf(fx::FiniteGP, y)               CLI/src/select.jl:49-50, CLI/src/mcmc.jl:35
#   posterior(fx, y) / mean_and_var       CLI/src/select.jl:51-52, src/plotting.jl:8,12
#   rand(fx)                              CLI/src/sample.jl:25
module GaPLACB200

using GaPLAC, AbstractGPs, LinearAlgebra, Random

const LIB = get(ENV, "GAPLAC_B200_LIB", "libgaplac_b200.so")

# ---- struct gpl_op (32 bytes) and node kinds ------------------------------------------------------------------------
struct GplOp
    kind::Int32
    col::Int32
    theta_slot::Int32
    var_slot::Int32
    value::Float64
    var::Float64
end
const SQEXP, OU, LINEAR, CAT, CONSTANT, NOISE, ADD, MUL = Int32.(0:7)

check(ctx, rc) = rc == 0 ? nothing :
    (msg = unsafe_string(ccall((:gpl_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx));
     rc == -4 ? throw(PosDefException(parse(Int, match(r"pivot (\d+)", msg)[1]))) : error("gaplac_b200 ($rc): $msg"))

# ---- context --------------------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = -1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:gpl_init, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || error(unsafe_string(ccall((:gpl_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        c = new(r[])
        finalizer(c -> ccall((:gpl_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c)
    end
end
const CTX = Ref{Context}()
ctx() = isassigned(CTX) ? CTX[] : (CTX[] = Context())

# ---- formula AST -> postfix kernel-program (leaf i reads column i, src/abstractgp_translations.jl:45-71) ---------------
# hyperparams[varname] overrides l / c exactly as makekernel(c, hyperparams[varname(c)]) does (:13-15, :33);
# a value of type `Slot` marks an entry of the per-item hyperparameter vector of a batched call.
struct Slot; index::Int; end
hyper(h::Slot) = (Int32(h.index - 1), 1.0)
hyper(h::Real) = (Int32(-1), Float64(h))

function flatten!(ops, c::GaPLAC.GPOperation, hp, col)
    col = flatten!(ops, c.lhs, hp, col)
    col = flatten!(ops, c.rhs, hp, col)
    c.op in (:add, :multiply) || throw(ArgumentError("Operation $(c.op) not yet supported"))
    push!(ops, GplOp(c.op == :add ? ADD : MUL, 0, -1, -1, 1.0, 1.0))
    return col
end
function flatten!(ops, c::GaPLAC.GPCompnent, hp, col)
    v = GaPLAC.varname(c)
    if c isa GaPLAC.Cat
        haskey(hp, v) && throw(MethodError(GaPLAC.makekernel, (c, hp[v])))
        push!(ops, GplOp(CAT, col, -1, -1, 1.0, 1.0))
    else
        h = get(hp, v, c isa GaPLAC.Linear ? c.intercept : c.lengthscale)
        slot, val = hyper(h)
        kind = c isa GaPLAC.SqExp ? SQEXP : c isa GaPLAC.OU ? OU : LINEAR
        push!(ops, GplOp(kind, col, slot, -1, val, 1.0))
    end
    return col + Int32(1)
end
flatten(formula; hyperparams = Dict()) = (ops = GplOp[]; flatten!(ops, formula, hyperparams, Int32(0)); ops)

mutable struct Program
    h::Ptr{Cvoid}
    function Program(ops::Vector{GplOp})
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx().h, ccall((:gpl_program_create, LIB), Cint, (Ptr{Cvoid}, Ptr{GplOp}, Cint, Ref{Ptr{Cvoid}}),
                             ctx().h, ops, length(ops), r))
        p = new(r[])
        finalizer(p -> ccall((:gpl_program_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), p)
    end
end

# A GP whose kernel is a B200 kernel-program: GP(B200Kernel(formula)) drops into make_gp (src/interface.jl:36-41).
struct B200Kernel <: AbstractGPs.Kernel
    prog::Program
    ncols::Int
end
B200Kernel(formula; hyperparams = Dict()) = B200Kernel(Program(flatten(formula; hyperparams)), length(GaPLAC.varnames(formula)))

const B200FiniteGP = AbstractGPs.FiniteGP{<:AbstractGPs.GP{<:AbstractGPs.ZeroMean, <:B200Kernel}}
rowmatrix(x::AbstractGPs.RowVecs) = Matrix{Float64}(x.X)                # n x d, column-major: exactly the ABI layout
rowmatrix(x::AbstractVector{<:Real}) = reshape(Vector{Float64}(x), :, 1)
noisevar(fx) = (d = diag(fx.Σy); all(==(d[1]), d) || error("heteroscedastic noise not supported"); d[1])

# ---- logpdf ----------------------------------------------------------------------------------------------------------
function batched_logpdf(fx::B200FiniteGP, Y::AbstractVecOrMat{Float64}, Theta::Matrix{Float64};
                        grad::Bool = false, jitter::Float64 = 0.0)
    X = rowmatrix(fx.x); n, d = size(X); p, B = size(Theta)
    lml = Vector{Float64}(undef, B); info = zeros(Cint, B); s2 = [noisevar(fx)]
    dth = grad ? Matrix{Float64}(undef, p, B) : Matrix{Float64}(undef, 0, 0)
    dy = grad ? Matrix{Float64}(undef, n, B) : Matrix{Float64}(undef, 0, 0)
    GC.@preserve X Y Theta s2 lml info dth dy begin
        check(ctx().h, ccall((:gpl_lml_batched, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint,
             Ptr{Float64}, Cint, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
            ctx().h, fx.f.kernel.prog.h, n, d, X, 0, Y, ndims(Y) == 2 ? 1 : 0, Theta, p, s2, 0, jitter, B,
            lml, grad ? pointer(dth) : C_NULL, grad ? pointer(dy) : C_NULL, info))
    end
    return grad ? (lml, info, dth, dy) : (lml, info)
end

function AbstractGPs.logpdf(fx::B200FiniteGP, y::AbstractVector{<:Real})
    lml, info = batched_logpdf(fx, Vector{Float64}(y), zeros(Float64, 0, 1))
    info[1] == 0 || throw(PosDefException(info[1]))          # what cholesky() throws [upstream]
    return lml[1]
end

# ForwardDiff adapter for the mcmc model body (CLI/src/mcmc.jl:31-37): strip the duals, evaluate value + analytic
# gradient on the GPU, reassemble Dual(value, sum_k d/dtheta_k * partials(theta_k) + sum_i d/dy_i * partials(y_i)).
# (ChainRulesCore.rrule is the same two lines with the pullback (dth, dy).)

# ---- posterior / mean_and_var / rand ------------------------------------------------------------------------------------
mutable struct B200Posterior
    h::Ptr{Cvoid}; n::Int; d::Int
end
function AbstractGPs.posterior(fx::B200FiniteGP, y::AbstractVector{<:Real}; theta = Float64[], jitter = 0.0)
    X = rowmatrix(fx.x); n, d = size(X); yv = Vector{Float64}(y); r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X yv theta check(ctx().h, ccall((:gpl_posterior_fit, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Float64, Ref{Ptr{Cvoid}}),
        ctx().h, fx.f.kernel.prog.h, n, d, X, yv, theta, length(theta), noisevar(fx), jitter, r))
    p = B200Posterior(r[], n, d)
    finalizer(p -> ccall((:gpl_posterior_free, LIB), Cint, (Ptr{Cvoid},), p.h), p)
end
function AbstractGPs.mean_and_var(p::B200Posterior, xtest)
    Xs = rowmatrix(xtest); m = size(Xs, 1); mu = Vector{Float64}(undef, m); v = similar(mu)
    GC.@preserve Xs mu v check(ctx().h, ccall((:gpl_posterior_mean_var, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), p.h, m, Xs, mu, v))
    return mu, v
end
# The loop behind `predict` / `fitplot` (CLI/src/main.jl:8-16): one posterior per column of Theta (p x B chain of
# hyperparameter draws) over the same (X, y), mean and variance at xtest for each - one library call.
# Returns (mean m x B, var m x B, lml B, info B); rows with info != 0 were not positive definite (NaN predictions).
function predict_chain(fx::B200FiniteGP, y::AbstractVector{<:Real}, Theta::Matrix{Float64}, xtest;
                       sigma2 = [noisevar(fx)], jitter::Float64 = 0.0)
    X = rowmatrix(fx.x); n, d = size(X); p, B = size(Theta); Xs = rowmatrix(xtest); m = size(Xs, 1)
    yv = Vector{Float64}(y); s2 = Vector{Float64}(sigma2)
    mu = Matrix{Float64}(undef, m, B); v = Matrix{Float64}(undef, m, B)
    lml = Vector{Float64}(undef, B); info = zeros(Cint, B)
    GC.@preserve X yv Theta s2 Xs mu v lml info check(ctx().h, ccall((:gpl_predict_batched, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64,
         Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
        ctx().h, fx.f.kernel.prog.h, n, d, X, yv, Theta, p, s2, length(s2) == B && B > 1 ? 1 : 0, jitter, B, m, Xs,
        mu, v, lml, info))
    return mu, v, lml, info
end

function Base.rand(rng::Random.AbstractRNG, fx::B200FiniteGP; theta = Float64[], jitter = 0.0)
    X = rowmatrix(fx.x); n, d = size(X); z = randn(rng, n); out = similar(z)   # the RNG stays Julia's
    GC.@preserve X z out theta check(ctx().h, ccall((:gpl_sample, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Float64, Ptr{Float64}, Cint, Ptr{Float64}),
        ctx().h, fx.f.kernel.prog.h, n, d, X, theta, length(theta), noisevar(fx), jitter, z, 1, out))
    return out
end

end # module
