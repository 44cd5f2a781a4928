# GaPLACB200.jl — Julia binding of libgaplac_b200.so (include/gaplac_b200.h) for GaPLAC.
#
# Source only: there is no Julia toolchain in the build image, so this file has not been executed.  The same C ABI
# is exercised through Python ctypes (gaplac_b200/_lib.py), which follows the rules `ccall` imposes (column-major
# Float64, Cint sizes, caller-owned buffers kept alive with GC.@preserve); the Dual / rrule adapter below has an
# executable line-by-line mirror in gaplac_b200/ad.py (tests/test_ad_adapter.py).
#
# What it replaces in GaPLAC (paths relative to the reference tree):
#   kernel(formula; hyperparams)          src/abstractgp_translations.jl:45-71   -> flatten(formula) :: Vector{GplOp}
#   logpdf(fx::FiniteGP, y)               CLI/src/select.jl:49-50, CLI/src/mcmc.jl:35
#   posterior(fx, y) / mean_and_var       CLI/src/select.jl:51-52, src/plotting.jl:8,12
#   rand(fx)                              CLI/src/sample.jl:25
module GaPLACB200

using GaPLAC, AbstractGPs, LinearAlgebra, Random
import ForwardDiff, ChainRulesCore

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
# hyperparams[varname] overrides l / c exactly as makekernel(c, hyperparams[varname(c)]) does (:13-15, :33).  Every value
# that arrives through `hyperparams` - a Float64, or the ForwardDiff.Dual that Turing passes for l (CLI/src/mcmc.jl:32-33) -
# gets a slot of the per-call hyperparameter vector `theta`, so the compiled program does not depend on the value and the
# analytic gradient with respect to it comes back from the GPU.  `Slot(k)` marks an entry of a caller-supplied theta.
struct Slot; index::Int; end

function flatten!(ops, theta, c::GaPLAC.GPOperation, hp, col)
    col = flatten!(ops, theta, c.lhs, hp, col)
    col = flatten!(ops, theta, c.rhs, hp, col)
    c.op in (:add, :multiply) || throw(ArgumentError("Operation $(c.op) not yet supported"))
    push!(ops, GplOp(c.op == :add ? ADD : MUL, 0, -1, -1, 1.0, 1.0))
    return col
end
function flatten!(ops, theta, c::GaPLAC.GPCompnent, hp, col)
    v = GaPLAC.varname(c)
    if c isa GaPLAC.Cat
        haskey(hp, v) && throw(MethodError(GaPLAC.makekernel, (c, hp[v])))
        push!(ops, GplOp(CAT, col, -1, -1, 1.0, 1.0))
    else
        kind = c isa GaPLAC.SqExp ? SQEXP : c isa GaPLAC.OU ? OU : LINEAR
        if haskey(hp, v)
            h = hp[v]
            if h isa Slot
                push!(ops, GplOp(kind, col, Int32(h.index - 1), -1, 1.0, 1.0))
            else                                  # a number (Float64 or Dual): its own slot, value kept in theta
                push!(theta, h)
                push!(ops, GplOp(kind, col, Int32(length(theta) - 1), -1, 1.0, 1.0))
            end
        else
            h = c isa GaPLAC.Linear ? c.intercept : c.lengthscale
            push!(ops, GplOp(kind, col, Int32(-1), -1, Float64(h), 1.0))
        end
    end
    return col + Int32(1)
end
function flatten(formula; hyperparams = Dict())
    ops = GplOp[]; theta = Real[]
    flatten!(ops, theta, formula, hyperparams, Int32(0))
    return ops, (isempty(theta) ? Float64[] : [t for t in theta])      # concretely typed: Vector{Float64} or Vector{<:Dual}
end

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
struct B200Kernel{T<:Real} <: AbstractGPs.Kernel
    prog::Program
    ncols::Int
    theta::Vector{T}          # values of the hyperparameters that came through `hyperparams` (Float64 or Dual)
end
function B200Kernel(formula; hyperparams = Dict())
    ops, theta = flatten(formula; hyperparams)
    return B200Kernel(Program(ops), length(GaPLAC.varnames(formula)), theta)
end

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

# value + analytic gradient of one log-density: (lml, dlml/dtheta, dlml/dy); PosDefException like cholesky() [upstream]
function value_and_gradient(fx::B200FiniteGP, y::Vector{Float64}, theta::Vector{Float64})
    lml, info, dth, dy = batched_logpdf(fx, y, reshape(theta, :, 1); grad = true)
    info[1] == 0 || throw(PosDefException(info[1]))
    return lml[1], vec(dth), vec(dy)
end

# Plain Float64 call (select: CLI/src/select.jl:49-50).
function b200_logpdf(fx::B200FiniteGP, y::Vector{Float64}, theta::Vector{Float64})
    lml, info = batched_logpdf(fx, y, reshape(theta, :, 1))
    info[1] == 0 || throw(PosDefException(info[1]))
    return lml[1]
end

# ForwardDiff adapter for the mcmc model body (CLI/src/mcmc.jl:31-37: Turing differentiates `fx ~ FiniteGP(...)` with Duals
# in l - through kernel(eq; hyperparams = Dict(v => l)) - and in fx).  Strip the duals, evaluate value + analytic gradient
# ONCE on the GPU, reassemble Dual(value, sum_k dlml/dtheta_k * partials(theta_k) + sum_i dlml/dy_i * partials(y_i)).
# (gaplac_b200/ad.py is the executable mirror of these lines; tests/test_ad_adapter.py drives it in chunk mode.)
function b200_logpdf(fx::B200FiniteGP, y::AbstractVector{<:Real}, theta::AbstractVector{<:Real})
    T = promote_type(eltype(y), eltype(theta))
    T <: ForwardDiff.Dual || return b200_logpdf(fx, Vector{Float64}(y), Vector{Float64}(theta))
    vy = Float64[ForwardDiff.value(v) for v in y]
    vth = Float64[ForwardDiff.value(t) for t in theta]
    lml, dth, dy = value_and_gradient(fx, vy, vth)
    acc = zero(ForwardDiff.partials(zero(T)))                          # Partials{N,V} of zeros
    for k in eachindex(theta)
        theta[k] isa ForwardDiff.Dual && (acc += dth[k] * ForwardDiff.partials(theta[k]))
    end
    for i in eachindex(y)
        y[i] isa ForwardDiff.Dual && (acc += dy[i] * ForwardDiff.partials(y[i]))
    end
    return T(lml, acc)                                                 # Dual{Tag}(value, partials)
end

# Reverse mode (Zygote / ReverseDiff through ChainRules): the pullback is the same analytic gradient.
function ChainRulesCore.rrule(::typeof(b200_logpdf), fx::B200FiniteGP, y::Vector{Float64}, theta::Vector{Float64})
    lml, dth, dy = value_and_gradient(fx, y, theta)
    pullback(dl) = (ChainRulesCore.NoTangent(), ChainRulesCore.NoTangent(), dl .* dy, dl .* dth)
    return lml, pullback
end

# logpdf(fx, y) with the hyperparameters the kernel was built with; y and / or theta may carry Duals.
AbstractGPs.logpdf(fx::B200FiniteGP, y::AbstractVector{<:Real}) = b200_logpdf(fx, y, fx.f.kernel.theta)

# ---- posterior / mean_and_var / rand ------------------------------------------------------------------------------------
mutable struct B200Posterior
    h::Ptr{Cvoid}; n::Int; d::Int
end
kernel_theta(fx) = Float64[ForwardDiff.value(t) for t in fx.f.kernel.theta]
function AbstractGPs.posterior(fx::B200FiniteGP, y::AbstractVector{<:Real}; theta = kernel_theta(fx), jitter = 0.0)
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

function Base.rand(rng::Random.AbstractRNG, fx::B200FiniteGP; theta = kernel_theta(fx), jitter = 0.0)
    X = rowmatrix(fx.x); n, d = size(X); z = randn(rng, n); out = similar(z)   # the RNG stays Julia's
    GC.@preserve X z out theta check(ctx().h, ccall((:gpl_sample, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Float64, Float64, Ptr{Float64}, Cint, Ptr{Float64}),
        ctx().h, fx.f.kernel.prog.h, n, d, X, theta, length(theta), noisevar(fx), jitter, z, 1, out))
    return out
end

# ---- the sampler on the device: replaces `sample(m, NUTS(0.65), N)` of CLI/src/mcmc.jl:39-41 -----------------------------
# One chain per column of Y (n x B), or `chains` chains of one response.  `kernel` must take its sampled hyperparameters
# from Slot(1..p); lo / hi are the Uniform prior bounds (the reference: l ~ Uniform(0, 20), mcmc.jl:32).  Returns a
# NamedTuple of arrays laid out as the C ABI documents them (theta: p x n_rec x B, lp: n_rec x B, ...).
struct McmcOpts
    n_samples::Int32; n_adapt::Int32; max_depth::Int32; latent::Int32; search_eps::Int32; adapt_mass::Int32
    record_warmup::Int32; chain_offset::Int32; delta::Float64; max_dh::Float64; obs_sd::Float64; eps0::Float64; seed::UInt64
end
function nuts(fx::B200FiniteGP, Y::AbstractVecOrMat{Float64}, lo::Vector{Float64}, hi::Vector{Float64};
              samples::Integer = 200, adapt::Integer = -1, seed::Integer = 0, chains::Integer = 1, latent::Bool = true,
              obs_sd = 1.0, jitter = 0.0, record_warmup::Bool = false)
    X = rowmatrix(fx.x); n, d = size(X); p = length(lo)
    B = ndims(Y) == 2 ? size(Y, 2) : chains
    dim = p + (latent ? n : 0)
    na = adapt >= 0 ? adapt : min(1000, samples ÷ 2)
    nrec = samples + (record_warmup ? na : 0)
    opts = Ref(McmcOpts(samples, na, 10, latent, 1, 1, record_warmup, 0, 0.65, 1000.0, obs_sd, 0.1, seed))
    q0 = zeros(dim, B); s2 = [noisevar(fx)]
    theta = Array{Float64}(undef, p, nrec, B); lp = Matrix{Float64}(undef, nrec, B)
    accept = similar(lp); eps = similar(lp)
    depth = Matrix{Cint}(undef, nrec, B); nleap = similar(depth); div = similar(depth); status = Vector{Cint}(undef, B)
    evals = Ref{Clonglong}(0)
    GC.@preserve X Y lo hi s2 q0 theta lp accept eps depth nleap div status check(ctx().h, ccall((:gpl_mcmc_nuts, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Cint, Float64, Cint, Ptr{Float64}, Ref{McmcOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ref{Clonglong}),
        ctx().h, fx.f.kernel.prog.h, n, d, X, 0, Y, ndims(Y) == 2 ? 1 : 0, p, lo, hi, s2, 0, jitter, B, q0, opts,
        theta, lp, C_NULL, accept, eps, depth, nleap, div, status, evals))
    return (; theta, lp, accept, eps, depth, n_leapfrog = nleap, divergent = div, status, grad_evals = evals[])
end

end # module
