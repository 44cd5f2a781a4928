"""AbstractGPs-facing call surface, backed by libgaplac_b200.so (no CPU fallback).

Mirrors exactly the calls GaPLAC makes [upstream AbstractGPs 0.5.12]:

    fx   = FiniteGP(GP(kernel), X, 0.1)      CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47, CLI/src/sample.jl:25,
                                             src/plotting.jl:6
    lp   = logpdf(fx, y)                     CLI/src/select.jl:49-50 (and per leapfrog step, CLI/src/mcmc.jl:35)
    post = posterior(fx, y)                  CLI/src/select.jl:51-52, src/plotting.jl:8
    m, v = mean_and_var(post, xtest)         src/plotting.jl:12
    y    = rand(fx)                          CLI/src/sample.jl:25

plus the batched form the B200 backend adds: `logpdf_batched` evaluates many hyperparameter proposals /
responses in one call (MCMC proposals and chains, per-feature models), optionally with the analytic
gradient the reference obtains by ForwardDiff through the model body (CLI/src/mcmc.jl:31-37).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .formula import KernelProgram

_ctx: dict[int, "_lib.Context"] = {}


def default_context(device: int = -1) -> "_lib.Context":
    """Process-wide context per device (created on first use; raises without a CUDA device)."""
    if device not in _ctx:
        _ctx[device] = _lib.Context(device)
    return _ctx[device]


class GP:
    """GP(kernel): zero-mean prior (the reference never passes a mean, src/interface.jl:40)."""

    def __init__(self, kernel: KernelProgram):
        self.kernel = kernel
        self._compiled = {}

    def compiled(self, ctx):
        key = id(ctx)
        if key not in self._compiled:
            self._compiled[key] = ctx.program(self.kernel.ops)
        return self._compiled[key]

    def __call__(self, X, sigma2: float = 1e-18, **kw):           # gp(X, 0.1)  (CLI/src/sample.jl:25)
        return FiniteGP(self, X, sigma2, **kw)


class FiniteGP:
    """Marginal of a GP at inputs X with observation-noise VARIANCE sigma2: N(0, K(X,X) + sigma2 I).

    X: (n, d) array, rows are observations (RowVecs(X) / obsdim=1), or an n-vector for d = 1.
    theta: per-call hyperparameter vector for kernels that use Slot(k) hyperparameters."""

    def __init__(self, gp: GP, X, sigma2: float = 1e-18, obsdim: int = 1, theta=(), jitter: float = 0.0, ctx=None):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        elif obsdim == 2:
            X = X.T
        self.gp, self.X, self.sigma2, self.jitter = gp, X, float(sigma2), float(jitter)
        self.theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        self.ctx = ctx or default_context()
        if self.theta.size < gp.kernel.n_theta:
            raise ValueError(f"kernel uses {gp.kernel.n_theta} hyperparameter slots, theta has {self.theta.size}")

    def __len__(self):
        return self.X.shape[0]


def logpdf(fx: FiniteGP, y) -> float:
    """logpdf(fx, y) = -1/2 (n log 2pi + logdet K_y + y' K_y^-1 y).  Raises PosDefException like `cholesky`."""
    prog = fx.gp.compiled(fx.ctx)
    theta = fx.theta.reshape(1, -1) if fx.theta.size else np.zeros((1, 0))
    lml, info = fx.ctx.lml_batched(prog, fx.X, np.asarray(y, dtype=np.float64), theta, fx.sigma2, fx.jitter)
    if info[0] != 0:
        raise _lib.PosDefException(_lib.GPL_ERR_NOTPD, f"covariance not positive definite: pivot {info[0]}")
    return float(lml[0])


def logpdf_batched(fx: FiniteGP, Y, Theta, sigma2=None, grad: bool = False):
    """Many independent evaluations in one call.  Theta: (B, p); Y: (n,) shared or (B, n); sigma2: None
    (fx.sigma2), scalar or (B,).  Non-PD items come back as -Inf with info != 0 (no exception: a sampler rejects)."""
    prog = fx.gp.compiled(fx.ctx)
    s2 = fx.sigma2 if sigma2 is None else sigma2
    return fx.ctx.lml_batched(prog, fx.X, Y, Theta, s2, fx.jitter, grad=grad)


class PosteriorGP:
    def __init__(self, fx: FiniteGP, handle: "_lib.Posterior"):
        self.fx, self.handle = fx, handle

    @property
    def alpha(self) -> np.ndarray:
        return self.handle.alpha()

    def chol_upper(self) -> np.ndarray:
        return self.handle.factor()


def posterior(fx: FiniteGP, y) -> PosteriorGP:
    """posterior(fx, y): Cholesky factor and alpha = K_y^-1 y stay resident in HBM."""
    prog = fx.gp.compiled(fx.ctx)
    h = fx.ctx.posterior_fit(prog, fx.X, np.asarray(y, dtype=np.float64), fx.theta, fx.sigma2, fx.jitter)
    return PosteriorGP(fx, h)


def mean_and_var(post: PosteriorGP, Xtest):
    """(mean, var) of the latent function at Xtest; var excludes sigma2 [upstream AbstractGPs]."""
    Xt = np.asarray(Xtest, dtype=np.float64)
    if Xt.ndim == 1:
        Xt = Xt.reshape(-1, 1)
    return post.handle.mean_and_var(Xt, want_var=True)


def mean(post: PosteriorGP, Xtest):
    Xt = np.asarray(Xtest, dtype=np.float64)
    if Xt.ndim == 1:
        Xt = Xt.reshape(-1, 1)
    return post.handle.mean_and_var(Xt, want_var=False)


def rand(fx: FiniteGP, rng=None, n_samples: int | None = None):
    """rand(fx) = U' z, z ~ N(0, I) drawn on the HOST (the seed/stream stays the caller's)."""
    rng = np.random.default_rng() if rng is None else rng
    S = 1 if n_samples is None else n_samples
    Z = rng.standard_normal((len(fx), S))
    prog = fx.gp.compiled(fx.ctx)
    out = fx.ctx.sample(prog, fx.X, fx.theta, fx.sigma2, Z, fx.jitter)
    return out[:, 0].copy() if n_samples is None else out
