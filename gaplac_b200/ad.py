"""Forward-mode AD adapter: how `mcmc` reaches the GPU.

The reference's sampler differentiates the model body with ForwardDiff (CLI/src/mcmc.jl:31-37 under `sample(m, NUTS...)`):
`logpdf(FiniteGP, fx)` is called with Dual numbers in the hyperparameter l and in the latent vector fx, and the generic
Julia Cholesky runs on Duals (several full factorisations per gradient).  The adapter replaces that by ONE value +
analytic-gradient call:

    strip:       values of (theta, y)                                      Float64 vectors
    evaluate:    lml, dlml/dtheta, dlml/dy  = gpl_lml_batched(..., grad)   one factorisation on the GPU
    reassemble:  Dual(lml, sum_k dlml/dtheta_k * partials(theta_k) + sum_i dlml/dy_i * partials(y_i))

`julia/GaPLACB200.jl` has the same three steps as methods on ForwardDiff.Dual (and the ChainRulesCore.rrule with the
pullback (dtheta, dy)); this module is their executable mirror: `Dual` is a minimal stand-in for ForwardDiff.Dual{T,V,N}
(value + N partials), `logpdf_dual` the adapter.  tests/test_ad_adapter.py drives it in ForwardDiff's chunked mode."""
from __future__ import annotations

import numpy as np


class Dual:
    """value + partials (ForwardDiff.Dual{Tag, Float64, N}); just enough arithmetic for the tests."""
    __slots__ = ("value", "partials")

    def __init__(self, value: float, partials):
        self.value = float(value)
        self.partials = np.asarray(partials, dtype=np.float64)

    def _lift(self, o):
        return o if isinstance(o, Dual) else Dual(o, np.zeros_like(self.partials))

    def __add__(self, o):
        o = self._lift(o)
        return Dual(self.value + o.value, self.partials + o.partials)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.value, -self.partials)

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        return Dual(self.value * o.value, self.value * o.partials + o.value * self.partials)

    __rmul__ = __mul__

    def __repr__(self):
        return f"Dual({self.value}, {self.partials})"


def value(x) -> float:
    return x.value if isinstance(x, Dual) else float(x)


def npartials(*vectors) -> int:
    for v in vectors:
        for x in v:
            if isinstance(x, Dual):
                return len(x.partials)
    return 0


def logpdf_dual(evaluate, y, theta):
    """logpdf with Dual inputs.  evaluate(y_values, theta_values) -> (lml, dlml/dtheta, dlml/dy) is the backend call
    (gaplac_b200.gp.logpdf_batched(..., grad=True) on the GPU).  Returns a float when no input is a Dual."""
    vy = np.array([value(v) for v in y], dtype=np.float64)
    vth = np.array([value(t) for t in theta], dtype=np.float64)
    N = npartials(y, theta)
    if N == 0:
        return float(evaluate(vy, vth)[0])
    lml, dth, dy = evaluate(vy, vth)
    acc = np.zeros(N)
    for k, t in enumerate(theta):
        if isinstance(t, Dual):
            acc += dth[k] * t.partials
    for i, v in enumerate(y):
        if isinstance(v, Dual):
            acc += dy[i] * v.partials
    return Dual(lml, acc)


def gpu_evaluator(fx):
    """Backend for logpdf_dual on a FiniteGP whose kernel reads its hyperparameters from theta slots."""
    from . import gp as _gp
    from . import _lib

    def evaluate(vy, vth):
        lml, info, dth, dy = _gp.logpdf_batched(fx, vy, vth.reshape(1, -1), grad=True)
        if info[0] != 0:
            raise _lib.PosDefException(_lib.GPL_ERR_NOTPD, f"covariance not positive definite: pivot {info[0]}")
        return float(lml[0]), dth[0], dy[0]

    return evaluate


def gradient_chunked(f, x, chunk: int = 12):
    """ForwardDiff.gradient in chunk mode: ceil(len(x) / chunk) passes of f over Duals seeded with unit partials
    (the reference needs ~5 passes for the 51-dimensional README model, each a full Dual Cholesky; here every pass is one
    GPU call)."""
    x = np.asarray(x, dtype=np.float64)
    g = np.empty_like(x)
    val = None
    for lo in range(0, len(x), chunk):
        hi = min(lo + chunk, len(x))
        xs = [Dual(v, np.eye(hi - lo)[k - lo] if lo <= k < hi else np.zeros(hi - lo)) for k, v in enumerate(x)]
        out = f(xs)
        val = out.value
        g[lo:hi] = out.partials
    return val, g
