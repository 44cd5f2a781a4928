"""The ForwardDiff-Dual adapter (gaplac_b200/ad.py, mirrored in julia/GaPLACB200.jl): the mcmc model body of
CLI/src/mcmc.jl:31-37 written over Dual numbers, differentiated in ForwardDiff's chunked mode, must give the gradient the
sampler needs.  CPU: the backend is the oracle (the strip / evaluate / reassemble logic is what is under test); the GPU
backend runs in test_gpu_ad_adapter below."""
import numpy as np
import pytest

from gaplac_b200 import ad, workloads as W
from oracle import gp_oracle as O


def _model_body(evaluate, Y):
    """log joint of mcmc.jl:31-37 in the constrained space: l ~ U(0, 20); fx ~ FiniteGP(...); Y .~ Normal(fx, 1)."""
    n = len(Y)

    def f(x):                                   # x = (l, fx_1..fx_n), Duals
        ell, fx = x[0], x[1:]
        lp = ad.logpdf_dual(evaluate, fx, [ell])                    # fx ~ FiniteGP(GP(k(l)), X, 0.1)
        quad = 0.0
        for yi, fi in zip(Y, fx):
            r = yi - fi
            quad = quad + r * r
        return lp + (-np.log(20.0) - 0.5 * n * O.LOG2PI) + (-0.5) * quad
    return f


def test_chunked_forward_mode_gradient_of_the_model_body_matches_the_oracle():
    d = W.make_c1()
    ops, X, Y = d["ops"], d["X"], d["y"]
    calls = []

    def evaluate(vy, vth):
        calls.append(1)
        return O.lml_grad(ops, X, vy, vth, 0.1)

    rng = np.random.default_rng(0)
    x = np.concatenate([[2.5], rng.standard_normal(50)])
    val, g = ad.gradient_chunked(_model_body(evaluate, Y), x, chunk=12)
    rval, rdth, rdfx = O.mcmc_logjoint(ops, X, Y, [2.5], x[1:])
    assert abs(val - rval) < 1e-12 * abs(rval)
    assert abs(g[0] - rdth[0]) < 1e-10 * max(1.0, abs(rdth[0]))
    assert np.max(np.abs(g[1:] - rdfx)) < 1e-10
    assert len(calls) == 5                      # ceil(51 / 12) passes: one backend call each (the reference: 5 Dual Choleskys)


def test_plain_floats_pass_through_and_partials_compose_linearly():
    d = W.make_c1(n=12)
    ev = lambda vy, vth: O.lml_grad(d["ops"], d["X"], vy, vth, 0.1)
    y = list(d["y"])
    assert isinstance(ad.logpdf_dual(ev, y, [1.5]), float)
    # a direction that moves l and every y_i at once: one partial = directional derivative
    v = np.linspace(-1, 1, 12)
    out = ad.logpdf_dual(ev, [ad.Dual(a, [b]) for a, b in zip(y, v)], [ad.Dual(1.5, [0.7])])
    _, dth, dy = ev(np.array(y), np.array([1.5]))
    assert abs(out.partials[0] - (0.7 * dth[0] + v @ dy)) < 1e-12
    # central finite difference of the same direction
    h = 1e-6
    fp = ev(np.array(y) + h * v, np.array([1.5 + 0.7 * h]))[0]
    fm = ev(np.array(y) - h * v, np.array([1.5 - 0.7 * h]))[0]
    assert abs(out.partials[0] - (fp - fm) / (2 * h)) < 1e-6 * max(1.0, abs(out.partials[0]))


@pytest.mark.gpu
def test_gpu_ad_adapter(ctx):
    """The same chunked gradient with the GPU as backend (one gpl_lml_batched value+gradient call per pass)."""
    import gaplac_b200 as G
    d = W.make_c1()
    gp = G.GP(G.kernel(G.SqExp("x", l=G.Slot(0)))[0])
    fxgp = G.FiniteGP(gp, d["X"], 0.1, theta=[1.0], ctx=ctx)
    rng = np.random.default_rng(1)
    x = np.concatenate([[3.1], rng.standard_normal(50)])
    val, g = ad.gradient_chunked(_model_body(ad.gpu_evaluator(fxgp), d["y"]), x, chunk=11)
    rval, rdth, rdfx = O.mcmc_logjoint(d["ops"], d["X"], d["y"], [3.1], x[1:])
    assert abs(val - rval) < 1e-9 * abs(rval)
    assert abs(g[0] - rdth[0]) < 1e-8 * max(1.0, abs(rdth[0]))
    assert np.max(np.abs(g[1:] - rdfx)) < 1e-8
