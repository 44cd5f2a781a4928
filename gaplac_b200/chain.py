"""Posterior predictions averaged over an MCMC chain of hyperparameters, and the chain-based evidence of `select`.

This is SURVEY.md 8(f) item 2/3: the compute behind the `predict` / `fitplot` commands the reference only stubs
(CLI/src/main.jl:8-16; flags bin/old_options.jl:62-87; output columns ymu / yQ050 / yQ950 per test/pred.jl:11-14) and the
`select --chains` evidence (CLI/src/select.jl:15-20).  Every chain row is one posterior fit + one mean_and_var on the
GPU - all rows in one batch through gpl_predict_batched, or one by one (batched=False) -; the mixture over rows is
summarised on the host.
"""
from __future__ import annotations

import numpy as np

from . import gp as _gp


def mixture_summary(mu: np.ndarray, var: np.ndarray, quantiles=(0.05, 0.5, 0.95)):
    """Summaries of the equal-weight Gaussian mixture  (1/S) sum_s N(mu[s, :], var[s, :])  per column.

    Returns (mean[m], sd[m], Q[len(quantiles), m]).  Quantiles by safeguarded Newton / bisection on the mixture CDF inside
    a bracket that contains every component's +-10 sd."""
    from scipy.special import ndtr

    mu = np.atleast_2d(np.asarray(mu, dtype=np.float64))
    sd_s = np.sqrt(np.maximum(np.atleast_2d(np.asarray(var, dtype=np.float64)), 0.0))
    mean = mu.mean(axis=0)
    second = (sd_s ** 2 + mu ** 2).mean(axis=0)
    sd = np.sqrt(np.maximum(second - mean ** 2, 0.0))
    lo0 = (mu - 10.0 * sd_s - 1e-300).min(axis=0)
    hi0 = (mu + 10.0 * sd_s + 1e-300).max(axis=0)
    safe = np.where(sd_s > 0.0, sd_s, 1.0)
    qs = np.asarray(quantiles, dtype=np.float64).reshape(-1, 1)          # all quantiles refined together
    lo = np.repeat(lo0[None, :], qs.shape[0], axis=0)
    hi = np.repeat(hi0[None, :], qs.shape[0], axis=0)
    atoms = sd_s <= 0.0

    def cdf_at(x):                                                        # x: (nq, m) -> mixture CDF, z-scores
        z = (x[:, None, :] - mu[None, :, :]) / safe[None, :, :]
        c = ndtr(z)
        if atoms.any():
            c = np.where(atoms[None, :, :], (x[:, None, :] >= mu[None, :, :]).astype(np.float64), c)
        return c.mean(axis=1), z

    # Safeguarded Newton on the monotone mixture CDF: every evaluation tightens the bracket [lo, hi]; the Newton step is
    # taken when it stays inside, the midpoint otherwise (always with point masses); stops when every bracket is below
    # 1e-13 of its scale (a few Newton steps for smooth mixtures, plain bisection in the worst case).
    scale = np.maximum(hi0 - lo0, 1e-300)[None, :]
    Q = 0.5 * (lo + hi)
    smooth = not atoms.any()
    for _ in range(200):
        c, z = cdf_at(Q)
        below = c < qs
        lo = np.where(below, Q, lo)
        hi = np.where(below, hi, Q)
        mid = 0.5 * (lo + hi)
        if smooth:
            pdf = (np.exp(-0.5 * z * z) / (safe[None, :, :] * np.sqrt(2.0 * np.pi))).mean(axis=1)
            newton = Q - (c - qs) / np.where(pdf > 0.0, pdf, 1.0)
            ok = (pdf > 0.0) & (newton > lo) & (newton < hi)
            nxt = np.where(ok, newton, mid)
        else:
            nxt = mid
        moved = np.abs(nxt - Q)
        Q = nxt
        # converged: the iterate no longer moves (Newton approaches from one side, so the bracket need not shrink) or
        # the bracket itself is at rounding level
        if np.all((moved <= 1e-14 * scale) | (hi - lo <= 1e-14 * scale)):
            break
    return mean, sd, Q


def predict_chain(gp: "_gp.GP", X, y, chain_theta, Xtest, sigma2=0.0, jitter: float = 0.0, obs_var=None,
                  quantiles=(0.05, 0.5, 0.95), ctx=None, batched: bool = True):
    """Posterior of the latent f (and of a new observation y*) at Xtest, averaged over the rows of `chain_theta`.

    chain_theta: (S, p) hyperparameter draws (one posterior fit per row).  sigma2 / obs_var: scalar or (S,): the noise
    variance on the diagonal of K_y, and the variance added to the latent variance for a NEW observation (defaults to
    sigma2; models with a Noise term pass that term's variance draw).  Returns a dict with fmu, fsd, fQ###, ymu, ysd,
    yQ### (### = 1000 q, e.g. yQ050, yQ950 as in test/pred.jl:11-14) and the per-row `mu`, `var` (S x m)."""
    Th = np.atleast_2d(np.asarray(chain_theta, dtype=np.float64))
    S = Th.shape[0]
    s2 = np.broadcast_to(np.asarray(sigma2, dtype=np.float64), (S,))
    ov = s2 if obs_var is None else np.broadcast_to(np.asarray(obs_var, dtype=np.float64), (S,))
    Xt = np.asarray(Xtest, dtype=np.float64)
    if Xt.ndim == 1:
        Xt = Xt.reshape(-1, 1)
    ctx = ctx or _gp.default_context()
    Xa = np.asarray(X, dtype=np.float64)
    if Xa.ndim == 1:
        Xa = Xa.reshape(-1, 1)
    if batched:
        # all rows factored in one lockstep batch, predictions over (row, slab) pairs: gpl_predict_batched
        mu, var, _lml, info = ctx.predict_batched(gp.compiled(ctx), Xa, np.asarray(y, dtype=np.float64), Th,
                                                  s2 if S > 1 else float(s2[0]), Xt, jitter)
        if info.any():
            raise _gp._lib.PosDefException(_gp._lib.GPL_ERR_NOTPD,
                                           f"covariance not positive definite for chain row {int(np.flatnonzero(info)[0])}")
    else:
        mu = np.empty((S, Xt.shape[0]))
        var = np.empty_like(mu)
        for s in range(S):
            fx = _gp.FiniteGP(gp, Xa, float(s2[s]), theta=Th[s], jitter=jitter, ctx=ctx)
            post = _gp.posterior(fx, y)
            mu[s], var[s] = _gp.mean_and_var(post, Xt)
            post.handle.free()
    out = {"mu": mu, "var": var}
    for name, v in (("f", var), ("y", var + ov[:, None])):
        mean, sd, Q = mixture_summary(mu, v, quantiles)
        out[name + "mu"], out[name + "sd"] = mean, sd
        for q, row in zip(quantiles, Q):
            out["%sQ%03d" % (name, int(round(1000 * q)))] = row
    return out


def log2_harmmean_exp2(lp_chain) -> float:
    """log2(harmmean(2 .^ lp)) exactly as `select --chains` computes it (CLI/src/select.jl:15-19, there with BigFloat
    powers of two), as a stable base-2 log-sum-exp:  log2 S - log2 sum_i 2^(-lp_i).  The reference treats the chain's
    natural-log densities as base-2 exponents; this reproduces that number, quirk included."""
    lp = np.asarray(lp_chain, dtype=np.float64)
    m = np.max(-lp)
    return float(np.log2(lp.size) - (m + np.log2(np.sum(np.exp2(-lp - m)))))


def select_chains_log2_bayes(lp_chain_1, lp_chain_2) -> float:
    """The "Log2 Bayes" figure of `select --chains a.tsv b.tsv` (CLI/src/select.jl:15-20): lp1 - lp2 with
    lp_k = log2(harmmean(2 .^ lp_k_chain)).  Positive values favour model 1 (select.jl:65)."""
    return log2_harmmean_exp2(lp_chain_1) - log2_harmmean_exp2(lp_chain_2)


def select_formulae_log2_bayes(lp1: float, lp2: float) -> float:
    """`select --formulae`: log2(2^lp1 / 2^lp2) = lp1 - lp2 (CLI/src/select.jl:54)."""
    return float(lp1) - float(lp2)


def log_evidence_harmonic(loglik_chain) -> float:
    """NOT the reference's number: the natural-log harmonic-mean evidence estimate  log Z = log S - logsumexp(-ll)  of a
    chain of log-likelihoods - what the reference's base-2 expression (see log2_harmmean_exp2) would be if it exponentiated
    with e.  Offered as the statistically meaningful variant; `select_chains_log2_bayes` is the drop-in."""
    ll = np.asarray(loglik_chain, dtype=np.float64)
    m = np.max(-ll)
    return float(np.log(ll.size) - (m + np.log(np.sum(np.exp(-ll - m)))))


def log_bayes_factor(loglik_chain_1, loglik_chain_2) -> float:
    """log10 Bayes factor of model 1 over model 2 from the natural-log harmonic means (a deliberate departure from the
    reference's base-2 figure; see select_chains_log2_bayes for that one)."""
    return (log_evidence_harmonic(loglik_chain_1) - log_evidence_harmonic(loglik_chain_2)) / np.log(10.0)
