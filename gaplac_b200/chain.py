"""Posterior predictions averaged over an MCMC chain of hyperparameters, and the chain-based evidence of `select`.

This is SURVEY.md 8(f) item 2/3: the compute behind the `predict` / `fitplot` commands the reference only stubs
(CLI/src/main.jl:8-16; flags bin/old_options.jl:62-87; output columns ymu / yQ050 / yQ950 per test/pred.jl:11-14) and the
`select --chains` evidence (CLI/src/select.jl:15-20).  Every chain row is one posterior fit + one mean_and_var on the
GPU (gaplac_b200.gp); the mixture over rows is summarised on the host.
"""
from __future__ import annotations

import numpy as np

from . import gp as _gp


def mixture_summary(mu: np.ndarray, var: np.ndarray, quantiles=(0.05, 0.5, 0.95), iters: int = 60):
    """Summaries of the equal-weight Gaussian mixture  (1/S) sum_s N(mu[s, :], var[s, :])  per column.

    Returns (mean[m], sd[m], Q[len(quantiles), m]).  Quantiles by bisection on the mixture CDF (monotone; 60 halvings
    of a bracket that contains every component's +-10 sd reach double precision)."""
    from scipy.special import ndtr

    mu = np.atleast_2d(np.asarray(mu, dtype=np.float64))
    sd_s = np.sqrt(np.maximum(np.atleast_2d(np.asarray(var, dtype=np.float64)), 0.0))
    mean = mu.mean(axis=0)
    second = (sd_s ** 2 + mu ** 2).mean(axis=0)
    sd = np.sqrt(np.maximum(second - mean ** 2, 0.0))
    lo0 = (mu - 10.0 * sd_s - 1e-300).min(axis=0)
    hi0 = (mu + 10.0 * sd_s + 1e-300).max(axis=0)
    safe = np.where(sd_s > 0.0, sd_s, 1.0)
    Q = np.empty((len(quantiles), mu.shape[1]))
    for qi, q in enumerate(quantiles):
        lo, hi = lo0.copy(), hi0.copy()
        for _ in range(iters):
            mid = 0.5 * (lo + hi)
            z = (mid[None, :] - mu) / safe
            cdf = np.where(sd_s > 0.0, ndtr(z), (mid[None, :] >= mu).astype(np.float64)).mean(axis=0)
            below = cdf < q
            lo = np.where(below, mid, lo)
            hi = np.where(below, hi, mid)
        Q[qi] = 0.5 * (lo + hi)
    return mean, sd, Q


def predict_chain(gp: "_gp.GP", X, y, chain_theta, Xtest, sigma2=0.0, jitter: float = 0.0, obs_var=None,
                  quantiles=(0.05, 0.5, 0.95), ctx=None):
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
    mu = np.empty((S, Xt.shape[0]))
    var = np.empty_like(mu)
    for s in range(S):
        fx = _gp.FiniteGP(gp, X, float(s2[s]), theta=Th[s], jitter=jitter, ctx=ctx)
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


def log_evidence_harmonic(loglik_chain) -> float:
    """log of the harmonic-mean evidence estimate of a chain of log-likelihoods, as a stable log-sum-exp:
    log Z = log S - logsumexp(-ll).  (The reference reaches for BigFloat powers instead: CLI/src/select.jl:15-20.)"""
    ll = np.asarray(loglik_chain, dtype=np.float64)
    m = np.max(-ll)
    return float(np.log(ll.size) - (m + np.log(np.sum(np.exp(-ll - m)))))


def log_bayes_factor(loglik_chain_1, loglik_chain_2) -> float:
    """log10 Bayes factor of model 1 over model 2 from two chains (`select --chains`)."""
    return (log_evidence_harmonic(loglik_chain_1) - log_evidence_harmonic(loglik_chain_2)) / np.log(10.0)
