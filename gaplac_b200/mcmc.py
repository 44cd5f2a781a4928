"""Batched on-device sampler for the `mcmc` command's model (CLI/src/mcmc.jl:31-41): host-side mirror.

The reference runs `sample(m, NUTS(0.65), N)` on ONE chain, serially, every gradient through ForwardDiff.  Here B
independent chains (one per feature of a microbiome table, or several chains of one model) advance in lockstep on the
GPU: gpl_mcmc_nuts keeps positions, momenta, trees, accept/reject decisions and the warm-up adaptation on the device; per
leapfrog step it launches one batched log-density + analytic-gradient evaluation for all chains and one state-machine
kernel (one warp per chain).  The host only replays a captured CUDA graph and polls a "chains done" counter.

`nuts(...)` returns the chains; `chain_table(...)` lays one chain out like the reference's output table
(CLI/src/mcmc.jl:42 -> src/utils.jl:30-40: one row per draw, parameter columns and `lp`)."""
from __future__ import annotations

import ctypes as C

import numpy as np

MC_MAX_P = 16


class McmcConfig(C.Structure):
    """struct McmcConfig of csrc/mcmc_core.h (internal; mirrored here for the host-compiled test harness)."""
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("dim", C.c_int), ("latent", C.c_int), ("max_depth", C.c_int),
                ("n_samples", C.c_int), ("n_adapt", C.c_int), ("search_eps", C.c_int), ("adapt_mass", C.c_int),
                ("record_warmup", C.c_int), ("record_q", C.c_int), ("w_enabled", C.c_int), ("w_init_buffer", C.c_int),
                ("w_term_buffer", C.c_int), ("w_base", C.c_int), ("delta", C.c_double), ("max_dh", C.c_double),
                ("obs_sd", C.c_double), ("eps0", C.c_double), ("seed", C.c_ulonglong), ("lo", C.c_double * MC_MAX_P),
                ("hi", C.c_double * MC_MAX_P)]


def default_n_adapt(n_samples: int) -> int:
    """Turing's NUTS(0.65) default: n_adapts = min(1000, N / 2) [upstream Turing 0.21 `NUTS(δ)` constructor]."""
    return min(1000, n_samples // 2)


def make_config(n, p, latent, n_samples, n_adapt, seed, lo, hi, obs_sd=1.0, eps0=0.1, search_eps=True, adapt_mass=True,
                delta=0.65, max_depth=10, max_dh=1000.0, record_warmup=False, record_q=False) -> McmcConfig:
    c = McmcConfig()
    c.n, c.p, c.latent = int(n), int(p), int(bool(latent))
    c.dim = c.p + (c.n if latent else 0)
    c.max_depth, c.n_samples, c.n_adapt = int(max_depth), int(n_samples), int(n_adapt)
    c.search_eps, c.adapt_mass = int(bool(search_eps)), int(bool(adapt_mass))
    c.record_warmup, c.record_q = int(bool(record_warmup)), int(bool(record_q))
    c.delta, c.max_dh, c.obs_sd, c.eps0, c.seed = float(delta), float(max_dh), float(obs_sd), float(eps0), int(seed)
    for k in range(c.p):
        c.lo[k], c.hi[k] = float(lo[k]), float(hi[k])
    # Stan's window schedule (mc_setup_windows in csrc/mcmc_core.h)
    ib, tb, base, en = 75, 50, 25, 1
    if n_adapt < 20:
        ib = tb = base = en = 0
    elif ib + base + tb > n_adapt:
        ib, tb = int(0.15 * n_adapt), int(0.1 * n_adapt)
        base = n_adapt - (ib + tb)
    c.w_enabled, c.w_init_buffer, c.w_term_buffer, c.w_base = en, ib, tb, base
    return c


class GplMcmcOpts(C.Structure):
    """struct gpl_mcmc_opts (include/gaplac_b200.h)."""
    _fields_ = [("n_samples", C.c_int32), ("n_adapt", C.c_int32), ("max_depth", C.c_int32), ("latent", C.c_int32),
                ("search_eps", C.c_int32), ("adapt_mass", C.c_int32), ("record_warmup", C.c_int32),
                ("chain_offset", C.c_int32), ("delta", C.c_double), ("max_dh", C.c_double), ("obs_sd", C.c_double),
                ("eps0", C.c_double), ("seed", C.c_uint64)]


def nuts(ctx, prog, X, Y, lo, hi, sigma2=0.1, n_samples=200, n_adapt=None, seed=0, q0=None, chains=None, latent=True,
         obs_sd=1.0, jitter=0.0, eps0=0.1, search_eps=True, adapt_mass=True, delta=0.65, max_depth=10, max_dh=1000.0,
         record_warmup=False, record_q=False, chain_offset=0):
    """Run B chains on the device.  Y: (n,) one response shared by `chains` chains, or (B, n): one chain per row.
    lo, hi: (p,) Uniform prior bounds of the hyperparameter slots (the reference: l ~ Uniform(0, 20), mcmc.jl:32).
    q0: (B, dim) initial unconstrained positions (default zeros: every hyperparameter at the middle of its range,
    latent values 0).  Returns a dict of arrays with leading dimensions (B, n_rec): theta (.., p), lp, accept, eps, depth,
    n_leapfrog, divergent, optionally q (.., dim); status (B,), grad_evals, seconds."""
    from . import _lib
    import time
    lib = _lib.load()
    X = np.asarray(X, dtype=np.float64)
    Xf = _lib._fa(X, 2) if X.ndim < 3 else np.ascontiguousarray(np.transpose(X, (0, 2, 1)))
    x_batched = X.ndim == 3
    n, d = (Xf.shape[2], Xf.shape[1]) if x_batched else Xf.shape
    Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64))
    y_batched = Y.ndim == 2
    B = Y.shape[0] if y_batched else int(chains or 1)
    lo = np.ascontiguousarray(np.atleast_1d(np.asarray(lo, dtype=np.float64)))
    hi = np.ascontiguousarray(np.atleast_1d(np.asarray(hi, dtype=np.float64)))
    p = lo.size
    dim = p + (n if latent else 0)
    if n_adapt is None:
        n_adapt = default_n_adapt(n_samples)
    n_rec = n_samples + (n_adapt if record_warmup else 0)
    q0a = np.zeros((B, dim)) if q0 is None else np.ascontiguousarray(np.broadcast_to(np.asarray(q0, dtype=np.float64), (B, dim)))
    s2 = np.ascontiguousarray(np.atleast_1d(np.asarray(sigma2, dtype=np.float64)))
    opts = GplMcmcOpts(n_samples, n_adapt, max_depth, int(bool(latent)), int(bool(search_eps)), int(bool(adapt_mass)),
                       int(bool(record_warmup)), int(chain_offset), delta, max_dh, obs_sd, eps0, int(seed))
    out = dict(theta=np.zeros((B, n_rec, p)), lp=np.zeros((B, n_rec)), accept=np.zeros((B, n_rec)), eps=np.zeros((B, n_rec)),
               depth=np.zeros((B, n_rec), dtype=np.int32), n_leapfrog=np.zeros((B, n_rec), dtype=np.int32),
               divergent=np.zeros((B, n_rec), dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    if record_q:
        out["q"] = np.zeros((B, n_rec, dim))
    evals = C.c_longlong(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    fn = lib.gpl_multi_mcmc_nuts if isinstance(ctx, _lib.MultiContext) else lib.gpl_mcmc_nuts   # same signature
    rc = fn(ctx.h, prog.h, n, d, vp(Xf), int(x_batched), vp(Y), int(y_batched), p, vp(lo), vp(hi), vp(s2),
                           int(s2.size > 1), C.c_double(jitter), B, vp(q0a), C.byref(opts), vp(out["theta"]), vp(out["lp"]),
                           vp(out["q"]) if record_q else None, vp(out["accept"]), vp(out["eps"]), vp(out["depth"]),
                           vp(out["n_leapfrog"]), vp(out["divergent"]), vp(out["status"]), C.byref(evals))
    ctx._check(rc)
    out["seconds"] = time.perf_counter() - t0
    out["grad_evals"] = evals.value
    out["n_adapt"], out["n_samples"] = n_adapt, n_samples
    return out


def chain_table(result, chain: int = 0, names=None):
    """One chain as the reference writes it (CLI/src/mcmc.jl:42): a dict of columns, one row per draw - the hyperparameters
    (named like the Turing parameters, default `ℓ`, `ℓ2`, ...) and `lp`; `select --chains` reads the `lp` column
    (CLI/src/select.jl:15-18)."""
    th = result["theta"][chain]
    p = th.shape[1]
    names = names or (["ℓ"] + [f"ℓ{k + 1}" for k in range(1, p)])
    cols = {"iteration": np.arange(1, th.shape[0] + 1)}
    for k in range(p):
        cols[names[k]] = th[:, k]
    cols["lp"] = result["lp"][chain]
    cols["n_steps"] = result["n_leapfrog"][chain]
    cols["acceptance_rate"] = result["accept"][chain]
    cols["tree_depth"] = result["depth"][chain]
    cols["numerical_error"] = result["divergent"][chain]
    cols["step_size"] = result["eps"][chain]
    return cols
