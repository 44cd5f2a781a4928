"""The sampler's per-chain state machine (gaplac_b200/csrc/mcmc_core.h - the header the CUDA driver runs one warp per
chain) compiled for the HOST and driven with gradients from the CPU oracle, against the independent NumPy reference
sampler (oracle/nuts_ref.py): same Philox stream, so every transition must agree (positions, step sizes, tree depths,
acceptance statistics), warm-up included.  CPU only; the device run of the same header is tests/test_gpu_mcmc.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from gaplac_b200 import workloads as W
from gaplac_b200.mcmc import McmcConfig, make_config
from oracle import gp_oracle as O
from oracle import nuts_ref as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib():
    src = os.path.join(ROOT, "tests", "host", "mcmc_host.cpp")
    out = os.path.join(ROOT, "build", "libmcmc_host_test.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(ROOT, "gaplac_b200", "csrc", "mcmc_core.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    assert lib.mc_host_config_size() == C.sizeof(McmcConfig)      # the ctypes mirror matches the C struct
    return lib


GRAD_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int),
                      C.POINTER(C.c_double), C.POINTER(C.c_double))


def run_host(lib, model, q0, n_samples, n_adapt, seed, chain=0, **kw):
    cfg = make_config(n=len(model.Y), p=model.p, latent=model.latent, n_samples=n_samples, n_adapt=n_adapt, seed=seed,
                      lo=model.lo, hi=model.hi, obs_sd=model.obs_sd, record_warmup=True, record_q=True, **kw)
    n, p, dim, T = cfg.n, cfg.p, cfg.dim, n_samples + n_adapt

    def grad(th, y, lml, info, dth, dy):
        theta = np.ctypeslib.as_array(th, (max(p, 1),))[:p].copy()
        f = np.ctypeslib.as_array(y, (n,)).copy()
        try:
            val, g_th, g_y = O.lml_grad(model.ops, model.X, f, theta, model.sigma2, model.jitter)
            ok = np.isfinite(val)
        except Exception:
            ok = False
        if not ok:
            lml[0], info[0] = -np.inf, 1
            return
        lml[0], info[0] = val, 0
        for k in range(p):
            dth[k] = g_th[k]
        np.ctypeslib.as_array(dy, (n,))[:] = g_y

    out = dict(theta=np.zeros((T, p)), lp=np.zeros(T), accept=np.zeros(T), eps=np.zeros(T), depth=np.zeros(T, dtype=np.int32),
               n_leapfrog=np.zeros(T, dtype=np.int32), divergent=np.zeros(T, dtype=np.int32), q=np.zeros((T, dim)))
    evals = C.c_longlong(0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lib.mc_host_run.restype = C.c_int
    cb = GRAD_FN(grad)
    Y = np.ascontiguousarray(model.Y, dtype=np.float64)
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    rc = lib.mc_host_run(C.byref(cfg), chain, dp(Y), dp(q0), cb, dp(out["theta"]), dp(out["lp"]), dp(out["accept"]),
                         dp(out["eps"]), ip(out["depth"]), ip(out["n_leapfrog"]), ip(out["divergent"]), dp(out["q"]),
                         C.byref(evals))
    assert rc == T
    out["evals"] = evals.value
    return out


def assert_same_chain(a, ref, T, tol=1e-9, K=12):
    """Hamiltonian trajectories amplify rounding differences (measured here: ~10x every 4-5 transitions, 1e-15 -> 1e-9
    after ~35), so 'identical chains' means: the first K transitions agree to `tol` in every continuous quantity, and the
    discrete history (tree depths, leapfrog counts, divergences) is identical over the first 2K; the step-size / metric
    adaptation keeps the two in step far longer (first K transitions, then distributional agreement)."""
    K = min(K, T)
    D = min(2 * K, T)
    assert np.array_equal(a["depth"][:D], ref["depth"][:D])
    assert np.array_equal(a["n_leapfrog"][:D], ref["n_leapfrog"][:D])
    assert np.array_equal(a["divergent"][:D].astype(bool), ref["divergent"][:D])
    for k in ("q", "theta", "lp", "eps", "accept"):
        err = np.max(np.abs(a[k][:K] - ref[k][:K]) / np.maximum(1.0, np.abs(ref[k][:K])))
        assert err < tol, (k, err)
    # later transitions: same trees most of the time, same adapted step size to a few digits
    assert np.mean(a["depth"][:T] == ref["depth"][:T]) > 0.8
    assert abs(a["eps"][T - 1] - ref["eps"][T - 1]) < 0.05 * ref["eps"][T - 1]


def test_readme_model_chain_matches_the_reference_sampler(host_lib):
    """C1: y ~| SqExp(:x), n = 50, l ~ Uniform(0, 20), latent fx, Y ~ N(fx, 1)  (CLI/src/mcmc.jl:31-37): warm-up with
    step-size search, dual averaging and two variance windows, then sampling."""
    d = W.make_c1()
    model = N.Model(d["ops"], d["X"], d["y"], np.array([0.0]), np.array([20.0]), 0.1)
    q0 = np.zeros(model.dim)
    ref = N.sample_chain(model, q0, 30, 60, seed=2024, chain=3)
    got = run_host(host_lib, model, q0, 30, 60, seed=2024, chain=3)
    assert_same_chain(got, ref, 90)
    assert abs(got["evals"] - model.evals) <= 0.05 * model.evals   # same number of gradient evaluations (identical at first)
    assert len(np.unique(ref["depth"])) > 1                  # trees of several depths were exercised
    assert np.ptp(ref["eps"][:60]) > 0 and np.ptp(ref["eps"][60:]) == 0


def test_marginal_model_two_hyperparameters(host_lib):
    """latent = False: hyperparameters only (the legacy sampler behind the golden chains): SqExp + Noise, two slots."""
    d = W.make_c5(n=40)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(40)
    model = N.Model(d["ops"], d["X"] / 10.0, y, np.array([0.05, 0.01]), np.array([10.0, 5.0]), 0.0, latent=False)
    q0 = np.array([0.3, -0.2])
    ref = N.sample_chain(model, q0, 25, 40, seed=7)
    got = run_host(host_lib, model, q0, 25, 40, seed=7)
    assert_same_chain(got, ref, 65)


def test_fixed_step_size_no_mass_adaptation_and_divergences(host_lib):
    """search_eps = False, adapt_mass = False, and a step size large enough to produce divergent trees."""
    d = W.make_c1(n=20)
    model = N.Model(d["ops"], d["X"], d["y"], np.array([0.0]), np.array([20.0]), 0.1)
    q0 = np.zeros(model.dim)
    ref = N.sample_chain(model, q0, 15, 10, seed=5, eps0=1.7, search_eps=False, adapt_mass=False, max_dh=3.0)
    got = run_host(host_lib, model, q0, 15, 10, seed=5, eps0=1.7, search_eps=False, adapt_mass=False, max_dh=3.0)
    assert ref["divergent"].any()
    assert_same_chain(got, ref, 25)


def test_window_schedule_matches_stan_for_typical_warmups():
    """The windowed-adaptation schedule (init buffer / doubling windows / terminal buffer) for n_adapt = 1000, 250, 100."""
    for W_, ends in ((1000, [99, 149, 249, 449, 949]), (250, [99, 199]), (100, [89]), (10, [])):
        wv = N.WindowedVariance(W_, 2)
        got = []
        for t in range(W_):
            if wv.learn(np.array([float(t), 1.0])) is not None:
                got.append(t)
        assert got == ends, (W_, got)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert N.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert N.philox4x32((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert N.philox4x32((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
