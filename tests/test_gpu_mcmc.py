"""The batched on-device sampler (gpl_mcmc_nuts) against the host reference sampler (oracle/nuts_ref.py) under a fixed
seed: both consume the same Philox stream, so the chains must agree draw by draw.  Hamiltonian trajectories amplify
rounding differences roughly tenfold every few transitions (tests/test_mcmc_core.py), hence: the first K transitions to
1e-8 (north_star tolerance), identical tree shapes over the first 2K, then distributional agreement on long runs.
Model of CLI/src/mcmc.jl:31-37.  Needs a B200."""
import numpy as np
import pytest

from gaplac_b200 import mcmc, workloads as W
from gaplac_b200 import chain as chain_mod
from oracle import nuts_ref as N

pytestmark = pytest.mark.gpu

TOL = 1e-8


def _compare(dev, b, ref, K, tol=TOL):
    D = min(2 * K, len(ref["lp"]))
    assert np.array_equal(dev["depth"][b][:D], ref["depth"][:D])
    assert np.array_equal(dev["n_leapfrog"][b][:D], ref["n_leapfrog"][:D])
    assert np.array_equal(dev["divergent"][b][:D].astype(bool), ref["divergent"][:D])
    worst = 0.0
    for k in ("q", "theta", "lp", "eps", "accept"):
        a, r = dev[k][b][:K], ref[k][:K]
        worst = max(worst, float(np.max(np.abs(a - r) / np.maximum(1.0, np.abs(r)))))
    assert worst < tol, worst
    return worst


def test_c1_readme_chains_match_the_reference_sampler(ctx):
    """gaplac mcmc "y ~| SqExp(:x)" on 50 points (BASELINE config[0]): l ~ Uniform(0, 20), latent fx, Y ~ N(fx, 1)."""
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    dev = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=8, n_adapt=24, seed=2024, chains=3,
                    record_warmup=True, record_q=True)
    assert not dev["status"].any()
    for b in range(3):
        model = N.Model(d["ops"], d["X"], d["y"], np.array([0.0]), np.array([20.0]), 0.1)
        ref = N.sample_chain(model, np.zeros(model.dim), 8, 24, seed=2024, chain=b)
        err = _compare(dev, b, ref, K=6)
        print(f"chain {b}: max rel err over the first 6 transitions {err:.2e}")
    # the three chains are different chains
    assert not np.array_equal(dev["theta"][0], dev["theta"][1])


def test_c3_feature_chains_match_the_reference_sampler(ctx):
    """16 features of the microbiome config (n = 300, Cat(:subject)*SqExp(:time)+Noise), one chain per feature; four of
    them replayed on the host."""
    d = W.make_c3(features=16)
    prog = ctx.program(d["ops"])
    lo, hi = np.array([0.0, 0.0]), np.array([100.0, 2.0])
    dev = mcmc.nuts(ctx, prog, d["X"], d["Y"], lo, hi, sigma2=0.0, n_samples=3, n_adapt=5, seed=11, record_warmup=True,
                    record_q=True)
    assert not dev["status"].any()
    for b in (0, 5, 10, 15):
        model = N.Model(d["ops"], d["X"], d["Y"][b], lo, hi, 0.0)
        ref = N.sample_chain(model, np.zeros(model.dim), 3, 5, seed=11, chain=b)
        err = _compare(dev, b, ref, K=4)
        print(f"feature {b}: max rel err over the first 4 transitions {err:.2e}")


def test_chain_streams_do_not_depend_on_the_batch(ctx):
    """Chain 2 of a 4-chain call == the same chain run alone with chain_offset = 2 (sharding chains over GPUs keeps them)."""
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    kw = dict(sigma2=0.1, n_samples=6, n_adapt=12, seed=5, record_warmup=True)
    four = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], chains=4, **kw)
    one = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], chains=1, chain_offset=2, **kw)
    for k in ("theta", "lp", "eps", "accept", "depth", "n_leapfrog"):
        assert np.array_equal(four[k][2], one[k][0]), k


def test_marginal_model_and_zero_density_start(ctx):
    """latent = 0 (hyperparameters only, Y observed directly) with two slots; a chain started at a NaN position reports
    status 1 and does not disturb its neighbours."""
    d = W.make_c5(n=70)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(70)
    X = d["X"] / 10.0
    prog = ctx.program(d["ops"])
    lo, hi = np.array([0.05, 0.01]), np.array([10.0, 5.0])
    q0 = np.array([[0.3, -0.2], [np.nan, 0.0], [0.3, -0.2]])
    dev = mcmc.nuts(ctx, prog, X, y, lo, hi, sigma2=0.0, n_samples=10, n_adapt=20, seed=7, q0=q0, chains=3, latent=False,
                    record_warmup=True, record_q=True)
    assert list(dev["status"]) == [0, 1, 0]
    model = N.Model(d["ops"], X, y, lo, hi, 0.0, latent=False)
    ref = N.sample_chain(model, q0[0], 10, 20, seed=7, chain=0)
    _compare(dev, 0, ref, K=8)
    ref2 = N.sample_chain(model, q0[2], 10, 20, seed=7, chain=2)
    _compare(dev, 2, ref2, K=8)


def test_long_run_distribution_agrees_with_the_reference(ctx):
    """After the trajectories have decorrelated: posterior summaries of l and lp from 16 device chains against 4 host
    chains (C1, 100 warm-up + 200 draws each), and the warm-up reaches the target acceptance."""
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    dev = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=200, n_adapt=100, seed=99, chains=16)
    assert not dev["status"].any()
    th_d, lp_d = dev["theta"][:, :, 0].ravel(), dev["lp"].ravel()
    th_r, lp_r = [], []
    for b in range(4):
        model = N.Model(d["ops"], d["X"], d["y"], np.array([0.0]), np.array([20.0]), 0.1)
        ref = N.sample_chain(model, np.zeros(model.dim), 200, 100, seed=99, chain=b)
        th_r.append(ref["theta"][100:, 0])
        lp_r.append(ref["lp"][100:])
    th_r, lp_r = np.concatenate(th_r), np.concatenate(lp_r)
    assert abs(th_d.mean() - th_r.mean()) < 0.25 * th_r.std()
    assert abs(np.log(th_d).std() - np.log(th_r).std()) < 0.25 * np.log(th_r).std()
    assert abs(lp_d.mean() - lp_r.mean()) < 0.25 * lp_r.std()
    assert 0.5 < dev["accept"].mean() < 0.95
    assert dev["divergent"].mean() < 0.05
    # the output table of one chain feeds `select --chains` (CLI/src/select.jl:15-20)
    t0, t1 = mcmc.chain_table(dev, 0), mcmc.chain_table(dev, 1)
    assert set(["lp", "ℓ"]) <= set(t0)
    assert np.isfinite(chain_mod.select_chains_log2_bayes(t0["lp"], t1["lp"]))
