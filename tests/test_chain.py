"""Chain-averaged predictions and chain evidence: host summaries on CPU, the per-row
posteriors against the oracle on the GPU."""
import os

import numpy as np
import pytest

import gaplac_b200 as G
from gaplac_b200.formula import KernelProgram
from oracle import c_oracle as CO
from oracle import gp_oracle as O


def test_mixture_summary_matches_root_finding():
    from scipy.optimize import brentq
    from scipy.special import ndtr

    rng = np.random.default_rng(1)
    mu = rng.normal(0, 2, (9, 6))
    var = rng.uniform(0.01, 3, (9, 6))
    var[:, 5] = 0.0                       # degenerate components: a mixture of point masses
    qs = (0.05, 0.25, 0.5, 0.95)
    mean, sd, Q = G.mixture_summary(mu, var, qs)
    assert np.allclose(mean, mu.mean(0), rtol=0, atol=1e-15)
    assert np.allclose(sd ** 2, (var + mu ** 2).mean(0) - mu.mean(0) ** 2, rtol=1e-12, atol=1e-14)
    for c in range(5):
        for qi, q in enumerate(qs):
            f = lambda x: ndtr((x - mu[:, c]) / np.sqrt(var[:, c])).mean() - q
            ref = brentq(f, mu[:, c].min() - 20, mu[:, c].max() + 20, xtol=1e-14, rtol=1e-15)
            assert abs(Q[qi, c] - ref) < 1e-10
    srt = np.sort(mu[:, 5])               # point masses: the quantile is an order statistic
    assert abs(Q[2, 5] - srt[4]) < 1e-9   # median of 9 atoms


def test_single_component_quantiles_are_gaussian():
    from scipy.special import ndtri
    mean, sd, Q = G.mixture_summary([[1.5, -2.0]], [[4.0, 0.25]], (0.05, 0.95))
    assert np.allclose(Q[0], np.array([1.5, -2.0]) + ndtri(0.05) * np.array([2.0, 0.5]), atol=1e-12)
    assert np.allclose(Q[1], np.array([1.5, -2.0]) + ndtri(0.95) * np.array([2.0, 0.5]), atol=1e-12)


def test_harmonic_evidence_is_stable_where_the_naive_form_overflows():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    ll = np.array([-939.2, -941.7, -938.8, -945.1, -940.0])       # the golden chains live here: exp(939) overflows
    ref = mp.log(len(ll) / sum(mp.exp(-mp.mpf(float(v))) for v in ll))
    assert abs(G.log_evidence_harmonic(ll) - float(ref)) < 1e-12 * abs(float(ref))
    ll2 = ll - 3.0
    assert abs(G.log_bayes_factor(ll, ll2) - 3.0 / np.log(10.0)) < 1e-12


def test_select_chains_reproduces_the_reference_base2_harmonic_mean():
    """CLI/src/select.jl:15-20: lp_k = log2(harmmean(BigFloat(2) .^ lp)), Bayes = lp1 - lp2 - evaluated here in 60-digit
    arithmetic exactly as written there, against the stable float64 form."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    rng = np.random.default_rng(0)
    lp1 = -939.0 + rng.normal(0, 2.0, 100)
    lp2 = -640.0 + rng.normal(0, 1.5, 100)

    def ref(lp):
        v = [mp.mpf(2) ** mp.mpf(float(x)) for x in lp]
        hm = len(v) / sum(1 / x for x in v)
        return mp.log(hm, 2)

    assert abs(G.log2_harmmean_exp2(lp1) - float(ref(lp1))) < 1e-12 * abs(float(ref(lp1)))
    bayes = mp.log(mp.mpf(2) ** ref(lp1) / mp.mpf(2) ** ref(lp2), 2)
    assert abs(G.select_chains_log2_bayes(lp1, lp2) - float(bayes)) < 1e-10
    const = np.full(7, -12.5)
    assert G.select_chains_log2_bayes(const, const - 2.0) == pytest.approx(2.0, abs=1e-13)   # constant chains: c1 - c2
    assert G.select_formulae_log2_bayes(-31.53397005887427, -35.97395926954643) == pytest.approx(4.44, abs=5e-3)  # README.md:111-117


@pytest.mark.gpu
def test_predict_chain_golden_model_matches_oracle():
    """`predict ... --mcmc mcmc_3206.tsv --at "nutrient=-5:0.5:5;PersonID=0;StoolPairs=0"` (test/pred.jl:22): per-row
    posterior mean / variance against the oracle, mixture columns from the same host summary."""
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    X, y, Th, s2, lpi, prior = O.load_golden("3206", gdir)
    ops = O.golden_program("3206")
    rows = [0, 17, 50, 99]
    nut = np.arange(-5.0, 5.0001, 0.5)
    Xs = np.column_stack([np.zeros_like(nut), np.zeros_like(nut), nut])
    gp = G.GP(KernelProgram(ops=ops, vars=["PersonID", "StoolPairs", "nutrient"], n_theta=4))
    out = G.predict_chain(gp, X, y, Th[rows], Xs, sigma2=0.0, jitter=1e-9, obs_var=Th[rows, 3])
    loop = G.predict_chain(gp, X, y, Th[rows], Xs, sigma2=0.0, jitter=1e-9, obs_var=Th[rows, 3], batched=False)
    assert np.max(np.abs(out["mu"] - loop["mu"])) < 1e-9 and np.max(np.abs(out["var"] - loop["var"])) < 1e-9
    mu = np.empty((len(rows), len(nut)))
    var = np.empty_like(mu)
    for k, r in enumerate(rows):
        U, alpha = CO.posterior(ops, X, y, Th[r], 0.0, jitter=1e-9)
        mu[k], var[k] = CO.mean_and_var(ops, X, U, alpha, Xs, Th[r])
    assert np.max(np.abs(out["mu"] - mu)) < 1e-8 * max(1.0, np.max(np.abs(mu)))
    assert np.max(np.abs(out["var"] - var)) < 1e-8 * max(1.0, np.max(np.abs(var)))
    fm, fs, fQ = G.mixture_summary(mu, var, (0.05, 0.5, 0.95))
    ym, ys, yQ = G.mixture_summary(mu, var + Th[rows, 3][:, None], (0.05, 0.5, 0.95))
    assert np.allclose(out["fmu"], fm, atol=1e-8) and np.allclose(out["ymu"], ym, atol=1e-8)
    assert np.allclose(out["yQ050"], yQ[0], atol=1e-7) and np.allclose(out["yQ950"], yQ[2], atol=1e-7)
    assert np.all(out["yQ950"] - out["yQ050"] > out["fQ950"] - out["fQ050"])   # observation noise widens the band


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,B", [(50, 7, 3), (130, 200, 5), (300, 33, 2)])
def test_predict_batched_matches_oracle_per_row(n, m, B):
    """gpl_predict_batched (lockstep factorisation of all rows + lk_post_kernel + batched prediction kernel) against the
    oracle's posterior / mean_and_var row by row: ragged n, per-row sigma2, a non-PD row reported through info."""
    from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
    from gaplac_b200.formula import Op
    ALL_KINDS = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
                 Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
                 Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
    THETA = np.array([1.3, 0.8, 0.4, 1.7, 0.2])

    def _data(k, seed=0):
        rng = np.random.default_rng(seed)
        return (np.column_stack([rng.uniform(-3, 3, k), rng.uniform(0, 5, k), rng.standard_normal(k),
                                 rng.integers(0, 4, k).astype(float)]), rng.standard_normal(k))
    ctx = G.default_context()
    X, y = _data(n, seed=40 + n)
    Xs, _ = _data(m, seed=41 + n)
    prog = ctx.program(ALL_KINDS)
    Th = np.vstack([THETA * (1.0 + 0.07 * b) for b in range(B)])
    s2 = np.linspace(0.08, 0.2, B)
    mean, var, lml, info = ctx.predict_batched(prog, X, y, Th, s2, Xs)
    assert not info.any()
    for b in range(B):
        U, alpha = CO.posterior(ALL_KINDS, X, y, Th[b], float(s2[b]))
        rm, rv = CO.mean_and_var(ALL_KINDS, X, U, alpha, Xs, Th[b])
        ref, _ = CO.lml(ALL_KINDS, X, y, Th[b], float(s2[b]))
        assert abs(lml[b] - ref) < 1e-9 * abs(ref)
        assert np.max(np.abs(mean[b] - rm)) < 1e-8 * max(1.0, np.max(np.abs(rm)))
        assert np.max(np.abs(var[b] - rv)) < 1e-8 * max(1.0, np.max(np.abs(rv)))
    Thb = Th.copy()
    Thb[1, 3] = -50.0                       # negative variance on the SqExp*OU term: row 1 is not positive definite
    mean, var, lml, info = ctx.predict_batched(prog, X, y, Thb, s2, Xs)
    if B > 1:
        assert info[1] != 0 and np.isinf(lml[1]) and np.all(np.isnan(mean[1])) and np.all(np.isnan(var[1]))
        assert info[0] == 0 and np.all(np.isfinite(mean[0]))
