"""Separable evaluation of 1-D OU leaves on sorted inputs (csrc/kfun.cuh SepCtx; option "ou_separable"): the library sorts
the observations by the OU column - logpdf does not depend on their order [upstream AbstractGPs: a permutation of a
multivariate normal] - and returns dlml/dy in the caller's order.  Forced on here whatever the batch size (value 2)."""
import numpy as np
import pytest

from gaplac_b200 import workloads as W
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, LINEAR, MUL, NOISE, OU, SQEXP
from oracle import c_oracle as CO
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def forced(ctx):
    ctx.set_option("ou_separable", 2)
    yield ctx
    ctx.set_option("ou_separable", 1)


@pytest.mark.parametrize("n", [65, 130, 200, 512, 577])
def test_c2_program_matches_oracle_and_the_direct_form(forced, n):
    ctx = forced
    d = W.make_c2(n=n, B=7)
    prog = ctx.program(d["ops"])
    lml, info, dth, dy = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    ref, rinfo = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"], 0.0)
    assert not info.any() and not rinfo.any()
    assert np.max(np.abs(lml - ref) / np.abs(ref)) < 1e-9
    for b in (0, 6):
        _, rdth, rdy = O.lml_grad(d["ops"], d["X"], d["y"], d["Theta"][b], 0.0)
        assert np.max(np.abs(dth[b] - rdth) / np.maximum(1.0, np.abs(rdth))) < 1e-8
        assert np.max(np.abs(dy[b] - rdy)) < 1e-8 * max(1.0, np.max(np.abs(rdy)))      # dy in the CALLER's order
    ctx.set_option("ou_separable", 0)
    lml0, _, dth0, dy0 = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    ctx.set_option("ou_separable", 2)
    assert np.max(np.abs(lml - lml0) / np.abs(lml0)) < 1e-12
    assert np.max(np.abs(dy - dy0)) < 1e-10 * max(1.0, np.max(np.abs(dy0)))


def test_two_ou_leaves_products_other_columns_and_per_item_responses(forced):
    """OU(col 1; l0) * SqExp(col 0) + OU(col 1; l1) * Cat(col 2) + Linear(col 0) + Noise: two OU leaves on the sorted column
    inside products, other leaves on other columns, one response per item; ties in the sorted column."""
    ctx = forced
    rng = np.random.default_rng(3)
    n, B = 150, 5
    X = np.column_stack([rng.uniform(-3, 3, n), np.round(rng.uniform(0, 10, n), 1), rng.integers(0, 3, n).astype(float)])
    ops = [Op(OU, col=1, theta_slot=0), Op(SQEXP, col=0, theta_slot=1), Op(MUL), Op(OU, col=1, theta_slot=2), Op(CAT, col=2), Op(MUL),
           Op(ADD), Op(LINEAR, col=0, value=0.5, var=0.3), Op(ADD), Op(NOISE, var_slot=3), Op(ADD)]
    Th = np.column_stack([rng.uniform(0.5, 4, B), rng.uniform(0.5, 2, B), rng.uniform(0.3, 6, B), rng.uniform(0.05, 0.4, B)])
    Y = rng.standard_normal((B, n))
    prog = ctx.program(ops)
    lml, info, dth, dy = ctx.lml_batched(prog, X, Y, Th, 0.0, grad=True)
    assert not info.any()
    for b in range(B):
        val, rdth, rdy = O.lml_grad(ops, X, Y[b], Th[b], 0.0)
        assert abs(lml[b] - val) < 1e-9 * abs(val)
        assert np.max(np.abs(dth[b] - rdth) / np.maximum(1.0, np.abs(rdth))) < 1e-8
        assert np.max(np.abs(dy[b] - rdy)) < 1e-8 * max(1.0, np.max(np.abs(rdy)))


def test_bad_items_and_tiny_length_scales(forced):
    ctx = forced
    d = W.make_c2(n=200, B=5)
    Th = d["Theta"].copy()
    Th[1, 1] = -2.0            # negative OU length scale: rejected per item
    Th[2, 1] = 1e-3            # exponents down to -1e4: deep underflow in the separable factors
    Th[3, 1] = np.nan
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], Th, 0.0)
    assert info[1] != 0 and info[3] != 0 and lml[1] == -np.inf and lml[3] == -np.inf
    ok = [0, 2, 4]
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], Th[ok], 0.0)
    assert np.max(np.abs(lml[ok] - ref) / np.abs(ref)) < 1e-9


def test_headline_batch_uses_it_by_default(ctx):
    """B = 4096 at n = 512 passes the work threshold: default settings, sampled items against the oracle."""
    d = W.make_c2(n=512, B=4096)
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0)
    idx = np.array([0, 1, 777, 2048, 4095])
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"][idx], 0.0)
    assert not info.any()
    assert np.max(np.abs(lml[idx] - ref) / np.abs(ref)) < 1e-9


def test_bits_do_not_depend_on_the_batch(ctx):
    """The separable path is chosen from the model alone (n, the program): an item evaluated alone, in a batch of 9, or in
    one part of a two-part multi-device call has the same bits; so does a sampler chain run alone or among others."""
    from gaplac_b200 import _lib, mcmc
    d = W.make_c2(n=256, B=9)
    prog = ctx.program(d["ops"])
    whole = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    alone = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"][4:5], 0.0, grad=True)
    for a, b in zip(whole, alone):
        assert np.array_equal(a[4], b[0])
    m = _lib.MultiContext([0, 0])
    try:
        parts = m.lml_batched(m.program(d["ops"]), d["X"], d["y"], d["Theta"], 0.0, grad=True)
    finally:
        m.close()
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b)
    # sampler on an OU model (marginal form: hyperparameters only), n = 256
    lo, hi = np.array([0.2, 0.2, 0.01]), np.array([5.0, 5.0, 1.0])
    kw = dict(sigma2=0.0, n_samples=3, n_adapt=4, seed=9, latent=False, record_warmup=True)
    three = mcmc.nuts(ctx, prog, d["X"], d["y"], lo, hi, chains=3, **kw)
    one = mcmc.nuts(ctx, prog, d["X"], d["y"], lo, hi, chains=1, chain_offset=1, **kw)
    assert not three["status"].any()
    for k in ("theta", "lp", "eps", "depth"):
        assert np.array_equal(three[k][1], one[k][0]), k
