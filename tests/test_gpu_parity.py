"""Parity of the CUDA path (through the C ABI) with the CPU oracle, on identical inputs.

Tolerances (north_star): lml 1e-9 relative; posterior mean/variance 1e-8.  Covariance entries are exp/abs/mul
only and agree to a few ulp.  Every test here needs a B200 (pytest -m gpu)."""
import numpy as np
import pytest

from gaplac_b200 import _lib, workloads as W
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
from oracle import c_oracle as CO
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

LML_RTOL = 1e-9
PRED_TOL = 1e-8

ALL_KINDS = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
             Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
             Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
THETA = np.array([1.3, 0.8, 0.4, 1.7, 0.2])


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    return X, rng.standard_normal(n)


# ---------------------------------------------------------------------------------------------- covariance
@pytest.mark.parametrize("n", [1, 2, 50, 63, 64, 65, 129, 300])
def test_cov_matches_oracle(ctx, n):
    X, _ = _data(n, seed=n)
    prog = ctx.program(ALL_KINDS)
    K = ctx.cov(prog, X, THETA, 0.1, 1e-9)
    Kref = CO.cov(ALL_KINDS, X, THETA, 0.1, 1e-9)
    assert K.shape == (n, n)
    assert np.max(np.abs(K - Kref) / np.maximum(1.0, np.abs(Kref))) < 1e-14


def test_cross_cov_matches_oracle_and_drops_noise(ctx):
    X, _ = _data(70, seed=1)
    Xs, _ = _data(133, seed=2)
    Xs[:5] = X[:5]                                      # identical points: Noise must still contribute 0
    prog = ctx.program(ALL_KINDS)
    Ks = ctx.cross_cov(prog, X, Xs, THETA)
    ref = O.eval_program(ALL_KINDS, X, Xs, THETA, same=False)
    assert np.max(np.abs(Ks - ref)) < 1e-14


# ---------------------------------------------------------------------------------------------- lml
@pytest.mark.parametrize("n", [1, 7, 50, 64, 65, 200, 300])
def test_lml_single_matches_oracle(ctx, n):
    X, y = _data(n, seed=10 + n)
    prog = ctx.program(ALL_KINDS)
    lml, info = ctx.lml_batched(prog, X, y, THETA[None, :], 0.1)
    ref, rinfo = CO.lml(ALL_KINDS, X, y, THETA, 0.1)
    assert info[0] == 0 and rinfo == 0
    assert abs(lml[0] - ref) <= LML_RTOL * abs(ref)


@pytest.mark.parametrize("n,B", [(1100, 3), (1300, 2)])
def test_lml_batched_beyond_the_shared_memory_z_window(ctx, n, B):
    """n > 1088: the diagonal phase reads z from the global workspace instead of its shared-memory copy, and the
    factor workspace of the lockstep schedule spans > 17 tile columns."""
    X, y = _data(n, seed=70 + n)
    prog = ctx.program(ALL_KINDS)
    Th = np.vstack([THETA * (1.0 + 0.05 * b) for b in range(B)])
    lml, info = ctx.lml_batched(prog, X, y, Th, 0.1)
    for b in range(B):
        ref, rinfo = CO.lml(ALL_KINDS, X, y, Th[b], 0.1)
        assert info[b] == 0 and rinfo == 0
        assert abs(lml[b] - ref) <= LML_RTOL * abs(ref)
    lml_g, info_g, dth, dy = ctx.lml_batched(prog, X, y, Th[:1], 0.1, grad=True)   # fused gradient kernel, same n
    assert abs(lml_g[0] - lml[0]) <= LML_RTOL * abs(lml[0])


def test_lml_batched_workspace_chunking_is_invisible(ctx):
    """A batch whose factor workspace exceeds the cap runs in chunks: same bits as in one piece."""
    d = W.make_c3(features=300)
    prog = ctx.program(d["ops"])
    whole, info = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0)
    ctx.set_option("lk_ws_limit_mb", 24)          # 15 tiles x 32 KiB per item: ~50 items per chunk
    try:
        parts, info2 = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0)
    finally:
        ctx.set_option("lk_ws_limit_mb", 24 * 1024)
    assert np.array_equal(whole, parts) and np.array_equal(info, info2)


def test_lml_c1_readme_shape(ctx):
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    th = np.array([[0.5], [1.0], [1.5], [7.0], [19.9]])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], th, d["sigma2"])
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], th, d["sigma2"])
    assert not info.any()
    assert np.max(np.abs(lml - ref) / np.abs(ref)) < LML_RTOL


def test_lml_c2_theta_batch(ctx):
    d = W.make_c2(n=512, B=48)
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0)
    ref, rinfo = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"], 0.0)
    assert not info.any() and not rinfo.any()
    assert np.max(np.abs(lml - ref) / np.abs(ref)) < LML_RTOL
    # the reference's own pairwise-distance expansion [upstream Distances] stays inside the tolerance too
    ref2, _ = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"][:8], 0.0, mode=1)
    assert np.max(np.abs(lml[:8] - ref2) / np.abs(ref2)) < LML_RTOL


def test_lml_c3_y_batch(ctx):
    d = W.make_c3(features=40)
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0)
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["Y"], d["Theta"], 0.0)
    assert not info.any()
    assert np.max(np.abs(lml - ref) / np.abs(ref)) < LML_RTOL


def test_lml_x_batched_and_sigma_batched(ctx):
    B, n = 6, 90
    rng = np.random.default_rng(3)
    Xb = np.stack([_data(n, seed=s)[0] for s in range(B)])
    Yb = rng.standard_normal((B, n))
    Th = THETA[None, :] * rng.uniform(0.8, 1.2, (B, 5))
    s2 = rng.uniform(0.05, 0.3, B)
    prog = ctx.program(ALL_KINDS)
    lml, info = ctx.lml_batched(prog, Xb, Yb, Th, s2)
    ref, _ = CO.lml_batched(ALL_KINDS, Xb, Yb, Th, s2, x_batched=True)
    assert not info.any()
    assert np.max(np.abs(lml - ref) / np.abs(ref)) < LML_RTOL


@pytest.mark.parametrize("tag", ["3206", "1003"])
def test_lml_golden_fixture_rows(ctx, tag, golden_dir):
    """The reference's own known answers: all 100 chain rows in one batch at n = 923."""
    X, y, Th, s2, lpi, prior = O.load_golden(tag, golden_dir)
    prog = ctx.program(O.golden_program(tag))
    lml, info = ctx.lml_batched(prog, X, y, Th, s2, jitter=O.GOLDEN_JITTER)
    assert not info.any()
    assert np.max(np.abs(lml + prior - lpi)) < 1e-9 * np.max(np.abs(lpi))
    assert np.max(np.abs(lml + prior - lpi)) < 1e-9          # observed ~1e-11 absolute


def test_lml_not_positive_definite_is_reported_per_item(ctx):
    X = np.array([[0.0], [0.0], [1.0], [2.0]])
    ops = [Op(SQEXP, col=0, theta_slot=0)]
    prog = ctx.program(ops)
    lml, info = ctx.lml_batched(prog, X, np.zeros(4), np.array([[1.0], [2.0]]), np.array([0.0, 0.5]))
    assert info[0] == 2 and lml[0] == -np.inf               # duplicated point, no noise: pivot 2 (LAPACK convention)
    assert info[1] == 0 and np.isfinite(lml[1])
    ref, rinfo = CO.lml_batched(ops, X, np.zeros(4), np.array([[1.0], [2.0]]), np.array([0.0, 0.5]))
    assert list(rinfo) == [2, 0] and abs(lml[1] - ref[1]) < LML_RTOL * abs(ref[1])


def test_lml_nan_hyperparameter_does_not_poison_the_batch(ctx):
    d = W.make_c2(n=100, B=4)
    Th = d["Theta"].copy()
    Th[1, 0] = np.nan
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], Th, 0.0)
    assert info[1] != 0 and lml[1] == -np.inf
    ok = [0, 2, 3]
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], Th[ok], 0.0)
    assert np.max(np.abs(lml[ok] - ref) / np.abs(ref)) < LML_RTOL


def test_lml_duplicate_items_are_bitwise_identical_and_order_free(ctx):
    d = W.make_c2(n=192, B=20)
    Th = np.vstack([d["Theta"], d["Theta"][::-1]])
    prog = ctx.program(d["ops"])
    lml, _ = ctx.lml_batched(prog, d["X"], d["y"], Th, 0.0)
    assert np.array_equal(lml[:20], lml[20:][::-1])


# ---------------------------------------------------------------------------------------------- gradient
@pytest.mark.parametrize("n", [20, 64, 150])
def test_gradient_matches_oracle(ctx, n):
    X, y = _data(n, seed=20 + n)
    prog = ctx.program(ALL_KINDS)
    Th = np.vstack([THETA, THETA * 1.1])
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, Th, 0.1, grad=True)
    for b in range(2):
        val, rdth, rdy = O.lml_grad(ALL_KINDS, X, y, Th[b], 0.1)
        assert abs(lml[b] - val) < LML_RTOL * abs(val)
        assert np.max(np.abs(dth[b] - rdth) / np.maximum(1.0, np.abs(rdth))) < 1e-8
        assert np.max(np.abs(dy[b] - rdy)) < 1e-8 * max(1.0, np.max(np.abs(rdy)))


def test_gradient_mcmc_model_body(ctx):
    """One log-density + gradient evaluation of CLI/src/mcmc.jl:31-37 (shared l for every --infer variable)."""
    d = W.make_c1()
    fx = np.random.default_rng(0).standard_normal(50)
    prog = ctx.program(d["ops"])
    lml, info, dth, dy = ctx.lml_batched(prog, d["X"], fx, np.array([[2.5]]), 0.1, grad=True)
    val, rdth, rdfx = O.mcmc_logjoint(d["ops"], d["X"], d["y"], [2.5], fx)
    got = lml[0] - np.log(20.0) - 0.5 * 50 * O.LOG2PI - 0.5 * np.sum((d["y"] - fx) ** 2)
    assert abs(got - val) < LML_RTOL * abs(val)
    assert abs(dth[0, 0] - rdth[0]) < 1e-8 * max(1.0, abs(rdth[0]))
    assert np.max(np.abs(dy[0] + (d["y"] - fx) - rdfx)) < 1e-8


# ---------------------------------------------------------------------------------------------- posterior / predict / sample
@pytest.mark.parametrize("n,variant", [(50, 0), (130, 0), (300, 0), (130, 1), (700, 0), (700, 4), (1000, 0), (1000, 3)])
def test_posterior_and_predict_match_oracle(ctx, n, variant):
    X, y = _data(n, seed=30 + n)
    Xs, _ = _data(150, seed=77)
    prog = ctx.program(ALL_KINDS)
    ctx.set_option("chol_variant", variant)                 # 1: force the multi-CTA large-n path
    try:
        post = ctx.posterior_fit(prog, X, y, THETA, 0.1)
    finally:
        ctx.set_option("chol_variant", 0)
    U, alpha = CO.posterior(ALL_KINDS, X, y, THETA, 0.1)
    ref_lml, _ = CO.lml(ALL_KINDS, X, y, THETA, 0.1)
    assert abs(post.logpdf() - ref_lml) < LML_RTOL * abs(ref_lml)
    assert np.max(np.abs(post.alpha() - alpha)) < PRED_TOL * max(1.0, np.max(np.abs(alpha)))
    Ug = post.factor()
    assert np.max(np.abs(Ug - U)) < 1e-10 * np.max(np.abs(U))
    assert np.all(np.tril(Ug, -1) == 0.0)
    mean, var = post.mean_and_var(Xs)
    rmean, rvar = CO.mean_and_var(ALL_KINDS, X, U, alpha, Xs, THETA)
    assert np.max(np.abs(mean - rmean)) < PRED_TOL * max(1.0, np.max(np.abs(rmean)))
    assert np.max(np.abs(var - rvar)) < PRED_TOL * max(1.0, np.max(np.abs(rvar)))
    assert np.max(np.abs(post.mean_and_var(Xs, want_var=False) - rmean)) < PRED_TOL * max(1.0, np.max(np.abs(rmean)))
    post.free()


@pytest.mark.parametrize("m", [1, 33, 9472, 12000, 14000])
def test_predict_every_slab_width(ctx, m):
    """The host picks slabs of 64, 48 or 32 test points by SM balance (9472 -> 64, 12000 and 14000 -> 48, small -> 32):
    all widths, ragged last slabs included, must agree with the oracle."""
    X, y = _data(130, seed=5)
    prog = ctx.program(ALL_KINDS)
    post = ctx.posterior_fit(prog, X, y, THETA, 0.1)
    rng = np.random.default_rng(m)
    Xs = np.column_stack([rng.uniform(-3, 3, m), rng.uniform(0, 5, m), rng.standard_normal(m),
                          rng.integers(0, 4, m).astype(float)])
    mean, var = post.mean_and_var(Xs)
    idx = np.unique(np.concatenate([np.arange(min(m, 70)), np.arange(max(0, m - 70), m), rng.integers(0, m, 100)]))
    U, alpha = CO.posterior(ALL_KINDS, X, y, THETA, 0.1)
    rm, rv = CO.mean_and_var(ALL_KINDS, X, U, alpha, np.ascontiguousarray(Xs[idx]), THETA)
    assert np.max(np.abs(mean[idx] - rm)) < PRED_TOL * max(1.0, np.max(np.abs(rm)))
    assert np.max(np.abs(var[idx] - rv)) < PRED_TOL * max(1.0, np.max(np.abs(rv)))
    assert np.all(np.isfinite(mean)) and np.all(np.isfinite(var))
    post.free()


def test_posterior_plot_call_sequence(ctx):
    """src/plotting.jl:6-12: FiniteGP(gp, x, 0.1) -> posterior -> mean_and_var at 100 grid points."""
    import gaplac_b200 as G
    d = W.make_c1()
    gp, _ = G.make_gp(G.gp_spec("y ~| SqExp(:x; l=1.5)"))
    x = d["X"][:, 0]
    fx = G.FiniteGP(gp, x, 0.1, ctx=ctx)
    pgp = G.posterior(fx, d["y"])
    xtest = np.linspace(x.min() - 1, x.max() + 1, 100)
    ym, yvar = G.mean_and_var(pgp, xtest)
    ops = [Op(SQEXP, col=0, value=1.5)]
    U, alpha = CO.posterior(ops, d["X"], d["y"], [], 0.1)
    rm, rv = CO.mean_and_var(ops, d["X"], U, alpha, xtest[:, None], [])
    assert np.max(np.abs(ym - rm)) < PRED_TOL and np.max(np.abs(yvar - rv)) < PRED_TOL
    assert abs(G.logpdf(fx, d["y"]) - CO.lml(ops, d["X"], d["y"], [], 0.1)[0]) < LML_RTOL * 100


def test_posdef_exception_mirrors_julia(ctx):
    import gaplac_b200 as G
    gp = G.GP(G.kernel(G.SqExp("x"))[0])
    fx = G.FiniteGP(gp, np.array([0.0, 0.0, 1.0]), 0.0, ctx=ctx)
    with pytest.raises(G.PosDefException) as e:
        G.logpdf(fx, np.zeros(3))
    assert e.value.info == 2
    with pytest.raises(G.PosDefException):
        G.posterior(fx, np.zeros(3))


@pytest.mark.parametrize("n", [50, 200])
def test_sample_matches_oracle(ctx, n):
    X, _ = _data(n, seed=40 + n)
    Z = np.random.default_rng(5).standard_normal((n, 5))
    prog = ctx.program(ALL_KINDS)
    out = ctx.sample(prog, X, THETA, 0.1, Z)
    ref = CO.sample(ALL_KINDS, X, THETA, 0.1, Z)
    assert np.max(np.abs(out - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))


# ---------------------------------------------------------------------------------------------- large n
@pytest.mark.parametrize("n", [100, 257, 1000])
def test_chol_logdet_matches_oracle(ctx, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    U, ld, info = ctx.chol_logdet(A)
    Uref, ldref, rc = CO.chol_logdet(A)
    assert info == 0 and rc == 0
    assert abs(ld - ldref) < 1e-10 * max(1.0, abs(ldref))
    assert np.max(np.abs(U - Uref)) < 1e-10 * np.max(np.abs(Uref))
    assert np.max(np.abs(U.T @ U - A)) < 1e-11 * np.max(np.abs(A)) * n
    _, ld2, _ = ctx.chol_logdet(A, want_factor=False)
    assert ld2 == ld


def test_chol_logdet_reports_failing_pivot(ctx):
    A = np.eye(200)
    A[130, 130] = -1.0
    _, ld, info = ctx.chol_logdet(A)
    assert info == 131 and np.isnan(ld)


def test_lml_large_matches_oracle(ctx):
    d = W.make_c5(n=1500)
    prog = ctx.program(d["ops"])
    lml, ld, info = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ref, rinfo = CO.lml(d["ops"], d["X"], d["y"], d["theta"], 0.0)
    assert info == 0 and rinfo == 0
    assert abs(lml - ref) < LML_RTOL * abs(ref)


# ---------------------------------------------------------------------------------------------- handle lifetimes
def test_posterior_outliving_its_context_is_detached_not_dangling():
    """gpl_destroy with a live posterior: the device block goes with the context, the handle stays valid for
    gpl_posterior_free and every other call on it fails cleanly (no use-after-free)."""
    c = _lib.Context(0)
    X, y = _data(80, seed=1)
    prog = c.program(ALL_KINDS)
    post = c.posterior_fit(prog, X, y, THETA, 0.1)
    keep = post.logpdf()
    c.close()
    assert post.logpdf() == keep                      # host-side value
    with pytest.raises(_lib.GaplacError):
        post.alpha()
    with pytest.raises(_lib.GaplacError):
        post.mean_and_var(X[:5])
    post.free()
    post.free()                                       # idempotent


def test_predict_batched_runs_in_passes_beyond_the_workspace_cap(ctx):
    """More chain rows than the factor workspace holds: several passes inside the call, same bits."""
    X, y = _data(200, seed=9)
    prog = ctx.program(ALL_KINDS)
    rng = np.random.default_rng(1)
    Th = THETA[None, :] * rng.uniform(0.8, 1.2, (40, 5))
    Xs, _ = _data(70, seed=11)
    whole = ctx.predict_batched(prog, X, y, Th, 0.1, Xs)
    ctx.set_option("lk_ws_limit_mb", 6)               # 10 tiles + 4 inverses per row: ~ 12 rows per pass
    try:
        parts = ctx.predict_batched(prog, X, y, Th, 0.1, Xs)
    finally:
        ctx.set_option("lk_ws_limit_mb", 24 * 1024)
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b)
    U, alpha = CO.posterior(ALL_KINDS, X, y, Th[33], 0.1)
    rm, rv = CO.mean_and_var(ALL_KINDS, X, U, alpha, Xs, Th[33])
    assert np.max(np.abs(parts[0][33] - rm)) < PRED_TOL * max(1.0, np.max(np.abs(rm)))
    assert np.max(np.abs(parts[1][33] - rv)) < PRED_TOL * max(1.0, np.max(np.abs(rv)))


# ---------------------------------------------------------------------------------------------- structurally zero tiles
def test_zero_tile_skipping_keeps_every_bit():
    """Block-diagonal covariances (Cat(:subject) * SqExp(:time) on rows grouped by subject: C3) leave most tiles of L exactly
    zero; the lockstep kernels skip the updates with them, their triangular solves and their covariance code.  Same bits as
    the dense walk - lml, gradient and predictions, including an item with a poisoned length scale and one without noise."""
    d = W.make_c3(features=24)
    X, Y = d["X"], d["Y"]
    c = _lib.Context(0)
    try:
        prog = c.program(d["ops"])
        Theta = d["Theta"].copy()
        Theta[3, 0] = -1.0   # a poisoned length scale: that item is NaN / not positive definite either way
        Theta[5, 1] = 0.0    # no noise: K is singular for tied categories or fine otherwise - same answer either way
        on = c.lml_batched(prog, X, Y, Theta, 0.0, grad=True)
        c.set_option("zero_tile_skip", 0)
        off = c.lml_batched(prog, X, Y, Theta, 0.0, grad=True)
        for a, b in zip(on, off):
            assert np.array_equal(a, b, equal_nan=True)
        good = np.ones(len(Theta), bool)
        good[[3, 5]] = False
        ref, _ = CO.lml_batched(d["ops"], X, Y[good], Theta[good], 0.0)
        assert np.max(np.abs(on[0][good] - ref) / np.abs(ref)) < LML_RTOL
        # the posterior path (keeps the factor) walks the same kernels
        c.set_option("zero_tile_skip", 1)
        m1, v1, l1, i1 = c.predict_batched(prog, X, Y[0], Theta[:3], 0.0, X[:40] + 0.25)
        c.set_option("zero_tile_skip", 0)
        m0, v0, l0, i0 = c.predict_batched(prog, X, Y[0], Theta[:3], 0.0, X[:40] + 0.25)
        assert np.array_equal(m1, m0) and np.array_equal(v1, v0) and np.array_equal(l1, l0)
    finally:
        c.close()


def test_rows_are_grouped_by_the_shared_category_column():
    """Rows in any order, option zero_tile_skip = 3: a program whose every term carries Cat on one column has its observations
    grouped by that column inside the call (lml and dtheta do not depend on the order; dy comes back in the caller's order)."""
    d = W.make_c3(features=12)
    perm = np.random.default_rng(0).permutation(d["X"].shape[0])
    X, Y = d["X"][perm], d["Y"][:, perm]
    c = _lib.Context(0)
    try:
        prog = c.program(d["ops"])
        c.set_option("zero_tile_skip", 3)
        lml, info, dth, dy = c.lml_batched(prog, X, Y, d["Theta"], 0.0, grad=True)
        c.set_option("zero_tile_skip", 0)  # no grouping, no flags: the dense walk over the shuffled rows
        lml0, info0, dth0, dy0 = c.lml_batched(prog, X, Y, d["Theta"], 0.0, grad=True)
        assert not info.any() and not info0.any()
        assert np.max(np.abs(lml - lml0) / np.abs(lml0)) < 1e-12
        assert np.max(np.abs(dth - dth0) / np.maximum(1.0, np.abs(dth0))) < 1e-9
        assert np.max(np.abs(dy - dy0)) < 1e-9 * max(1.0, np.max(np.abs(dy0)))
        ref, _ = CO.lml_batched(d["ops"], X, Y, d["Theta"], 0.0)
        assert np.max(np.abs(lml - ref) / np.abs(ref)) < LML_RTOL
    finally:
        c.close()


def test_zero_tile_flags_next_to_the_separable_ou_sort():
    """Cat(:subject) * OU(:time) + Noise at n = 300 qualifies for the zero-tile flags and for the separable OU form, whose sort by
    time would interleave the subjects.  With the flags on the rows keep the caller's (grouped) order and the OU sort is not
    used; with the flags off it is.  Every combination against the oracle; flags on / off bitwise equal without the sort."""
    d = W.make_c3(features=8)
    ops = [Op(CAT, col=0), Op(OU, col=1, theta_slot=0), Op(MUL), Op(NOISE, var_slot=1), Op(ADD)]
    Theta = np.column_stack([np.linspace(20, 90, 8), np.linspace(0.1, 0.4, 8)])
    c = _lib.Context(0)
    try:
        prog = c.program(ops)
        res = {}
        for skip in (1, 0):
            for sep in (1, 0):
                c.set_option("zero_tile_skip", skip)
                c.set_option("ou_separable", sep)
                res[(skip, sep)] = c.lml_batched(prog, d["X"], d["Y"], Theta, 0.0, grad=True)
        for a, b in zip(res[(1, 0)], res[(0, 0)]):
            assert np.array_equal(a, b)
        for a, b in zip(res[(1, 1)], res[(1, 0)]):  # flags on: the separable-OU option changes nothing for this program
            assert np.array_equal(a, b)
        ref, _ = CO.lml_batched(ops, d["X"], d["Y"], Theta, 0.0)
        for k, r in res.items():
            assert not r[1].any()
            assert np.max(np.abs(r[0] - ref) / np.abs(ref)) < LML_RTOL, k
        assert np.max(np.abs(res[(1, 1)][2] - res[(0, 0)][2]) / np.maximum(1.0, np.abs(res[(0, 0)][2]))) < 1e-8
    finally:
        c.close()
