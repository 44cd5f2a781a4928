"""Seeded differential fuzzing of the kernel-program interpreter: random formulas (sums of products of leaves, variance
slots, shared slots, constants, noise) and ragged sizes, GPU against the NumPy oracle - covariance entries, lml,
analytic gradient.  Every case is deterministic (seed = case index)."""
import numpy as np
import pytest

from gaplac_b200 import _lib
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
from gaplac_b200.formula import Op
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
P_SLOTS = 6


def _random_program(rng):
    """Postfix program: 1-3 terms, each a product of 1-3 leaves; hyperparameters fixed or in slots 0..P_SLOTS-1
    (slots may repeat: one draw can drive several leaves, as the mcmc model body does)."""
    ops = []
    n_terms = int(rng.integers(1, 4))
    for t in range(n_terms):
        n_leaves = int(rng.integers(1, 4))
        for l in range(n_leaves):
            kind = [SQEXP, OU, LINEAR, CAT][int(rng.integers(0, 4))]
            col = {SQEXP: 0, OU: 1, LINEAR: 2, CAT: 3}[kind]
            kw = dict(col=col)
            if kind != CAT:
                if rng.random() < 0.6:
                    kw["theta_slot"] = int(rng.integers(0, P_SLOTS))
                else:
                    kw["value"] = float(rng.uniform(0.5, 2.0))
            if rng.random() < 0.3:
                kw["var_slot"] = int(rng.integers(0, P_SLOTS))
            elif rng.random() < 0.3:
                kw["var"] = float(rng.uniform(0.3, 1.5))
            ops.append(Op(kind, **kw))
            if l > 0:
                ops.append(Op(MUL))
        if t > 0:
            ops.append(Op(ADD))
    extra = rng.random()
    if extra < 0.4:
        ops += [Op(NOISE, var_slot=int(rng.integers(0, P_SLOTS))), Op(ADD)]
    elif extra < 0.6:
        ops += [Op(CONSTANT, theta_slot=int(rng.integers(0, P_SLOTS))), Op(ADD)]
    elif extra < 0.7:
        ops += [Op(CONSTANT, value=0.25), Op(ADD)]
    return ops


@pytest.fixture(scope="module")
def ctx():
    return _lib.Context(0)


@pytest.mark.parametrize("case", range(40))
def test_random_program_against_oracle(ctx, case):
    rng = np.random.default_rng(1000 + case)
    ops = _random_program(rng)
    n = int(rng.choice([1, 3, 17, 63, 64, 65, 100, 129, 200]))
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    y = rng.standard_normal(n)
    theta = rng.uniform(0.4, 1.6, P_SLOTS)
    sigma2 = float(rng.uniform(0.3, 1.0))           # keeps K_y positive definite whatever the formula
    try:
        prog = ctx.program(ops)
    except _lib.GaplacError as e:                   # expansion limits (terms / factors): a documented API error
        assert "more than" in str(e) or "limit" in str(e).lower()
        return
    K = ctx.cov(prog, X, theta, sigma2, 1e-10)
    Kref = O.cov(ops, X, theta, sigma2, 1e-10)
    assert np.max(np.abs(K - Kref) / np.maximum(1.0, np.abs(Kref))) < 1e-13
    ref, rdth, rdy = O.lml_grad(ops, X, y, theta, sigma2, 1e-10)
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, theta[None, :], sigma2, 1e-10, grad=True)
    assert info[0] == 0
    assert abs(lml[0] - ref) <= 1e-9 * abs(ref)
    lml2, info2 = ctx.lml_batched(prog, X, y, np.vstack([theta, theta]), sigma2, 1e-10)        # lockstep path
    assert abs(lml2[0] - ref) <= 1e-9 * abs(ref) and lml2[0] == lml2[1]
    assert np.max(np.abs(dth[0] - rdth) / np.maximum(1.0, np.abs(rdth))) < 1e-8
    assert np.max(np.abs(dy[0] - rdy)) < 1e-8 * max(1.0, np.max(np.abs(rdy)))


@pytest.mark.parametrize("case", range(0, 40, 3))
def test_random_program_posterior_and_predictions(ctx, case):
    """Same random formulas through posterior_fit / mean_and_var and the batched chain path."""
    rng = np.random.default_rng(1000 + case)
    ops = _random_program(rng)
    n = int(rng.choice([5, 64, 100, 150]))
    m = int(rng.choice([1, 40, 100]))
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    Xs = np.column_stack([rng.uniform(-3, 3, m), rng.uniform(0, 5, m), rng.standard_normal(m),
                          rng.integers(0, 4, m).astype(float)])
    y = rng.standard_normal(n)
    Th = rng.uniform(0.4, 1.6, (3, P_SLOTS))
    sigma2 = float(rng.uniform(0.3, 1.0))
    try:
        prog = ctx.program(ops)
    except _lib.GaplacError:
        return
    post = ctx.posterior_fit(prog, X, y, Th[0], sigma2, 1e-10)
    mean, var = post.mean_and_var(Xs)
    post.free()
    bm, bv, blml, binfo = ctx.predict_batched(prog, X, y, Th, sigma2, Xs, 1e-10)
    assert not binfo.any()
    for b in range(3):
        U, alpha = O.posterior(ops, X, y, Th[b], sigma2, 1e-10)
        rm, rv = O.mean_and_var(ops, X, U, alpha, Xs, Th[b])
        tol_m, tol_v = 1e-8 * max(1.0, np.max(np.abs(rm))), 1e-8 * max(1.0, np.max(np.abs(rv)))
        assert np.max(np.abs(bm[b] - rm)) < tol_m and np.max(np.abs(bv[b] - rv)) < tol_v
        if b == 0:
            assert np.max(np.abs(mean - rm)) < tol_m and np.max(np.abs(var - rv)) < tol_v


@pytest.mark.parametrize("case", range(16))
def test_random_program_on_the_sorted_separable_path(ctx, case):
    """n > 192 with at least one OU leaf: the library sorts the observations by the OU column and evaluates those leaves in
    separable form below the diagonal (option ou_separable, on by default).  Same random formulas, several tile columns,
    ties in the sorted column, per-item responses; lml, dtheta and dy (in the caller's order) against the oracle."""
    rng = np.random.default_rng(5000 + case)
    ops = _random_program(rng)
    if not any(o.kind == OU for o in ops):
        ops += [Op(OU, col=1, theta_slot=int(rng.integers(0, P_SLOTS)), var=0.5), Op(ADD)]
    n = int(rng.choice([193, 200, 256, 257, 330]))
    X = np.column_stack([rng.uniform(-3, 3, n), np.round(rng.uniform(0, 5, n), 2), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    Y = rng.standard_normal((3, n))
    Th = rng.uniform(0.4, 1.6, (3, P_SLOTS))
    sigma2 = float(rng.uniform(0.3, 1.0))
    try:
        prog = ctx.program(ops)
    except _lib.GaplacError:
        return
    lml, info, dth, dy = ctx.lml_batched(prog, X, Y, Th, sigma2, 1e-10, grad=True)
    assert not info.any()
    for b in range(3):
        ref, rdth, rdy = O.lml_grad(ops, X, Y[b], Th[b], sigma2, 1e-10)
        assert abs(lml[b] - ref) <= 1e-9 * abs(ref)
        assert np.max(np.abs(dth[b] - rdth) / np.maximum(1.0, np.abs(rdth))) < 1e-8
        assert np.max(np.abs(dy[b] - rdy)) < 1e-8 * max(1.0, np.max(np.abs(rdy)))
