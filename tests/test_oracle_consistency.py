"""The two oracle restatements (NumPy/SciPy LAPACK and plain C) agree with each other, with an extended-
precision (mpmath) evaluation on small problems, and with finite differences for the gradient."""
import numpy as np
import pytest

from gaplac_b200 import workloads as W
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
from oracle import c_oracle as CO
from oracle import gp_oracle as O

ALL_KINDS = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
             Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
             Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
THETA = np.array([1.3, 0.8, 0.4, 1.7, 0.2])


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    return X, rng.standard_normal(n)


def test_cov_c_vs_numpy():
    X, _ = _data(37)
    Kn = O.cov(ALL_KINDS, X, THETA, 0.1, 1e-9)
    Kc = CO.cov(ALL_KINDS, X, THETA, 0.1, 1e-9)
    assert np.max(np.abs(Kn - Kc)) < 1e-14
    assert np.allclose(Kn, Kn.T)


def test_gemm_expansion_mode_gap_is_small():
    """[upstream Distances] pairwise uses |a|^2+|b|^2-2ab; the direct (a-b)^2 differs by O(1e-13) here."""
    d = W.make_c2(n=128, B=4)
    v0, _ = CO.lml(d["ops"], d["X"], d["y"], d["Theta"][0], 0.0, mode=0)
    v1, _ = CO.lml(d["ops"], d["X"], d["y"], d["Theta"][0], 0.0, mode=1)
    assert abs(v0 - v1) < 1e-9 * abs(v0)


@pytest.mark.parametrize("n", [1, 5, 33])
def test_lml_posterior_predict_sample_c_vs_numpy(n):
    X, y = _data(n, seed=n)
    vn = O.lml(ALL_KINDS, X, y, THETA, 0.1)
    vc, info = CO.lml(ALL_KINDS, X, y, THETA, 0.1)
    assert info == 0 and abs(vn - vc) < 1e-11 * max(1, abs(vn))
    Un, an = O.posterior(ALL_KINDS, X, y, THETA, 0.1)
    Uc, ac = CO.posterior(ALL_KINDS, X, y, THETA, 0.1)
    assert np.max(np.abs(np.triu(Un) - Uc)) < 1e-12 and np.max(np.abs(an - ac)) < 1e-10
    Xs, _ = _data(11, seed=99)
    mn, vn_ = O.mean_and_var(ALL_KINDS, X, Un, an, Xs, THETA)
    mc, vc_ = CO.mean_and_var(ALL_KINDS, X, Uc, ac, Xs, THETA)
    assert np.max(np.abs(mn - mc)) < 1e-10 and np.max(np.abs(vn_ - vc_)) < 1e-10
    Z = np.random.default_rng(1).standard_normal((n, 3))
    assert np.max(np.abs(O.sample(ALL_KINDS, X, THETA, 0.1, Z) - CO.sample(ALL_KINDS, X, THETA, 0.1, Z))) < 1e-12


def test_not_positive_definite_reports_pivot():
    X = np.array([[0.0], [0.0], [1.0]])
    ops = [Op(SQEXP, col=0, value=1.0)]
    v, info = CO.lml(ops, X, np.zeros(3), [], 0.0)        # duplicated point, no noise -> singular at pivot 2
    assert v == -np.inf and info == 2
    assert O.lml(ops, X, np.zeros(3), [], 0.0) == -np.inf


def test_lml_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    n = 12
    X, y = _data(n, seed=5)
    ops = W.prog_c2()
    th = [1.5, 1.0, 0.1]
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            d = mp.mpf(X[i, 0]) - mp.mpf(X[j, 0])
            K[i, j] = mp.exp(-d * d / (2 * mp.mpf(th[0]) ** 2)) + mp.exp(-abs(d) / mp.mpf(th[1]))
            if i == j:
                K[i, j] += mp.mpf(th[2])
    L = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(v) for v in y])
    z = mp.lu_solve(L, yv)
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    truth = -(n * mp.log(2 * mp.pi) + logdet + sum(v * v for v in z)) / 2
    got = O.lml(ops, X[:, :1], y, th, 0.0)
    assert abs(got - float(truth)) < 1e-12 * abs(float(truth))


def test_gradient_matches_finite_differences():
    n = 20
    X, y = _data(n, seed=3)
    val, dth, dy = O.lml_grad(ALL_KINDS, X, y, THETA, 0.1)
    h = 1e-6
    for s in range(len(THETA)):
        tp, tm = THETA.copy(), THETA.copy()
        tp[s] += h
        tm[s] -= h
        fd = (O.lml(ALL_KINDS, X, y, tp, 0.1) - O.lml(ALL_KINDS, X, y, tm, 0.1)) / (2 * h)
        assert abs(fd - dth[s]) < 1e-6 * max(1.0, abs(fd))
    for i in (0, 7, n - 1):
        yp, ym = y.copy(), y.copy()
        yp[i] += h
        ym[i] -= h
        fd = (O.lml(ALL_KINDS, X, yp, THETA, 0.1) - O.lml(ALL_KINDS, X, ym, THETA, 0.1)) / (2 * h)
        assert abs(fd - dy[i]) < 1e-6 * max(1.0, abs(fd))


def test_lml_invariant_under_row_permutation():
    X, y = _data(25, seed=8)
    perm = np.random.default_rng(0).permutation(25)
    a = O.lml(ALL_KINDS, X, y, THETA, 0.1)
    b = O.lml(ALL_KINDS, X[perm], y[perm], THETA, 0.1)
    assert abs(a - b) < 1e-11 * abs(a)
