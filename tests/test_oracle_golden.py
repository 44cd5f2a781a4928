"""Pin the oracle: the reference's own legacy fixtures give 200 known lml values at n=923.

lπ (chain column) = lml(y = bug; K(θ) + 1e-9 I) + closed-form legacy prior.  Commands that produced the
fixtures: /root/reference/test/pred.jl:3,22.  tests/golden/*.csv were extracted by tools/make_golden.py.
"""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import gp_oracle as O

TOL = 5e-11   # absolute, on |lπ| ~ 640..940  (observed: <= 3.1e-12)


@pytest.mark.parametrize("tag", ["3206", "1003"])
def test_numpy_oracle_reproduces_fixture_rows(tag, golden_dir):
    X, y, Th, s2, lpi, prior = O.load_golden(tag, golden_dir)
    assert X.shape == (923, 3) and len(Th) == 100
    ops = O.golden_program(tag)
    rows = list(range(0, 100, 9))          # SciPy path: a spread of rows keeps the CPU suite short
    got = np.array([O.lml(ops, X, y, Th[r], s2[r], O.GOLDEN_JITTER) for r in rows])
    assert np.max(np.abs(got + prior[rows] - lpi[rows])) < TOL


@pytest.mark.parametrize("tag", ["3206", "1003"])
def test_c_oracle_reproduces_all_fixture_rows(tag, golden_dir):
    X, y, Th, s2, lpi, prior = O.load_golden(tag, golden_dir)
    ops = O.golden_program(tag)
    CO.use_openblas(1)                     # the reference's LAPACK path (dpotrf / dtrtrs)
    try:
        got, info = CO.lml_batched(ops, X, y, Th, s2, O.GOLDEN_JITTER, threads=0)
    finally:
        CO.use_plain_c()
    assert not info.any()
    assert np.max(np.abs(got + prior - lpi)) < TOL


def test_plain_c_cholesky_matches_lapack_on_fixture(golden_dir):
    X, y, Th, s2, lpi, prior = O.load_golden("3206", golden_dir)
    ops = O.golden_program("3206")
    got, info = CO.lml_batched(ops, X, y, Th[:8], s2[:8], O.GOLDEN_JITTER, threads=0)
    assert not info.any()
    assert np.max(np.abs(got + prior[:8] - lpi[:8])) < TOL


def test_fixture_without_jitter_is_measurably_off(golden_dir):
    """The 1e-9 jitter is part of the identity: without it the residual is ~2e-7."""
    X, y, Th, s2, lpi, prior = O.load_golden("3206", golden_dir)
    v, _ = CO.lml(O.golden_program("3206"), X, y, Th[0], s2[0], 0.0)
    assert 1e-8 < abs(v + prior[0] - lpi[0]) < 1e-5
