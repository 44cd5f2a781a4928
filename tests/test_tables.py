"""Host-side table ingestion against the reference's definitions and its own fixture."""
import os

import numpy as np
import pytest

from gaplac_b200 import tables as T


def test_getrank_and_invnormal_follow_the_reference_definition():
    from scipy.special import ndtri
    v = np.array([0.3, 0.0, 2.5, 0.0, -1.0, 0.3])
    # sortperm (stable): -1.0(5) 0.0(2) 0.0(4) 0.3(1) 0.3(6) 2.5(3)  ->  ranks by position: 4 2 6 3 1 5; zeros -> 1
    assert T.getrank(v, flattenzeros=False).tolist() == [4, 2, 6, 3, 1, 5]
    assert T.getrank(v).tolist() == [4, 1, 6, 1, 1, 5]
    r = np.array([4, 1, 6, 1, 1, 5], dtype=float)
    assert np.allclose(T.invnormaltransform(v), ndtri((r - 0.375) / (6 - 0.75 + 1)), rtol=0, atol=1e-15)
    z = T.invnormaltransform(np.random.default_rng(0).gamma(2.0, size=500))
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1.0) < 0.02      # symmetric plotting positions


def test_design_matrix_on_the_reference_fixture_layout():
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    tab = T.read_table(os.path.join(gdir, "input_pair_3206.csv"))
    tab = T.complete_cases(tab)
    X, levels = T.design_matrix(tab, ["PersonID", "StoolPairs", "PersonID", "nutrient"])   # Cat*Cat + Cat + Linear
    assert X.shape == (923, 4) and not levels                      # ids are numeric in the fixture
    assert np.array_equal(X[:, 0], X[:, 2])
    from oracle import gp_oracle as O
    Xo, yo, *_ = O.load_golden("3206", gdir)
    assert np.array_equal(X[:, [0, 1, 3]], Xo)
    assert np.array_equal(np.array([float(b) for b in tab["bug"]]), yo)


def test_categorical_encoding_missing_rows_and_packing():
    text = "subject\ttime\tf1\tf2\nA\t0.5\t1.0\t0\nB\t1.5\tNA\t3\nA\t2.5\t2.0\t0\nC\t3.5\t0.5\t7\n"
    tab = T.read_table(text)
    assert list(tab) == ["subject", "time", "f1", "f2"] and len(tab["time"]) == 4
    cc = T.complete_cases(tab)
    assert cc["subject"] == ["A", "A", "C"]
    X, levels = T.design_matrix(cc, ["subject", "time"])
    assert levels == {"subject": ["A", "C"]} and X[:, 0].tolist() == [1.0, 1.0, 2.0] and X[:, 1].tolist() == [0.5, 2.5, 3.5]
    Y = T.pack_responses(cc, ["f1", "f2"])
    assert Y.shape == (2, 3) and np.all(np.isfinite(Y))
    assert Y[1, 0] == Y[1, 1]                                       # the two zeros of f2 share rank 1
    with pytest.raises(KeyError):
        T.design_matrix(cc, ["nope"])
    with pytest.raises(ValueError):
        T.read_table("a,b\n1\n")
