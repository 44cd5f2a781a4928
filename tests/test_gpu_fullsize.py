"""BASELINE.json's full sizes: spot checks against the oracle plus size-independent properties."""
import numpy as np
import pytest

from gaplac_b200 import workloads as W
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def test_c2_full_batch_4096(ctx):
    d = W.make_c2()                                        # n = 512, B = 4096
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0)
    assert not info.any() and np.all(np.isfinite(lml))
    idx = np.random.default_rng(0).choice(4096, 24, replace=False)
    CO.use_openblas(1)
    try:
        ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"][idx], 0.0)
    finally:
        CO.use_plain_c()
    assert np.max(np.abs(lml[idx] - ref) / np.abs(ref)) < 1e-9
    # property: the result of an item does not depend on its position in the batch or on its neighbours
    perm = np.random.default_rng(1).permutation(4096)
    lml2, _ = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"][perm], 0.0)
    assert np.array_equal(lml2, lml[perm])
    # property: permuting the observations leaves lml unchanged (to rounding)
    rp = np.random.default_rng(2).permutation(512)
    lml3, _ = ctx.lml_batched(prog, d["X"][rp], d["y"][rp], d["Theta"][:64], 0.0)
    assert np.max(np.abs(lml3 - lml[:64]) / np.abs(lml[:64])) < 1e-10


def test_c3_all_features(ctx):
    d = W.make_c3()                                        # 2000 features x n = 300
    prog = ctx.program(d["ops"])
    lml, info = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0)
    assert not info.any()
    idx = np.arange(0, 2000, 97)
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["Y"][idx], d["Theta"][idx], 0.0)
    assert np.max(np.abs(lml[idx] - ref) / np.abs(ref)) < 1e-9


def test_c4_predict_2048_to_20000(ctx):
    d = W.make_c4()
    prog = ctx.program(d["ops"])
    post = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)
    mean, var = post.mean_and_var(d["Xs"])
    assert mean.shape == (20000,) and np.all(np.isfinite(mean)) and np.all(var > -1e-8)
    CO.use_openblas(8)
    try:
        U, alpha = CO.posterior(d["ops"], d["X"], d["y"], d["theta"], 0.0)
        idx = np.arange(0, 20000, 331)
        rm, rv = CO.mean_and_var(d["ops"], d["X"], U, alpha, d["Xs"][idx], d["theta"])
    finally:
        CO.use_plain_c()
    assert np.max(np.abs(mean[idx] - rm)) < 1e-8 * max(1.0, np.max(np.abs(rm)))
    assert np.max(np.abs(var[idx] - rv)) < 1e-8 * max(1.0, np.max(np.abs(rv)))
    # property: predicting at the training inputs reproduces K alpha = y - sigma_n^2 alpha
    m2 = post.mean_and_var(d["X"][:256], want_var=False)
    assert np.max(np.abs(m2 - (d["y"][:256] - d["theta"][2] * alpha[:256]))) < 1e-8
    post.free()


def test_c5_large_n_8192(ctx):
    d = W.make_c5()
    prog = ctx.program(d["ops"])
    lml, ld, info = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    assert info == 0
    import scipy.linalg as sla
    from oracle import gp_oracle as O
    K = O.cov(d["ops"], d["X"], d["theta"], 0.0)
    c = sla.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    ldref = 2.0 * np.sum(np.log(np.diag(c)))
    z = sla.solve_triangular(c, d["y"], lower=True, check_finite=False)
    ref = -0.5 * (8192 * O.LOG2PI + ldref + z @ z)
    assert abs(ld - ldref) < 1e-9 * abs(ldref)
    assert abs(lml - ref) < 1e-9 * abs(ref)
