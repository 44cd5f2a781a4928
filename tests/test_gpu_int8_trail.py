"""Large-n factorisation with trailing updates on the INT8 tensor path (option "trail_int8", csrc/trail_int8.cu +
int8_syrk.cuh: tcgen05.mma.kind::i8, TMEM, TMA).  Replaces the DMMA updates of the same [upstream] cholesky the reference
reaches from logpdf / posterior at large n (kernel 3); the bar is the same: lml 1e-9, posterior 1e-8.
"""
import numpy as np
import pytest

from gaplac_b200 import _lib, workloads as W
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


@pytest.fixture()
def own_ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _lapack_lml(d):
    import scipy.linalg as sla
    from oracle import gp_oracle as O
    K = O.cov(d["ops"], d["X"], d["theta"], 0.0)
    c = sla.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    z = sla.solve_triangular(c, d["y"], lower=True, check_finite=False)
    ld = 2.0 * np.sum(np.log(np.diag(c)))
    return -0.5 * (len(d["y"]) * O.LOG2PI + ld + z @ z), ld


@pytest.mark.parametrize("n", [4096, 4305, 5000])  # 4305 -> 68 tile rows (odd 128-blocks at the block ends), 5000 -> ragged last tile
def test_eight_slices_equal_the_fp64_path_and_lapack(own_ctx, n):
    ctx = own_ctx
    d = W.make_c5(n=n)
    prog = ctx.program(d["ops"])
    ctx.set_option("trail_int8", 0)
    base = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("trail_int8", 8)
    got = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    assert got[2] == 0 and base[2] == 0
    assert abs(got[0] - base[0]) < 1e-12 * abs(base[0]) and abs(got[1] - base[1]) < 1e-12 * abs(base[1])
    ref, ldref = _lapack_lml(d)
    assert abs(got[0] - ref) < 1e-11 * abs(ref) and abs(got[1] - ldref) < 1e-11 * abs(ldref)
    # run to run: integer products and a fixed FP64 summation order
    assert ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0) == got


@pytest.mark.parametrize("slices,tol", [(7, 1e-10), (6, 1e-9)])
def test_fewer_slices_stay_inside_the_lml_tolerance(own_ctx, slices, tol):
    ctx = own_ctx
    d = W.make_c5(n=4096)
    prog = ctx.program(d["ops"])
    ctx.set_option("trail_int8", 0)
    base = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("trail_int8", slices)
    got = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    assert got[2] == 0 and abs(got[0] - base[0]) < tol * abs(base[0])


def test_automatic_mode_switches_to_int8_for_large_n_and_reports_the_same_answer(own_ctx):
    """Default option value: DMMA below n = 6144 (bitwise the DMMA path), the INT8 path from there on."""
    ctx = own_ctx
    d = W.make_c5(n=4096)
    prog = ctx.program(d["ops"])
    auto = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("trail_int8", 0)
    assert ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0) == auto
    d = W.make_c5(n=8192)
    n0 = ctx.launch_count()
    off = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    n1 = ctx.launch_count()
    ctx.set_option("trail_int8", -1)
    auto = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    n2 = ctx.launch_count()
    assert n2 - n1 != n1 - n0                                   # a different kernel sequence ran ...
    assert abs(auto[0] - off[0]) < 1e-12 * abs(off[0])          # ... to the same answer
    with pytest.raises(_lib.GaplacError):
        ctx.set_option("trail_int8", 3)


def test_posterior_through_the_int8_path(own_ctx):
    """gpl_posterior_fit shares the factorisation: alpha and predictions against the FP64 path and the oracle."""
    ctx = own_ctx
    d = W.make_c4(n=4096, m=64)
    prog = ctx.program(d["ops"])
    ctx.set_option("trail_int8", 0)
    p0 = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)
    a0, (m0, v0) = p0.alpha(), p0.mean_and_var(d["Xs"])
    p0.free()
    ctx.set_option("trail_int8", 8)
    p1 = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)
    a1, (m1, v1) = p1.alpha(), p1.mean_and_var(d["Xs"])
    assert abs(p1.logpdf() - CO.lml(d["ops"], d["X"], d["y"], d["theta"], 0.0)[0]) < 1e-9 * abs(p1.logpdf())
    p1.free()
    assert np.max(np.abs(a1 - a0)) < 1e-9 * np.max(np.abs(a0))
    assert np.max(np.abs(m1 - m0)) < 1e-8 and np.max(np.abs(v1 - v0)) < 1e-8


def test_indefinite_matrix_still_reports_the_first_failing_pivot(own_ctx):
    ctx = own_ctx
    ctx.set_option("trail_int8", 8)
    n = 4200
    rng = np.random.default_rng(1)
    A = np.eye(n) * 4.0 + 0.5 * np.add.outer(np.sin(np.arange(n)), np.sin(np.arange(n))) / n
    A[3000, 3000] = -1.0
    _, ld, info = ctx.chol_logdet(A, want_factor=False)
    assert info == 3001 and np.isnan(ld)
    A[3000, 3000] = 4.0
    _, ld, info = ctx.chol_logdet(A, want_factor=False)
    assert info == 0 and abs(ld - np.linalg.slogdet(A)[1]) < 1e-10 * abs(ld)


def test_workspace_poisoning_does_not_reach_the_result(own_ctx):
    """Slices, row scales and breadcrumb live in the context workspace: NaN-filled before the call, same bits after."""
    ctx = own_ctx
    d = W.make_c5(n=4096)
    prog = ctx.program(d["ops"])
    ctx.set_option("trail_int8", 8)
    clean = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("poison_ws", 1)
    assert ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0) == clean


def test_look_ahead_without_the_worker_cta(own_ctx):
    """chol_variant 3 (what every n > 8960 takes: the worker protocol stops at 140 tile rows) with the INT8 passes."""
    ctx = own_ctx
    d = W.make_c5(n=4608)
    prog = ctx.program(d["ops"])
    ctx.set_option("trail_int8", 0)
    base = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("chol_variant", 3)
    ctx.set_option("trail_int8", 8)
    got = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    assert got[2] == 0 and abs(got[0] - base[0]) < 1e-12 * abs(base[0])


def test_automatic_slice_count_holds_the_tolerance_on_an_ill_conditioned_covariance(own_ctx):
    """SqExp with l = 3 on 6144 points in (-50, 50) plus 1e-6 noise: condition number ~1e10, FP64 itself is good to ~2e-10 here.
    The automatic setting (9 slices) stays at that level; 8 slices - 56 bits below the row maxima - lose another digit (4e-9),
    which is why they are not the default."""
    ctx = own_ctx
    d = W.make_c5(n=6144)
    d["theta"] = np.array([3.0, 1e-6])
    prog = ctx.program(d["ops"])
    ref, _ = _lapack_lml(d)
    auto = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("trail_int8", 0)
    dmma = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("trail_int8", 8)
    eight = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    assert auto[2] == 0 and dmma[2] == 0 and eight[2] == 0
    e_auto, e_dmma, e_eight = (abs(v[0] - ref) / abs(ref) for v in (auto, dmma, eight))
    assert e_auto < 1e-9 and e_dmma < 1e-9
    assert e_auto < 3 * e_dmma + 1e-12
    assert e_eight < 1e-7  # usable, but a digit behind
