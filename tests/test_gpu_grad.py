"""Analytic gradient on the lockstep schedule (csrc/lml_grad_lockstep.cu) against the CPU oracle, at the shapes the
reference's sampler would call it on: the mcmc model body (CLI/src/mcmc.jl:31-37) needs d lml / d(hyperparameters) and
d lml / d(latent fx) at every leapfrog step.  Tolerance (north_star): 1e-8.  Through the C ABI; needs a B200."""
import numpy as np
import pytest

from gaplac_b200 import workloads as W
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-8
LML_RTOL = 1e-9

ALL_KINDS = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
             Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
             Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
THETA = np.array([1.3, 0.8, 0.4, 1.7, 0.2])


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    return X, rng.standard_normal(n)


def _check(lml, dth, dy, ops, X, y, theta, sigma2, jitter=0.0):
    val, rdth, rdy = O.lml_grad(ops, X, y, theta, sigma2, jitter)
    assert abs(lml - val) <= LML_RTOL * abs(val)
    assert np.max(np.abs(dth - rdth) / np.maximum(1.0, np.abs(rdth))) < GRAD_TOL
    assert np.max(np.abs(dy - rdy)) < GRAD_TOL * max(1.0, np.max(np.abs(rdy)))


def test_gradient_c2_headline_shape(ctx):
    """SqExp(:x)+OU(:x)+Noise at n = 512 (BASELINE config[1])."""
    d = W.make_c2(n=512, B=12)
    prog = ctx.program(d["ops"])
    lml, info, dth, dy = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    assert not info.any()
    for b in (0, 5, 11):
        _check(lml[b], dth[b], dy[b], d["ops"], d["X"], d["y"], d["Theta"][b], 0.0)
    plain, _ = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0)
    assert np.array_equal(plain, lml)                       # same factorisation kernels with and without the gradient


def test_gradient_c3_microbiome_shape(ctx):
    """Cat(:subject)*SqExp(:time)+Noise at n = 300 (not a multiple of 64), one response per feature (config[2])."""
    d = W.make_c3(features=10)
    prog = ctx.program(d["ops"])
    lml, info, dth, dy = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    assert not info.any()
    for b in (0, 4, 9):
        _check(lml[b], dth[b], dy[b], d["ops"], d["X"], d["Y"][b], d["Theta"][b], 0.0)


def test_gradient_golden_program_n923(ctx, golden_dir):
    """The reference's legacy model (4 variances, Cat x Cat + Cat + Linear + Noise, jitter 1e-9) at n = 923."""
    X, y, Th, s2, _, _ = O.load_golden("3206", golden_dir)
    ops = O.golden_program("3206")
    prog = ctx.program(ops)
    rows = [0, 57, 99]
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, Th[rows], s2[rows], jitter=O.GOLDEN_JITTER, grad=True)
    assert not info.any()
    for k, r in enumerate(rows):
        _check(lml[k], dth[k], dy[k], ops, X, y, Th[r], s2[r], O.GOLDEN_JITTER)


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 129, 200, 321])
def test_gradient_ragged_sizes_every_leaf_kind(ctx, n):
    X, y = _data(n, seed=20 + n)
    prog = ctx.program(ALL_KINDS)
    Th = np.vstack([THETA, THETA * 1.1, THETA * 0.9])
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, Th, 0.1, grad=True)
    assert not info.any()
    for b in range(3):
        _check(lml[b], dth[b], dy[b], ALL_KINDS, X, y, Th[b], 0.1)


def test_gradient_beyond_the_shared_memory_z_window(ctx):
    n = 1100
    X, y = _data(n, seed=70 + n)
    prog = ctx.program(ALL_KINDS)
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, THETA[None, :], 0.1, grad=True)
    assert info[0] == 0
    _check(lml[0], dth[0], dy[0], ALL_KINDS, X, y, THETA, 0.1)


def test_gradient_per_item_inputs(ctx):
    """X, y and sigma2 per item (fully independent models in one call)."""
    B, n = 5, 90
    rng = np.random.default_rng(3)
    Xb = np.stack([_data(n, seed=s)[0] for s in range(B)])
    Yb = rng.standard_normal((B, n))
    Th = THETA[None, :] * rng.uniform(0.8, 1.2, (B, 5))
    s2 = rng.uniform(0.05, 0.3, B)
    prog = ctx.program(ALL_KINDS)
    lml, info, dth, dy = ctx.lml_batched(prog, Xb, Yb, Th, s2, grad=True)
    assert not info.any()
    for b in range(B):
        _check(lml[b], dth[b], dy[b], ALL_KINDS, Xb[b], Yb[b], Th[b], s2[b])


def test_gradient_agrees_with_the_fused_per_item_kernel(ctx):
    """lml_variant = 1 selects the one-CTA-per-item kernel of round 1: an independent implementation of the same maths."""
    d = W.make_c3(features=6)
    prog = ctx.program(d["ops"])
    a = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    ctx.set_option("lml_variant", 1)
    try:
        b = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    finally:
        ctx.set_option("lml_variant", 0)
    assert np.max(np.abs(a[0] - b[0]) / np.abs(a[0])) < LML_RTOL
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(1.0, np.abs(a[2]))) < GRAD_TOL
    assert np.max(np.abs(a[3] - b[3])) < GRAD_TOL * max(1.0, np.max(np.abs(a[3])))


def test_gradient_is_bitwise_reproducible_and_chunking_is_invisible(ctx):
    d = W.make_c3(features=120)
    prog = ctx.program(d["ops"])
    a = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    b = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    ctx.set_option("lk_ws_limit_mb", 24)        # ~ 20 items per pass with the gradient workspace
    try:
        c = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    finally:
        ctx.set_option("lk_ws_limit_mb", 24 * 1024)
    for k in (0, 2, 3):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


def test_gradient_of_bad_items_is_nan_and_neighbours_are_untouched(ctx):
    d = W.make_c2(n=130, B=5)
    Th = d["Theta"].copy()
    Th[1, 0] = -1.0          # negative length scale: rejected per item (ScaleTransform throws in the reference)
    Th[3, 1] = np.nan
    prog = ctx.program(d["ops"])
    lml, info, dth, dy = ctx.lml_batched(prog, d["X"], d["y"], Th, 0.0, grad=True)
    for b in (1, 3):
        assert info[b] != 0 and lml[b] == -np.inf and np.all(np.isnan(dth[b])) and np.all(np.isnan(dy[b]))
    for b in (0, 2, 4):
        assert info[b] == 0
        _check(lml[b], dth[b], dy[b], d["ops"], d["X"], d["y"], Th[b], 0.0)


def test_gradient_through_the_device_pointer_entry(ctx):
    torch = pytest.importorskip("torch")
    d = W.make_c2(n=200, B=7)
    prog = ctx.program(d["ops"])
    dev = torch.device("cuda", 0)
    X = torch.from_numpy(np.ascontiguousarray(d["X"].T)).to(dev)
    y = torch.from_numpy(d["y"]).to(dev)
    Th = torch.from_numpy(np.ascontiguousarray(d["Theta"])).to(dev)
    s2 = torch.zeros(1, dtype=torch.float64, device=dev)
    B, p, n = 7, 3, 200
    lml = torch.empty(B, dtype=torch.float64, device=dev)
    dth = torch.empty(B, p, dtype=torch.float64, device=dev)
    dy = torch.empty(B, n, dtype=torch.float64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ctx.lml_batched_dev(prog, n, 1, X.data_ptr(), False, y.data_ptr(), False, Th.data_ptr(), p, s2.data_ptr(), False,
                            0.0, B, lml.data_ptr(), dth.data_ptr(), dy.data_ptr(), info.data_ptr(), st.cuda_stream)
    # a second call on ANOTHER stream right behind it: the context orders workspace reuse across streams
    lml2 = torch.empty_like(lml)
    ctx.lml_batched_dev(prog, n, 1, X.data_ptr(), False, y.data_ptr(), False, Th.data_ptr(), p, s2.data_ptr(), False,
                        0.0, B, lml2.data_ptr(), 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    assert np.array_equal(lml.cpu().numpy(), h[0]) and np.array_equal(lml2.cpu().numpy(), h[0])
    assert np.array_equal(dth.cpu().numpy(), h[2]) and np.array_equal(dy.cpu().numpy(), h[3])
