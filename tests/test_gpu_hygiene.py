"""Memory hygiene without compute-sanitizer (closed on this pool: profiles/sanitizer_r02.txt):
  * canaries - the device-pointer entry points write only inside the caller's output buffers;
  * poisoning - with the whole workspace filled with NaN payloads before every call, all results keep their bits
    (no kernel reads workspace that was not written in the same call);
  * determinism - every path twice, identical bits (a race between CTAs / warps shows as run-to-run differences)."""
import numpy as np
import pytest

from gaplac_b200 import _lib, mcmc, workloads as W
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP

pytestmark = pytest.mark.gpu

ALL_KINDS = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
             Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
             Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
THETA = np.array([1.3, 0.8, 0.4, 1.7, 0.2])
CANARY = -7.0e300


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    return (np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                             rng.integers(0, 4, n).astype(float)]), rng.standard_normal(n))


def _everything(c, big_n=700):
    """One call of every kernel family; returns all results as a flat list of arrays."""
    prog = c.program(ALL_KINDS)
    X, y = _data(150, 1)
    Th = np.vstack([THETA, THETA * 1.1, THETA * 0.9])
    out = []
    out += list(c.lml_batched(prog, X, y, Th, 0.1))
    out += list(c.lml_batched(prog, X, y, Th, 0.1, grad=True))
    out.append(c.cov(prog, X, THETA, 0.1))
    out.append(c.cross_cov(prog, X, _data(70, 2)[0], THETA))
    post = c.posterior_fit(prog, X, y, THETA, 0.1)
    out += list(post.mean_and_var(_data(333, 3)[0]))
    out.append(post.factor())
    post.free()
    out += [a for a in c.predict_batched(prog, X, y, Th, 0.1, _data(40, 4)[0])]
    out.append(c.sample(prog, X, THETA, 0.1, np.random.default_rng(5).standard_normal((150, 2))))
    Xb, yb = _data(big_n, 6)                                           # large-n path: look-ahead streams + worker CTA
    p2 = c.posterior_fit(prog, Xb, yb, THETA, 0.1)
    out += list(p2.mean_and_var(_data(50, 7)[0]))
    out.append(p2.alpha())
    p2.free()
    d5 = W.make_c5(n=1500)
    out.append(np.array(c.lml_large(c.program(d5["ops"]), d5["X"], d5["y"], d5["theta"], 0.0)))
    A = np.cov(np.random.default_rng(8).standard_normal((300, 900))) + np.eye(300)
    U, ld, info = c.chol_logdet(A)
    out += [U, np.array([ld, info])]
    d1 = W.make_c1(n=20)
    r = mcmc.nuts(c, c.program(d1["ops"]), d1["X"], d1["y"], [0.0], [20.0], sigma2=0.1, n_samples=4, n_adapt=4, seed=1, chains=3)
    out += [r["theta"], r["lp"], r["eps"]]
    return out


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True)


def test_poisoned_workspace_and_reruns_keep_every_bit():
    c = _lib.Context(0)
    try:
        first = _everything(c)
        second = _everything(c)                     # determinism: same context, warm workspaces
        _same(first, second)
        c.set_option("poison_ws", 1)
        _same(first, _everything(c))                # every call starts from a NaN-filled workspace
        c.set_option("poison_ws", 0)
    finally:
        c.close()
    fresh = _lib.Context(0)                         # ... and from a freshly allocated one
    try:
        _same(first, _everything(fresh))
    finally:
        fresh.close()


def test_large_n_worker_protocol_is_deterministic(ctx):
    d = W.make_c5(n=4096)
    prog = ctx.program(d["ops"])
    a = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    for _ in range(3):
        assert ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0) == a


@pytest.mark.parametrize("n,B", [(1, 1), (63, 3), (200, 5), (321, 2)])
def test_device_entry_points_stay_inside_the_callers_buffers(ctx, n, B):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    X, y = _data(n, n)
    prog = ctx.program(ALL_KINDS)
    p, G = 5, 64                                                    # guard band of 64 elements on both sides
    Th = np.vstack([THETA * (1.0 + 0.03 * b) for b in range(B)])
    dX = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
    dY = torch.from_numpy(y).to(dev)
    dTh = torch.from_numpy(np.ascontiguousarray(Th)).to(dev)
    dS = torch.tensor([0.1], dtype=torch.float64, device=dev)

    def guarded(count, dtype=torch.float64, fill=CANARY):
        t = torch.full((count + 2 * G,), fill, dtype=dtype, device=dev)
        return t, t[G:G + count]

    lml_g, lml = guarded(B)
    dth_g, dth = guarded(B * p)
    dy_g, dy = guarded(B * n)
    info_g, info = guarded(B, torch.int32, -77)
    st = torch.cuda.current_stream().cuda_stream
    ctx.lml_batched_dev(prog, n, 4, dX.data_ptr(), False, dY.data_ptr(), False, dTh.data_ptr(), p, dS.data_ptr(), False, 0.0,
                        B, lml.data_ptr(), dth.data_ptr(), dy.data_ptr(), info.data_ptr(), st)
    K_g, K = guarded(n * n)
    ctx.cov_dev(prog, n, 4, dX.data_ptr(), dTh.data_ptr(), p, 0.1, 0.0, K.data_ptr(), st)
    A = np.cov(np.random.default_rng(n).standard_normal((n, 3 * n + 3))).reshape(n, n) + np.eye(n)
    A_g, Ad = guarded(n * n)
    Ad.copy_(torch.from_numpy(np.asfortranarray(A).ravel(order="K")).to(dev))
    ld_g, ld = guarded(1)
    ci_g, ci = guarded(1, torch.int32, -77)
    ctx.chol_logdet_dev(n, Ad.data_ptr(), True, ld.data_ptr(), ci.data_ptr(), st)
    torch.cuda.synchronize()
    for g, fill in ((lml_g, CANARY), (dth_g, CANARY), (dy_g, CANARY), (K_g, CANARY), (A_g, CANARY), (ld_g, CANARY)):
        assert bool((g[:G] == fill).all()) and bool((g[-G:] == fill).all())
    for g in (info_g, ci_g):
        assert bool((g[:G] == -77).all()) and bool((g[-G:] == -77).all())
    # ... and everything inside was written
    for t in (lml, dth, dy, K, ld):
        assert not bool((t == CANARY).any())
    assert int(info.abs().sum()) == 0 and int(ci[0]) == 0
    host = ctx.lml_batched(prog, X, y, Th, 0.1, grad=True)
    assert np.array_equal(lml.cpu().numpy(), host[0]) and np.array_equal(dy.cpu().numpy().reshape(B, n), host[3])


def test_release_workspace_and_regrow(ctx):
    d = W.make_c2(n=200, B=9)
    prog = ctx.program(d["ops"])
    a = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    ctx.release_workspace()
    b = ctx.lml_batched(prog, d["X"], d["y"], d["Theta"], 0.0, grad=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
