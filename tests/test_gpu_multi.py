"""Several devices behind one call (gpl_multi_*): the batch is split into contiguous blocks, every device writes its slice
into the caller's buffers.  On a one-GPU box the same device is listed twice (two contexts, two streams) so the sharding
and gathering logic runs everywhere; with two or more GPUs the real thing runs too."""
import threading

import numpy as np
import pytest

from gaplac_b200 import _lib, mcmc, workloads as W

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1], [0, 1, 2, 3, 4, 5, 6, 7]])
def test_multi_lml_and_gradient_equal_the_single_device_call(ctx, devices):
    if max(devices) >= _device_count():
        pytest.skip(f"needs {max(devices) + 1} devices")
    d = W.make_c3(features=11)                          # 11 items over 2 / 3 / 8 parts: uneven blocks
    m = _lib.MultiContext(devices)
    try:
        prog_m = m.program(d["ops"])
        got = m.lml_batched(prog_m, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    finally:
        m.close()
    ref = ctx.lml_batched(ctx.program(d["ops"]), d["X"], d["Y"], d["Theta"], 0.0, grad=True)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


def test_multi_sampler_keeps_every_chain(ctx):
    """Chains sharded over parts keep their random streams (chain_offset): same chains as on one device."""
    d = W.make_c1()
    kw = dict(sigma2=0.1, n_samples=5, n_adapt=10, seed=3, chains=5, record_warmup=True)
    one = mcmc.nuts(ctx, ctx.program(d["ops"]), d["X"], d["y"], [0.0], [20.0], **kw)
    m = _lib.MultiContext([0, 0])
    try:
        two = mcmc.nuts(m, m.program(d["ops"]), d["X"], d["y"], [0.0], [20.0], **kw)
    finally:
        m.close()
    for k in ("theta", "lp", "eps", "depth", "n_leapfrog"):
        assert np.array_equal(one[k], two[k]), k


def test_multi_reports_which_part_failed():
    m = _lib.MultiContext([0, 0])
    try:
        prog = m.program(W.prog_c2())
        with pytest.raises(_lib.GaplacError) as e:
            m.lml_batched(prog, np.zeros((5, 1)), np.zeros(5), np.ones((4, 2)), 0.0)      # p = 2 < 3 slots
        assert "part" in str(e.value) and "hyperparameter slots" in str(e.value)
    finally:
        m.close()


@pytest.mark.parametrize("trail_int8", [0, 8])
def test_large_n_path_survives_a_busy_device(trail_int8):
    """The look-ahead factorisation hands work between concurrently resident kernels through flags (bounded waits).  Run
    it while a second context keeps the GPU saturated with batched evaluations: it must finish, and correctly - with the
    DMMA trailing updates and with the persistent INT8 kernels (which need whole SMs) alike."""
    from oracle import c_oracle as CO
    busy = _lib.Context(0)
    big = _lib.Context(0)
    big.set_option("trail_int8", trail_int8)
    d2 = W.make_c2(n=512, B=1024)
    stop = threading.Event()

    def hammer():
        prog = busy.program(d2["ops"])
        while not stop.is_set():
            busy.lml_batched(prog, d2["X"], d2["y"], d2["Theta"], 0.0)

    t = threading.Thread(target=hammer)
    t.start()
    try:
        d = W.make_c5(n=4096)
        prog = big.program(d["ops"])
        res = [big.lml_large(prog, d["X"], d["y"], d["theta"], 0.0) for _ in range(3)]
    finally:
        stop.set()
        t.join()
        busy.close()
    CO.use_openblas(4)
    ref, rinfo = CO.lml(d["ops"], d["X"], d["y"], d["theta"], 0.0)
    CO.use_openblas(1)
    for lml, ld, info in res:
        assert info == 0 and rinfo == 0
        assert abs(lml - ref) < 1e-9 * abs(ref)
    big.close()
