#!/usr/bin/env python3
"""Step-by-step diagnostics on a GPU box: exercises every C-ABI entry against the oracle and prints the errors
instead of asserting, so one gpurun round trip tells as much as possible.  Not part of the test suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaplac_b200 import _lib, workloads as W   # noqa: E402
from gaplac_b200.formula import Op             # noqa: E402
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP  # noqa: E402
from oracle import c_oracle as CO, gp_oracle as O  # noqa: E402

ALL = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL),
       Op(LINEAR, col=2, theta_slot=2), Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD),
       Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
TH = np.array([1.3, 0.8, 0.4, 1.7, 0.2])


def data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                         rng.integers(0, 4, n).astype(float)])
    return X, rng.standard_normal(n)


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[ok ] {name}: {r}  ({time.time() - t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


ctx = _lib.Context(0)
print(ctx.device_info())
prog = ctx.program(ALL)


def t_cov(n):
    X, _ = data(n, n)
    K = ctx.cov(prog, X, TH, 0.1, 1e-9)
    return float(np.max(np.abs(K - CO.cov(ALL, X, TH, 0.1, 1e-9))))


def t_lml(n, B=3):
    X, y = data(n, n)
    Th = np.vstack([TH * (1 + 0.05 * b) for b in range(B)])
    lml, info = ctx.lml_batched(prog, X, y, Th, 0.1)
    ref, _ = CO.lml_batched(ALL, X, y, Th, 0.1)
    return f"rel {np.max(np.abs(lml - ref) / np.abs(ref)):.2e} info {info.tolist()} lml {lml[0]:.12g} ref {ref[0]:.12g}"


def t_grad(n):
    X, y = data(n, n)
    lml, info, dth, dy = ctx.lml_batched(prog, X, y, TH[None, :], 0.1, grad=True)
    val, rdth, rdy = O.lml_grad(ALL, X, y, TH, 0.1)
    return (f"lml rel {abs(lml[0] - val) / abs(val):.2e} dth rel {np.max(np.abs(dth[0] - rdth) / np.maximum(1, np.abs(rdth))):.2e} "
            f"dy abs {np.max(np.abs(dy[0] - rdy)):.2e}")


def t_post(n, variant=0):
    X, y = data(n, n)
    Xs, _ = data(150, 77)
    ctx.set_option("chol_variant", variant)
    try:
        post = ctx.posterior_fit(prog, X, y, TH, 0.1)
    finally:
        ctx.set_option("chol_variant", 0)
    U, alpha = CO.posterior(ALL, X, y, TH, 0.1)
    ref, _ = CO.lml(ALL, X, y, TH, 0.1)
    a = post.alpha()
    Ug = post.factor()
    m, v = post.mean_and_var(Xs)
    rm, rv = CO.mean_and_var(ALL, X, U, alpha, Xs, TH)
    return (f"lml rel {abs(post.logpdf() - ref) / abs(ref):.2e} alpha {np.max(np.abs(a - alpha)):.2e} "
            f"U {np.max(np.abs(Ug - U)):.2e} mean {np.max(np.abs(m - rm)):.2e} var {np.max(np.abs(v - rv)):.2e}")


def t_sample(n):
    X, _ = data(n, n)
    Z = np.random.default_rng(5).standard_normal((n, 5))
    return float(np.max(np.abs(ctx.sample(prog, X, TH, 0.1, Z) - CO.sample(ALL, X, TH, 0.1, Z))))


def t_chol(n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    U, ld, info = ctx.chol_logdet(A)
    Ur, ldr, _ = CO.chol_logdet(A)
    return f"info {info} logdet rel {abs(ld - ldr) / abs(ldr):.2e} U {np.max(np.abs(U - Ur)):.2e}"


def t_large(n):
    d = W.make_c5(n=n)
    p5 = ctx.program(d["ops"])
    t = time.time()
    lml, ld, info = ctx.lml_large(p5, d["X"], d["y"], d["theta"], 0.0)
    dt = time.time() - t
    CO.use_openblas(8)
    ref, _ = CO.lml(d["ops"], d["X"], d["y"], d["theta"], 0.0)
    CO.use_plain_c()
    return f"info {info} rel {abs(lml - ref) / abs(ref):.2e} gpu wall {dt * 1e3:.1f} ms"


def t_c2(B):
    d = W.make_c2(n=512, B=B)
    p2 = ctx.program(d["ops"])
    ctx.lml_batched(p2, d["X"], d["y"], d["Theta"][:8], 0.0)
    t = time.time()
    lml, info = ctx.lml_batched(p2, d["X"], d["y"], d["Theta"], 0.0)
    dt = time.time() - t
    CO.use_openblas(1)
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"][:16], 0.0)
    CO.use_plain_c()
    return f"rel {np.max(np.abs(lml[:16] - ref) / np.abs(ref)):.2e} bad {int((info != 0).sum())} {B / dt:.0f} evals/s e2e ({dt * 1e3:.1f} ms)"


for n in (5, 64, 65, 130):
    step(f"cov n={n}", lambda n=n: t_cov(n))
for n in (5, 64, 65, 130, 300):
    step(f"lml n={n}", lambda n=n: t_lml(n))
for n in (20, 130):
    step(f"grad n={n}", lambda n=n: t_grad(n))
for n, v in ((50, 0), (130, 0), (130, 1), (700, 0)):
    step(f"posterior n={n} variant={v}", lambda n=n, v=v: t_post(n, v))
step("sample n=200", lambda: t_sample(200))
for n in (100, 257, 1000):
    step(f"chol n={n}", lambda n=n: t_chol(n))
step("lml_large n=1500", lambda: t_large(1500))
step("c2 B=512", lambda: t_c2(512))
step("c2 B=4096", lambda: t_c2(4096))
if "--big" in sys.argv:
    step("lml_large n=8192", lambda: t_large(8192))
print("launches", ctx.launch_count())
