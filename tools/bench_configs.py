#!/usr/bin/env python3
"""Timings of the other BASELINE.json configs (C1, C3, golden, C4, C5) through the host-buffer C ABI (wall clock around
blocking calls, best of a few repetitions, after a warm-up call).  The headline config C2 is bench.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaplac_b200 import _lib, workloads as W  # noqa: E402
from oracle import gp_oracle as O             # noqa: E402  (golden fixture loader only)


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


ctx = _lib.Context(0)
out = {}

d = W.make_c1()
prog = ctx.program(d["ops"])
fx = np.random.default_rng(0).standard_normal(50)
t = best(lambda: ctx.lml_batched(prog, d["X"], fx, np.array([[2.5]]), 0.1, grad=True), 5)
out["C1 logdensity+gradient call (n=50)"] = {"ms": t * 1e3, "calls_per_s": 1 / t}

d = W.make_c3()
prog = ctx.program(d["ops"])
t = best(lambda: ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0))
out["C3 2000 features x n=300 lml"] = {"ms": t * 1e3, "evals_per_s": 2000 / t, "tflops": 2000 * (300 ** 3 / 3 + 2 * 300 ** 2) / t * 1e-12}
t = best(lambda: ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True), 2)
out["C3 2000 features x n=300 lml+gradient"] = {"ms": t * 1e3, "evals_per_s": 2000 / t}

X, y, Th, s2, lpi, prior = O.load_golden("3206", os.path.join(ROOT, "tests", "golden"))
prog = ctx.program(O.golden_program("3206"))
Th2 = np.tile(Th, (2, 1))
t = best(lambda: ctx.lml_batched(prog, X, y, Th2, 0.0, jitter=1e-9))
out["golden 3206 n=923 x 200 rows lml"] = {"ms": t * 1e3, "evals_per_s": 200 / t, "tflops": 200 * (923 ** 3 / 3 + 2 * 923 ** 2) / t * 1e-12}

d = W.make_c4()
prog = ctx.program(d["ops"])
post = [None]


def fit():
    if post[0] is not None:
        post[0].free()
    post[0] = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)


tf = best(fit, 6)
tp = best(lambda: post[0].mean_and_var(d["Xs"]), 4)
n, m = 2048, 20000
out["C4 posterior fit n=2048"] = {"ms": tf * 1e3, "tflops": n ** 3 / 3 / tf * 1e-12}
out["C4 predict 20000 points"] = {"ms": tp * 1e3, "points_per_s": m / tp, "tflops": (n * n * m + 2 * n * m) / tp * 1e-12}

d = W.make_c5()
prog = ctx.program(d["ops"])
t = best(lambda: ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0), 2)
out["C5 n=8192 lml_large (K build + Cholesky + solve)"] = {"ms": t * 1e3, "tflops": 8192 ** 3 / 3 / t * 1e-12,
                                                           "frac_fp64_peak": 8192 ** 3 / 3 / t * 1e-12 / 37.0}
print(json.dumps(out, indent=1))
