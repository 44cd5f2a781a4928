#!/usr/bin/env python3
"""Small-batch sweep: lockstep schedule (lml_variant 0) vs fused per-item kernel (1) through the host-buffer ABI."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaplac_b200 import _lib, workloads as W

ctx = _lib.Context(0)
for n in (128, 512, 1024):
    d = W.make_c2(n=n, B=4096)
    prog = ctx.program(d["ops"])
    for B in (1, 8, 64, 256, 1024, 4096):
        th = np.ascontiguousarray(d["Theta"][:B])
        row = []
        for v in (0, 1):
            ctx.set_option("lml_variant", v)
            f = lambda: ctx.lml_batched(prog, d["X"], d["y"], th, 0.0)
            f(); f()
            t = min((lambda t0: (f(), time.perf_counter() - t0)[1])(time.perf_counter()) for _ in range(5))
            row.append(t * 1e3)
        print(f"n={n:5d} B={B:5d}  lockstep {row[0]:8.3f} ms   fused {row[1]:8.3f} ms")
