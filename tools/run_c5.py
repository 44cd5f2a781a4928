"""C5 alone: gpl_lml_large at n = 8192 with the library's own per-phase CUDA-event times (covariance build, factorisation +
forward solve) next to the wall time of the blocking call.   python tools/run_c5.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaplac_b200 import _lib, workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = W.make_c5(n=n)
ctx = _lib.Context(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
prog = ctx.program(d["ops"])
ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
ctx.set_option("profile_events", 1)
for _ in range(5):
    t = time.perf_counter()
    r = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    wall = (time.perf_counter() - t) * 1e3
    ms, _ = ctx.last_timing()
    print(f"n={n} lml={r[0]:.6f} info={r[2]}  cov build {ms[0]:.3f} ms  factorisation {ms[1]:.3f} ms = {n ** 3 / 3 / ms[1] * 1e-9:.2f} TF  wall {wall:.3f} ms")
