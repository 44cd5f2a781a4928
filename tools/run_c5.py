import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaplac_b200 import _lib, workloads as W
ctx = _lib.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = W.make_c5(n=n); prog = ctx.program(d["ops"])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t = time.perf_counter(); r = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0); print(r, (time.perf_counter() - t) * 1e3, "ms")
