import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaplac_b200 import _lib, workloads as W
ctx = _lib.Context(0)
d = W.make_c3(); prog = ctx.program(d["ops"])
for _ in range(2):
    r = ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
print(r[0][:3])
