import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaplac_b200 import _lib, workloads as W
import ctypes as C
ctx = _lib.Context(0)
ctx.set_option('profile_events', int(os.environ.get('PROF', '1')))
if os.environ.get('CHOLV'): ctx.set_option('chol_variant', int(os.environ['CHOLV']))
lib = _lib.load(); lib.gpl_debug_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = W.make_c5(n=n); prog = ctx.program(d["ops"])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t = time.perf_counter(); r = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0); wall = (time.perf_counter() - t) * 1e3; ms = (C.c_double * 3)(); l3 = (C.c_int * 3)(); lib.gpl_debug_last_timing(ctx.h, ms, l3)
    print(r, "wall %.3f ms; cov build %.3f ms; factorisation %.3f ms = %.2f TFLOP/s (%.1f%% of 37.0)" % (wall, ms[0], ms[1], n**3/3/ms[1]*1e-9, n**3/3/ms[1]*1e-9/37.0*100))
