"""gpl_lml_large with FP64 (DMMA) trailing updates against option trail_int8 = S (INT8 split path on tcgen05): lml and the
library's own CUDA-event time of the factorisation + forward solve.   python tools/i8_large.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaplac_b200 import _lib, workloads as W  # noqa: E402

sizes = [int(v) for v in sys.argv[1:]] or [4096, 8192, 16384]
ctx = _lib.Context(0)
ctx.set_option("profile_events", 1)
out = []
for n in sizes:
    d = W.make_c5(n=n)
    prog = ctx.program(d["ops"])
    row = dict(n=n)
    base = None
    for S in (0, 8, 7, 6):
        ctx.set_option("trail_int8", S)
        best, lml, info = 1e30, None, None
        for _ in range(4):
            r = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
            ms, _ = ctx.last_timing()
            best = min(best, ms[1])
            lml, info = r[0], r[2]
        if S == 0:
            base = lml
        row[f"S{S}"] = dict(lml=lml, info=int(info), factor_ms=best, tf_equiv=n ** 3 / 3 / best * 1e-9, rel_vs_fp64=abs(lml - base) / abs(base))
        print(f"n={n:6d} trail_int8={S}: lml {lml:.10f} info {info} factor {best:8.3f} ms = {n ** 3 / 3 / best * 1e-9:6.2f} TF-equivalent"
              f"  rel diff vs FP64 path {abs(lml - base) / abs(base):.2e}", flush=True)
    out.append(row)
ctx.set_option("trail_int8", 0)
if os.path.isdir("gpurun_out"):
    json.dump(out, open("gpurun_out/i8_large.json", "w"), indent=1)
