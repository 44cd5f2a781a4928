#!/bin/bash
# ncu launch lists: gpl_lml_large n=8192 (INT8 trailing passes) and the C3 batched lml + gradient (zero-tile skipping)
timeout 120 python tools/run_c5.py 8192 > gpurun_out/c5_56.log 2>&1 || exit 1; tail -2 gpurun_out/c5_56.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c5_int8.csv python tools/run_c5.py 8192 > gpurun_out/ncu56.log 2>&1; echo rc=$?
cat > /tmp/c3run.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from gaplac_b200 import _lib, workloads as W
d = W.make_c3()
ctx = _lib.Context(0)
prog = ctx.program(d["ops"])
for _ in range(3):
    ctx.lml_batched(prog, d["X"], d["Y"], d["Theta"], 0.0, grad=True)
print("ok")
PY
timeout 120 python /tmp/c3run.py || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c3_grad.csv python /tmp/c3run.py > gpurun_out/ncu56b.log 2>&1; echo rc=$?
