#!/bin/bash
python -m pytest tests/test_gpu_mcmc.py tests/test_gpu_multi.py -x -q > gpurun_out/t5.log 2>&1; tail -4 gpurun_out/t5.log
python tools/sanitize_target.py > gpurun_out/sanitize_plain.out 2>&1 && \
compute-sanitizer --tool memcheck --leak-check no --log-file gpurun_out/sanitize_memcheck.log python tools/sanitize_target.py > gpurun_out/sanitize_memcheck.out 2>&1
echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitize_memcheck.out; tail -5 gpurun_out/sanitize_memcheck.log 2>/dev/null
python - <<'PY' > gpurun_out/mcmc_compaction.json 2> gpurun_out/mcmc_compaction.err
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from gaplac_b200 import _lib, mcmc, workloads as W
ctx = _lib.Context(0)
d = W.make_c3(features=512)
prog = ctx.program(d["ops"])
out = {}
for name, kw in (("c3_512_chains", dict(n_samples=50, n_adapt=100)),):
    r = mcmc.nuts(ctx, prog, d["X"], d["Y"], [0.0, 0.0], [100.0, 2.0], sigma2=0.0, seed=3, **kw)
    out[name] = dict(seconds=r["seconds"], grad_evals=int(r["grad_evals"]), accept=float(r["accept"].mean()),
                     depth=float(r["depth"].mean()), div=float(r["divergent"].mean()), status=int((r["status"] != 0).sum()),
                     leapfrogs_total=int(r["n_leapfrog"].sum()))
d1 = W.make_c1()
r = mcmc.nuts(ctx, ctx.program(d1["ops"]), d1["X"], d1["y"], [0.0], [20.0], sigma2=0.1, n_samples=500, seed=1, chains=64)
out["c1_64"] = dict(seconds=r["seconds"], grad_evals=int(r["grad_evals"]), accept=float(r["accept"].mean()))
print(json.dumps(out))
PY
cat gpurun_out/mcmc_compaction.json; tail -c 300 gpurun_out/mcmc_compaction.err
