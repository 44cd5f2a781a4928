timeout 400 python -m pytest tests/test_chain.py -m gpu -x -q --timeout 200 2>&1 | tail -8
python - <<PY
import os, time, numpy as np, gaplac_b200 as G
from gaplac_b200.formula import KernelProgram
from oracle import gp_oracle as O
X, y, Th, s2, lpi, prior = O.load_golden("3206", "tests/golden")
gp = G.GP(KernelProgram(ops=O.golden_program("3206"), vars=["PersonID","StoolPairs","nutrient"], n_theta=4))
nut = np.arange(-5.0, 5.0001, 0.1); Xs = np.column_stack([np.zeros_like(nut), np.zeros_like(nut), nut])
for batched in (True, False):
    G.predict_chain(gp, X, y, Th[:3], Xs, jitter=1e-9, obs_var=Th[:3,3], batched=batched)
    t=time.perf_counter(); out = G.predict_chain(gp, X, y, Th, Xs, jitter=1e-9, obs_var=Th[:,3], batched=batched); dt=time.perf_counter()-t
    print("batched" if batched else "looped ", "predict over %d chain rows x %d points (n=923): %.1f ms" % (len(Th), len(nut), dt*1e3), out["ymu"][:2], out["yQ950"][:2])
ctx = G.default_context(); prog = gp.compiled(ctx)
ctx.predict_batched(prog, X, y, Th, 0.0, Xs, 1e-9)
t=time.perf_counter(); ctx.predict_batched(prog, X, y, Th, 0.0, Xs, 1e-9); print("gpl_predict_batched alone: %.2f ms" % ((time.perf_counter()-t)*1e3))
PY
