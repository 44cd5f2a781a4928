#!/usr/bin/env python3
"""Per-phase clock totals of the fused lml kernel (profile build: -DGPL_LML_PROFILE).

    python -c "from gaplac_b200 import build as B; B.build_variant('prof', ['-DGPL_LML_PROFILE'])"
    GAPLAC_B200_LIB=gaplac_b200/libgaplac_b200_prof.so python tools/phase_profile.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaplac_b200 import _lib, workloads as W  # noqa: E402

ctx = _lib.Context(0)
d = W.make_c2()
prog = ctx.program(d["ops"])
dev = torch.device("cuda", 0)
X = torch.from_numpy(np.ascontiguousarray(d["X"].T)).to(dev)
Y = torch.from_numpy(d["y"]).to(dev)
Th = torch.from_numpy(np.ascontiguousarray(d["Theta"])).to(dev)
S2 = torch.zeros(1, dtype=torch.float64, device=dev)
B = Th.shape[0]
lml = torch.empty(B, dtype=torch.float64, device=dev)
info = torch.zeros(B, dtype=torch.int32, device=dev)
raw = torch.zeros(148 * 8 * 8, dtype=torch.float64, device=dev)
grid = C.c_int()
lib = _lib.load()
lib.gpl_debug_phase_profile.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
for _ in range(2):
    rc = lib.gpl_debug_phase_profile(ctx.h, prog.h, 512, 1, X.data_ptr(), Y.data_ptr(), Th.data_ptr(), 3, S2.data_ptr(), B,
                                     lml.data_ptr(), info.data_ptr(), raw.data_ptr(), C.byref(grid))
    assert rc == 0
torch.cuda.synchronize()
g = grid.value
r = raw.cpu().numpy()[: g * 8].reshape(g, 8)
names = ["top-of-tile barrier", "K-gen (covariance tile)", "update loop (DMMA)", "potrf (diag)", "diag tail (solve, stores)",
         "trsm", "store L_ij", "between tiles"]
tot = r.sum(axis=1).mean()
print(f"grid {g} CTAs; mean clocks per CTA {tot:.3e}")
for k, nme in enumerate(names):
    print(f"  {nme:28s} {100 * r[:, k].mean() / tot:6.2f}%   {r[:, k].mean() / (B / g):10.0f} clk per matrix")
