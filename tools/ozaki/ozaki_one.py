"""One shape, a few launches of the split + update pair (the command profiled under ncu).  python tools/ozaki/ozaki_one.py n K S [reps]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ozaki_bench as ob  # noqa: E402

n, K, S = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
A = torch.randn(n, K, dtype=torch.float64, device="cuda")
C = ob.colmajor(n, n, 0.0)
ws = None
for _ in range(reps):
    ws = ob.ozaki(A, C, S, 0, ws)
ob.sync("one")
print("ok")
