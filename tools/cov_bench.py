"""Covariance-construction kernels alone (north_star kernel 1): achieved HBM write bandwidth at n = 8192.

    python tools/cov_bench.py [--n 8192] [--reps 20]
    ncu --set full -k regex:cov_ ... python tools/cov_bench.py --reps 2     (profiles/ncu_r02_cov_*.txt)

cov_dense_kernel writes the full n x n matrix (8 n^2 bytes), cov_tiles_kernel the lower triangle in the tile-major layout
of the large-n factorisation (8 * 4096 * nt(nt+1)/2 bytes).  Programs: SqExp (one exp per entry), SqExp+OU (two), and the
C5 model SqExp+Noise."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gaplac_b200 import _lib, workloads as W
    from gaplac_b200.formula import Op
    from gaplac_b200._lib import SQEXP
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--profile", action="store_true", help="one launch per kernel / program, no warm-up (for ncu -c 4)")
    a = ap.parse_args()
    n = a.n
    if a.profile:
        a.reps = 1
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    rng = np.random.default_rng(5)
    x = rng.uniform(-50, 50, n)
    dX = torch.from_numpy(x).to(dev)
    dK = torch.empty(n * n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6550.7
    out = {"n": n, "hbm_peak_gbs": peak, "kernels": {}}
    progs = {"SqExp": ([Op(SQEXP, col=0, theta_slot=0)], [1.0]), "SqExp+OU+Noise (C2)": (W.prog_c2(), [1.5, 1.0, 0.1]),
             "SqExp+Noise (C5)": (W.prog_c5(), [1.0, 0.1])}
    for name, (ops, th) in progs.items():
        prog = ctx.program(ops)
        dth = torch.tensor(th, dtype=torch.float64, device=dev)
        for _ in range(0 if a.profile else 3):
            ctx.cov_dev(prog, n, 1, dX.data_ptr(), dth.data_ptr(), len(th), 0.0, 0.0, dK.data_ptr(), stream.cuda_stream)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for _ in range(a.reps):
            ctx.cov_dev(prog, n, 1, dX.data_ptr(), dth.data_ptr(), len(th), 0.0, 0.0, dK.data_ptr(), stream.cuda_stream)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / a.reps
        gbs = 8.0 * n * n / (ms * 1e-3) * 1e-9
        out["kernels"]["cov_dense_kernel " + name] = {"ms": ms, "gb_written": 8.0 * n * n * 1e-9, "gbs": gbs, "frac_hbm": gbs / peak}
    # tile-major build inside gpl_lml_large (phase 0 of its timing)
    d = W.make_c5(n=n)
    prog = ctx.program(d["ops"])
    if not a.profile:
        ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
    ctx.set_option("profile_events", 1)
    ms = []
    for _ in range(1 if a.profile else max(2, a.reps // 4)):
        ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
        ms.append(ctx.last_timing()[0][0])
    ctx.set_option("profile_events", 0)
    nt = (n + 63) // 64
    nbytes = 8.0 * 4096 * nt * (nt + 1) / 2
    m = float(np.mean(ms))
    out["kernels"]["cov_tiles_kernel SqExp+Noise (C5)"] = {"ms": m, "gb_written": nbytes * 1e-9, "gbs": nbytes / (m * 1e-3) * 1e-9,
                                                          "frac_hbm": nbytes / (m * 1e-3) * 1e-9 / peak}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
