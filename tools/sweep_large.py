"""n-sweep of the single-model paths against cuSOLVER: gpl_lml_large (covariance build + blocked Cholesky + forward solve +
logdet) and gpl_posterior_fit (+ alpha, diagonal-tile inverses) for n = 256 .. 8192, next to torch.linalg.cholesky
(cuSOLVER potrf) on the same matrix.     python tools/sweep_large.py > profiles/sweep_large_r02.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gaplac_b200 import _lib, workloads as W
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def ev(fn, reps=5):
        fn()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    print(f"{'n':>6s} {'lml_large ms':>13s} {'factor ms':>10s} {'TF':>6s} {'% peak':>7s} {'posterior_fit ms':>17s} {'cuSOLVER potrf ms':>18s} {'TF':>6s}   path")
    for n in (256, 512, 1024, 2048, 3072, 4096, 6144, 8192):
        d = W.make_c5(n=n)
        prog = ctx.program(d["ops"])
        whole = ev(lambda: ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0))
        ctx.set_option("profile_events", 1)
        fac = []
        for _ in range(4):
            ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
            fac.append(ctx.last_timing()[0][1])
        ctx.set_option("profile_events", 0)
        f_ms = float(np.median(fac))
        holder = {}

        def fit():
            if "p" in holder:
                holder["p"].free()
            holder["p"] = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)

        fit_ms = ev(fit)
        holder["p"].free()
        K = torch.from_numpy(ctx.cov(prog, d["X"], d["theta"], 0.0)).to(dev)
        cs = ev(lambda: torch.linalg.cholesky(K))
        tf = n ** 3 / 3.0 / (f_ms * 1e-3) * 1e-12
        nt = (n + 63) // 64
        path = "fused one-CTA kernel (fit)" if n <= 128 else ("panels on one stream" if (nt + 3) // 4 <= 2 else "look-ahead panels + worker CTA")
        print(f"{n:6d} {whole:13.3f} {f_ms:10.3f} {tf:6.2f} {100 * tf / 37.0:7.1f} {fit_ms:17.3f} {cs:18.3f} {n ** 3 / 3.0 / (cs * 1e-3) * 1e-12:6.2f}   {path}")
    ctx.set_stream(0)
    print("\nlml_large / posterior_fit: CUDA events on the library's stream around the blocking host-buffer calls (H2D of X, y and D2H of the "
          "scalars inside); factor: the library's own events around the factorisation + forward solve; cuSOLVER: torch.linalg.cholesky of the "
          "same K resident in HBM (factorisation only).")


if __name__ == "__main__":
    main()
