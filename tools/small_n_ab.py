"""A/B of the two value+gradient paths for one-tile models (n <= 64): fused per-item kernel (one launch, the default) vs the
lockstep schedule (six launches; option lml_variant = 3), over batch sizes; plus the README sampler either way."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gaplac_b200 import _lib, mcmc, workloads as W
    ctx = _lib.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    out = {}
    for n in (50, 64):
        dd = W.make_c1(n=n)
        X = torch.from_numpy(np.ascontiguousarray(dd["X"].T)).to(dev)
        for B in (1, 64, 1024, 16384):
            rng = np.random.default_rng(B)
            Y = torch.from_numpy(rng.standard_normal((B, n))).to(dev)
            Th = torch.from_numpy(rng.uniform(0.5, 5, (B, 1))).to(dev)
            s2 = torch.tensor([0.1], dtype=torch.float64, device=dev)
            lml = torch.empty(B, dtype=torch.float64, device=dev)
            dth = torch.empty(B, dtype=torch.float64, device=dev)
            dy = torch.empty(B * n, dtype=torch.float64, device=dev)
            info = torch.zeros(B, dtype=torch.int32, device=dev)
            res = {}
            for variant in (0, 3):
                ctx.set_option("lml_variant", variant)

                def call():
                    ctx.lml_batched_dev(prog, n, 1, X.data_ptr(), False, Y.data_ptr(), True, Th.data_ptr(), 1, s2.data_ptr(), False,
                                        0.0, B, lml.data_ptr(), dth.data_ptr(), dy.data_ptr(), info.data_ptr(), stream.cuda_stream)
                for _ in range(5):
                    call()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(50):
                    call()
                b.record(stream)
                torch.cuda.synchronize()
                res["fused" if variant == 0 else "lockstep"] = a.elapsed_time(b) / 50
                res["lml_" + ("fused" if variant == 0 else "lockstep")] = lml.cpu().numpy().copy()
            out[f"n={n} B={B}"] = {"fused_ms": res["fused"], "lockstep_ms": res["lockstep"],
                                   "max_rel_diff": float(np.max(np.abs(res["lml_fused"] - res["lml_lockstep"]) / np.abs(res["lml_lockstep"])))}
    for variant in (0, 3):
        ctx.set_option("lml_variant", variant)
        mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=20, n_adapt=10, seed=1)
        for chains in (1, 64):
            r = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=500, seed=1, chains=chains)
            out[f"c1_mcmc chains={chains} {'fused' if variant == 0 else 'lockstep'}"] = {"seconds": r["seconds"], "grad_evals": int(r["grad_evals"])}
    ctx.set_option("lml_variant", 0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
