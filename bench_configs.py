"""The `configs` block of bench.py: the BASELINE.json configs other than the headline, each timed with CUDA events on
the stream the library runs on and checked against the CPU oracle on a sample (outside the timed regions).

    C2 +gradient   4096 proposals at n=512, lml + dtheta + dy            (the mcmc inner loop at the headline shape)
    C3             2000 features x n=300, Cat*SqExp+Noise: lml, lml+gradient
    golden         the reference's legacy fixture model, n=923, 200 chain rows (answers = the fixture's own values)
    C1             README workflow: one log-density + gradient call of the mcmc model body, n=50 (host-buffer ABI latency)
    C1 mcmc        the README command itself: gaplac mcmc "y ~| SqExp(:x)" --samples 500 on the device sampler (1 chain, and 64)
    C3 mcmc        one chain per feature: 256 of the 2000 chains x n=300 in lockstep (50 draws after 100 warm-up: bounded sample)
    C4             posterior fit n=2048 + mean/variance at 20 000 test points
    C5             single large GP n=8192: covariance build + blocked Cholesky + solve (gpl_lml_large)

Not part of the headline `value`; N = 1 only.  The oracle is used as the checker only."""
from __future__ import annotations

import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _events(torch, stream, fn, reps, flush=None):
    """Average CUDA-event time of fn() on `stream` (fn enqueues on it, or blocks on work enqueued on it)."""
    ms = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / reps


def run_configs(ctx, dev, flush, quick: bool = False):
    import torch
    from bench import DevArm, algorithmic_flops, fp64_peak, phase_times, workload, PHASES, host_cores, cpu_lml
    from gaplac_b200 import workloads as W
    from oracle import c_oracle as CO, gp_oracle as O

    peak, _ = fp64_peak()
    stream = torch.cuda.current_stream()
    out = {}

    def batched(name, wl, grad, reps=3):
        prog = ctx.program(wl["ops"])
        arm = DevArm(ctx, prog, wl, dev, grad)
        for _ in range(2):
            arm.launch(stream)
        ms = _events(torch, stream, lambda: arm.launch(stream), reps, flush)
        kms, _ = phase_times(arm, flush, 2)
        fl = arm.B * algorithmic_flops(arm.n, grad)
        e = {"workload": wl["name"], "n": arm.n, "batch": arm.B, "gradient": grad, "ms": ms, "evals_per_s": arm.B / (ms * 1e-3),
             "tflops": fl / (ms * 1e-3) * 1e-12, "frac_fp64_peak": fl / (ms * 1e-3) * 1e-12 / peak,
             "per_kernel_ms": {nm: float(v) for nm, v in zip(PHASES, kms) if v > 0},
             "not_pd_items": int((arm.dinfo != 0).sum().item())}
        lml = arm.dlml.cpu().numpy()
        idx = np.unique(np.linspace(0, arm.B - 1, 4).astype(int))
        if "known" in wl:       # the fixture's own answers
            e["max_abs_err_vs_reference_fixture"] = float(np.max(np.abs(lml - wl["known"])))
        ref = cpu_lml(wl, idx, min(host_cores(), len(idx)))
        e["oracle_max_rel_err"] = float(np.max(np.abs(lml[idx] - ref) / np.abs(ref)))
        if grad:
            dth = arm.ddth.cpu().numpy().reshape(arm.B, arm.p)
            dy = arm.ddy.cpu().numpy().reshape(arm.B, arm.n)
            b = int(idx[1])
            yb = wl["Y"] if wl["Y"].ndim == 1 else wl["Y"][b]
            _, rdth, rdy = O.lml_grad(wl["ops"], wl["X"], yb, wl["Theta"][b], float(wl["sigma2"][0]), wl["jitter"])
            e["oracle_grad_max_rel_err"] = float(max(np.max(np.abs(dth[b] - rdth) / np.maximum(1.0, np.abs(rdth))),
                                                     np.max(np.abs(dy[b] - rdy)) / max(1.0, np.max(np.abs(rdy)))))
        out[name] = e

    batched("c2_grad", workload("c2", 0), True)
    c3 = workload("c3", 0)
    batched("c3_lml", c3, False)
    batched("c3_grad", c3, True)
    batched("golden_n923", workload("golden", 0), False)

    # ---- C1: README workflow, one log-density + gradient call of the mcmc model body through the host-buffer ABI ----------
    d = W.make_c1()
    prog = ctx.program(d["ops"])
    fx = np.random.default_rng(0).standard_normal(50)
    th = np.array([[2.5]])
    for _ in range(20):
        res = ctx.lml_batched(prog, d["X"], fx, th, d["sigma2"], grad=True)
    t0 = time.perf_counter()
    for _ in range(200):
        res = ctx.lml_batched(prog, d["X"], fx, th, d["sigma2"], grad=True)
    ms = (time.perf_counter() - t0) * 1e3 / 200
    _, rdth, rdy = O.lml_grad(d["ops"], d["X"], fx, th[0], d["sigma2"])
    ref = CO.lml(d["ops"], d["X"], fx, th[0], d["sigma2"])[0]
    out["c1_call"] = {"workload": "C1 y ~| SqExp(:x), n=50: one logpdf + gradient call (host buffers, blocking)", "ms": ms,
                      "calls_per_s": 1e3 / ms, "timing": "wall clock over 200 blocking calls (latency-bound: 13 launches + 2 copies)",
                      "oracle_max_rel_err": float(abs(res[0][0] - ref) / abs(ref)),
                      "oracle_grad_max_rel_err": float(max(abs(res[2][0, 0] - rdth[0]) / max(1.0, abs(rdth[0])),
                                                           np.max(np.abs(res[3][0] - rdy))))}

    # ---- C1 / C3 through the on-device sampler (gpl_mcmc_nuts) ------------------------------------------------------------------
    from gaplac_b200 import mcmc
    from oracle import nuts_ref as NR

    def sampler_entry(label, res, n_chains, n, note):
        T = res["n_adapt"] + res["n_samples"]
        return {"workload": label, "chains": n_chains, "n": n, "n_samples": res["n_samples"], "n_adapt": res["n_adapt"],
                "seconds": res["seconds"], "samples_per_s": n_chains * res["n_samples"] / res["seconds"],
                "transitions_per_s": n_chains * T / res["seconds"], "grad_evals": int(res["grad_evals"]),
                "grad_evals_per_s": res["grad_evals"] / res["seconds"], "mean_accept": float(res["accept"].mean()),
                "mean_tree_depth": float(res["depth"].mean()), "max_tree_depth": int(res["depth"].max()),
                "divergent_frac": float(res["divergent"].mean()),
                "leapfrogs_per_chain_mean_max": [float(res["n_leapfrog"].sum(axis=1).mean()), int(res["n_leapfrog"].sum(axis=1).max())],
                "failed_chains": int((res["status"] != 0).sum()), "timing": "wall clock of the blocking call (H2D, graph replay of "
                "(batched lml+gradient, chain state machine) per leapfrog step, D2H of the chains)", "note": note}

    d = W.make_c1()
    prog = ctx.program(d["ops"])
    mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=20, n_adapt=10, seed=1)      # warm the graph path
    r1 = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=500, seed=1, chains=1)
    e = sampler_entry("C1 gaplac mcmc \"y ~| SqExp(:x)\" n=50, 500 samples (+250 warm-up), 1 chain", r1, 1, 50,
                      "the reference's own configuration: one chain; latency-bound (7 small launches per leapfrog step)")
    # fixed-seed parity: the first transitions of the same chain on the host reference sampler
    rw = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=4, n_adapt=8, seed=1, chains=1,
                   record_warmup=True, record_q=True)
    model = NR.Model(d["ops"], d["X"], d["y"], np.array([0.0]), np.array([20.0]), 0.1)
    ref = NR.sample_chain(model, np.zeros(model.dim), 4, 8, seed=1, chain=0)
    e["reference_sampler_max_rel_err_first_6_transitions"] = float(max(
        np.max(np.abs(rw[k][0][:6] - ref[k][:6]) / np.maximum(1.0, np.abs(ref[k][:6]))) for k in ("q", "theta", "lp", "eps", "accept")))
    e["reference_sampler_same_trees_first_12"] = bool(np.array_equal(rw["depth"][0][:12], ref["depth"][:12]))
    out["c1_mcmc"] = e
    r64 = mcmc.nuts(ctx, prog, d["X"], d["y"], [0.0], [20.0], sigma2=0.1, n_samples=500, seed=1, chains=64)
    out["c1_mcmc_64_chains"] = sampler_entry("C1 model, 64 chains x 500 samples in lockstep", r64, 64, 50, "same launches, 64 chains")
    nch = 256
    r3 = mcmc.nuts(ctx, ctx.program(c3["ops"]), c3["X"], c3["Y"][:nch], [0.0, 0.0], [100.0, 2.0], sigma2=0.0, n_samples=50, n_adapt=100, seed=3)
    out["c3_mcmc"] = sampler_entry(f"C3 {nch} of the 2000 feature chains x n=300, Cat*SqExp+Noise, l ~ U(0,100), s2 ~ U(0,2): 50 draws + 100 warm-up each",
                                   r3, nch, 300, "bounded sample of the config (an eighth of the features, short chains); finished chains are compacted out of the batch; the wall time of a lockstep batch is set by its slowest chain (deepest trees)")

    # ---- C4: posterior fit n=2048 + 20 000 test points (host-buffer ABI on torch's stream: CUDA events see it) -------------
    d = W.make_c4()
    prog = ctx.program(d["ops"])
    ctx.set_stream(stream.cuda_stream)
    try:
        post = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)
        post.free()
        holder = {}

        def fit():
            if "p" in holder:
                holder["p"].free()
            holder["p"] = ctx.posterior_fit(prog, d["X"], d["y"], d["theta"], 0.0)

        fit_ms = _events(torch, stream, fit, 3)
        post = holder["p"]
        post.mean_and_var(d["Xs"])
        res = {}

        def pred():
            res["mv"] = post.mean_and_var(d["Xs"])

        pred_ms = _events(torch, stream, pred, 3)
        n, m = 2048, len(d["Xs"])
        e = {"workload": "C4 SqExp(:x)+Linear(:z)+Noise n=2048 train, 20000 test points", "fit_ms": fit_ms,
             "fit_tflops": n ** 3 / 3.0 / (fit_ms * 1e-3) * 1e-12, "predict_ms": pred_ms, "points_per_s": m / (pred_ms * 1e-3),
             "predict_tflops": (n * n * m + 2.0 * n * m) / (pred_ms * 1e-3) * 1e-12,
             "predict_frac_fp64_peak": (n * n * m + 2.0 * n * m) / (pred_ms * 1e-3) * 1e-12 / peak,
             "timing": "CUDA events on the library's stream around the blocking host-buffer calls (H2D/D2H inside)"}
        idx = np.unique(np.linspace(0, m - 1, 16).astype(int))
        U, alpha = CO.posterior(d["ops"], d["X"], d["y"], d["theta"], 0.0)
        rm, rv = CO.mean_and_var(d["ops"], d["X"], U, alpha, np.ascontiguousarray(d["Xs"][idx]), d["theta"])
        mean, var = res["mv"]
        e["oracle_mean_max_err"] = float(np.max(np.abs(mean[idx] - rm)) / max(1.0, np.max(np.abs(rm))))
        e["oracle_var_max_err"] = float(np.max(np.abs(var[idx] - rv)) / max(1.0, np.max(np.abs(rv))))
        out["c4_predict"] = e
        post.free()

        # ---- C5: n = 8192 --------------------------------------------------------------------------------------------------
        d = W.make_c5()
        prog = ctx.program(d["ops"])
        ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
        ctx.set_option("profile_events", 1)
        fac, cov = [], []
        for _ in range(3):
            lml, ld, info = ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0)
            ms7, _ = ctx.last_timing()
            cov.append(ms7[0])
            fac.append(ms7[1])
        ctx.set_option("profile_events", 0)
        whole_ms = _events(torch, stream, lambda: ctx.lml_large(prog, d["X"], d["y"], d["theta"], 0.0), 3)
        n = 8192
        f_ms = float(np.mean(fac))

        def factor_ms(nn, option, reps=3):  # the library's own events around the factorisation, for one option value
            dd = d if nn == 8192 else W.make_c5(n=nn)
            ctx.set_option("trail_int8", option)
            ctx.set_option("profile_events", 1)
            ts, val = [], None
            for _ in range(reps + 1):
                val = ctx.lml_large(prog, dd["X"], dd["y"], dd["theta"], 0.0)
                ts.append(ctx.last_timing()[0][1])
            ctx.set_option("profile_events", 0)
            ctx.set_option("trail_int8", -1)
            return float(np.mean(ts[1:])), val

        dmma_ms, dmma_val = factor_ms(8192, 0)
        eight_ms, eight_val = factor_ms(8192, 8)
        e = {"workload": "C5 SqExp(:x; l=1)+Noise n=8192 single model", "factorisation_ms": f_ms,
             "factorisation_tflops": n ** 3 / 3.0 / (f_ms * 1e-3) * 1e-12,
             "factorisation_frac_fp64_peak": n ** 3 / 3.0 / (f_ms * 1e-3) * 1e-12 / peak,
             "trailing_updates": "INT8 split path on tcgen05 (option trail_int8 = -1: 9 slices from n = 6144 on, as accurate as the FP64 "
                                 "path on ill-conditioned covariances too); FLOPs are those of the FP64 factorisation it replaces",
             "int8_8_slices": {"factorisation_ms": eight_ms, "factorisation_tflops": n ** 3 / 3.0 / (eight_ms * 1e-3) * 1e-12,
                               "lml_rel_diff": float(abs(eight_val[0] - dmma_val[0]) / abs(dmma_val[0])),
                               "note": "56 bits below the row maxima: within 1e-14 of the FP64 path on well-conditioned problems like this one"},
             "fp64_dmma_only": {"factorisation_ms": dmma_ms, "factorisation_tflops": n ** 3 / 3.0 / (dmma_ms * 1e-3) * 1e-12,
                                "lml_rel_diff": float(abs(dmma_val[0] - lml) / abs(lml))},
             "cov_build_ms": float(np.mean(cov)), "cov_build_gbs": 8.0 * (n * (n + 64) / 2) / (np.mean(cov) * 1e-3) * 1e-9,
             "whole_call_ms": whole_ms, "info": int(info),
             "timing": "CUDA events inside the library around the build and the factorisation (+ forward solve); whole call by "
                       "events on the stream around the blocking host-buffer call"}
        if not quick:  # beyond the north_star size: where the trailing updates are most of the work
            big = {}
            for nn in (12288, 16384):
                a_ms, a_val = factor_ms(nn, -1, reps=2)
                b_ms, b_val = factor_ms(nn, 0, reps=2)
                big[f"n{nn}"] = {"factorisation_ms": a_ms, "factorisation_tflops": nn ** 3 / 3.0 / (a_ms * 1e-3) * 1e-12,
                                 "factorisation_frac_fp64_peak": nn ** 3 / 3.0 / (a_ms * 1e-3) * 1e-12 / peak,
                                 "fp64_dmma_only_ms": b_ms, "lml_rel_diff_vs_fp64_path": float(abs(a_val[0] - b_val[0]) / abs(b_val[0])),
                                 "info": int(a_val[2])}
            e["larger_n"] = big
        if not quick:
            import scipy.linalg as sla
            K = O.cov(d["ops"], d["X"], d["theta"], 0.0)
            c, _ = sla.cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
            z = sla.solve_triangular(c, d["y"], lower=True, check_finite=False)
            rld = 2.0 * np.sum(np.log(np.diag(c)))
            ref = -0.5 * (n * np.log(2 * np.pi) + rld + float(z @ z))
            e["oracle_lml_rel_err"] = float(abs(lml - ref) / abs(ref))
            e["oracle_logdet_rel_err"] = float(abs(ld - rld) / abs(rld))
        out["c5_large_n"] = e
    finally:
        ctx.set_stream(0)
    return out
