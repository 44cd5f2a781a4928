#!/usr/bin/env python3
"""Headline benchmark: FP64 GP log-marginal-likelihood evals/sec (n=512, batch 4096) — BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|golden]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: B = 4096 hyperparameter proposals of
`SqExp(:x)+OU(:x)+Noise` at n = 512 (BASELINE config[1]), shared X and y.
Multi-GPU is weak scaling: every rank owns its own 4096 proposals; the only collective is the NCCL
all-gather of the per-rank log-densities.

  value     whole-job evals/s with inputs resident in HBM (device-pointer C-ABI entry, CUDA events on the
            launching stream, L2 flushed between timed steps, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call (gpl_lml_batched): H2D of X, y, Theta and D2H
            of lml/info inside the timed region
  roofline  the fused kernel against the FP64 pipe peak (measured with tools/fp64_peak.cu)
  cpu_baseline  the CPU oracle (C + OpenBLAS dpotrf/dtrtrs, the reference's LAPACK path) on the host cores

`--impl reference` times that CPU path alone (the reference itself is Julia; no julia binary exists here,
so the arm runs the oracle port — DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 GP log-marginal-likelihood evals/sec (n=512, batch 4096)"
UNIT = "evals/s"
NOMINAL_FP64_TFLOPS = 37.2   # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload(name: str, rank: int):
    from gaplac_b200 import workloads as W
    if name == "c2":
        d = W.make_c2(seed=2, n=512, B=4096)
        if rank:   # every rank evaluates its own proposals on the shared (X, y)
            rng = np.random.default_rng(1000 + rank)
            d["Theta"] = np.column_stack([rng.uniform(0.2, 5, 4096), rng.uniform(0.2, 5, 4096), rng.uniform(0.05, 0.5, 4096)])
        return dict(name="C2 SqExp(:x)+OU(:x)+Noise n=512 B=4096 theta-batch", ops=d["ops"], X=d["X"], Y=d["y"],
                    Theta=d["Theta"], sigma2=np.array([0.0]), jitter=0.0)
    if name == "c3":
        d = W.make_c3(seed=3 + rank)
        return dict(name="C3 Cat(:subject)*SqExp(:time)+Noise n=300 2000 features y-batch", ops=d["ops"], X=d["X"],
                    Y=d["Y"], Theta=d["Theta"], sigma2=np.array([0.0]), jitter=0.0)
    if name == "golden":
        from oracle import gp_oracle as O
        X, y, Th, s2, _, _ = O.load_golden("3206", os.path.join(ROOT, "tests", "golden"))
        Th = np.tile(Th, (2, 1))
        return dict(name="golden 3206 n=923 200 chain rows", ops=O.golden_program("3206"), X=X, Y=y, Theta=Th,
                    sigma2=np.array([0.0]), jitter=1e-9)
    raise SystemExit(f"unknown workload {name}")


def algorithmic_flops(n: int) -> float:
    return n ** 3 / 3.0 + 2.0 * n * n       # Cholesky + one triangular solve + K build as n^2


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_rate(wl, sample: int, threads: int) -> float:
    from oracle import c_oracle as CO
    CO.use_openblas(1)
    Th = wl["Theta"][:sample]
    Y = wl["Y"] if wl["Y"].ndim == 1 else wl["Y"][:sample]
    t0 = time.perf_counter()
    CO.lml_batched(wl["ops"], wl["X"], Y, Th, wl["sigma2"], wl["jitter"], threads=threads)
    return sample / (time.perf_counter() - t0)


def cpu_baseline(wl, target_s: float = 12.0):
    cores = host_cores()
    probe = max(cores, 8)
    r0 = cpu_rate(wl, probe, cores)
    sample = int(min(len(wl["Theta"]), max(probe, (r0 * target_s) // cores * cores)))
    rate = cpu_rate(wl, sample, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} of the {len(wl['Theta'])} evals of one step, one eval per core at a time, "
                      f"OpenBLAS dpotrf/dtrtrs threads=1 (C oracle: oracle/gp_oracle.c)"}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    wl = workload(args.workload, 0)
    cores = host_cores()
    r0 = cpu_rate(wl, max(cores, 8), cores)
    per_step_s = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    sample = int(min(len(wl["Theta"]), max(cores, (r0 * per_step_s) // cores * cores)))
    for _ in range(args.warmup):
        cpu_rate(wl, sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_rate(wl, sample, cores)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = sample / dt
    desc = (f"{sample} evals per step (bounded sample of the {len(wl['Theta'])}-eval batch), {cores} host threads, "
            "C oracle + OpenBLAS dpotrf/dtrtrs threads=1 per eval; Julia is not installed, so this is the port")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "sample_evals_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- GPU arm
def fp64_peak():
    """Measured FP64 pipe peak of this pool's B200 (tools/fp64_peak.cu, committed under profiles/), else nominal."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["peak_tflops"]), j.get("source", "profiles/fp64_peak.json")
        except Exception:
            pass
    return NOMINAL_FP64_TFLOPS, "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (no measurement committed yet)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel, from the committed ncu launch list (profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def kernel_flops(n: int, B: int):
    """Algorithmic FLOPs of one step split over the three kernels of the lockstep schedule:
    n^3/3 = sum over 64x64 tile operations: potrf t^3/3 per diagonal tile, trsm t^3 and gemm 2 t^3 j per tile
    below the diagonal of column j, syrk t^3 j per diagonal tile; the 2 n^2 of the covariance build and solve go
    to the kernels in proportion to the tiles they generate."""
    t, nt = 64.0, (n + 63) // 64
    below_tiles = nt * (nt - 1) // 2
    gemm = sum(j * (nt - 1 - j) for j in range(nt))
    syrk = sum(range(nt))
    ntri = nt * (nt + 1) // 2
    total = algorithmic_flops(n)
    tile3 = (n / nt) ** 3  # n need not be a multiple of 64: scale the tile cube so the parts sum to n^3/3
    scale = (n ** 3 / 3.0) / (tile3 * (nt / 3.0 + below_tiles + 2 * gemm + syrk))
    # lk_below_kernel also forms the diagonal tiles of columns 1..nt-1 (syrk + their covariance entries);
    # lk_diag_kernel is left with the covariance of tile (0, 0)
    f_below = scale * tile3 * (below_tiles + 2 * gemm + syrk) + 2.0 * n * n * (below_tiles + nt - 1) / ntri
    f_diag = 2.0 * n * n * 1 / ntri
    f_potrf = total - f_below - f_diag
    return {"below": B * f_below, "diag": B * f_diag, "potrf": B * f_potrf}


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from gaplac_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload(args.workload, rank)
    ctx = _lib.Context(local_rank)
    prog = ctx.program(wl["ops"])
    X = np.asfortranarray(wl["X"])
    n, d = X.shape
    Theta = np.ascontiguousarray(wl["Theta"])
    B, p = Theta.shape
    Y = np.ascontiguousarray(wl["Y"])
    y_batched = Y.ndim == 2

    # ---- device-resident arm -----------------------------------------------------------------------------
    dX = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)           # column-major n x d == C-order d x n
    dY = torch.from_numpy(Y).to(dev)
    dTh = torch.from_numpy(Theta).to(dev)
    dS2 = torch.from_numpy(wl["sigma2"]).to(dev)
    dlml = torch.empty(B, dtype=torch.float64, device=dev)
    dinfo = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_dev():
        ctx.lml_batched_dev(prog, n, d, dX.data_ptr(), False, dY.data_ptr(), y_batched, dTh.data_ptr(), p,
                            dS2.data_ptr(), False, wl["jitter"], B, dlml.data_ptr(), 0, 0, dinfo.data_ptr(),
                            stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, dlml)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush between timed steps (not timed)
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        ctx.lml_batched_dev(prog, n, d, dX.data_ptr(), False, dY.data_ptr(), y_batched, dTh.data_ptr(), p,
                            dS2.data_ptr(), False, wl["jitter"], B, dlml.data_ptr(), 0, 0, dinfo.data_ptr(),
                            stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, dlml)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    launches = ctx.launch_count() - l0
    if world > 1:
        dist.barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    bad = int((dinfo != 0).sum().item())
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = t.tolist()

    # ---- per-kernel pass (untimed for `value`): CUDA events around every launch of the lockstep schedule ----------
    import ctypes as C
    lib = _lib.load()
    lib.gpl_debug_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    ctx.set_option("profile_events", 1)
    kms = np.zeros(3)
    kl = np.zeros(3, dtype=np.int64)
    reps = max(2, min(args.steps, 5))
    for _ in range(reps):
        flush.zero_()
        ctx.lml_batched_dev(prog, n, d, dX.data_ptr(), False, dY.data_ptr(), y_batched, dTh.data_ptr(), p,
                            dS2.data_ptr(), False, wl["jitter"], B, dlml.data_ptr(), 0, 0, dinfo.data_ptr(),
                            stream.cuda_stream)
        ms3 = (C.c_double * 3)()
        l3 = (C.c_int * 3)()
        lib.gpl_debug_last_timing(ctx.h, ms3, l3)
        kms += np.array(list(ms3))
        kl += np.array(list(l3))
    ctx.set_option("profile_events", 0)
    kms /= reps
    kl //= reps
    torch.cuda.synchronize()

    # ---- end-to-end arm: host buffers through gpl_lml_batched -----------------------------------------------
    hlml = None

    def step_e2e():
        nonlocal hlml
        hlml, hinfo = ctx.lml_batched(prog, wl["X"], wl["Y"], Theta, wl["sigma2"], wl["jitter"])
        if world > 1:
            g = torch.from_numpy(hlml).to(dev)
            dist.all_gather_into_tensor(gathered, g)
            if rank == 0:
                gathered.cpu()
        return hinfo

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - e0) * 1e3 / args.steps
    t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_ms = t.item()
    clocks = sampler.stop(t0, time.perf_counter()) if sampler else None
    parity = float(np.max(np.abs(hlml - dlml.cpu().numpy()) / np.abs(hlml)))   # the two arms compute the same thing
    h2d = X.nbytes + Y.nbytes + Theta.nbytes + wl["sigma2"].nbytes
    d2h = B * 8 + B * 4

    if rank == 0:
        peak, peak_src = fp64_peak()
        flops = B * algorithmic_flops(n)
        step_ach = flops / (kern_ms * 1e-3) * 1e-12
        kf = kernel_flops(n, B)
        names = ["diag", "potrf", "below"]
        dom = int(np.argmax(kms))            # dominant kernel by device time
        dom_name = "lk_%s_kernel" % names[dom]
        launches_dom = max(int(kl[dom]), 1)
        ach = kf[names[dom]] / (kms[dom] * 1e-3) * 1e-12 if kms[dom] > 0 else 0.0
        out = {
            "metric": METRIC, "value": world * B / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "n": n, "batch_per_gpu": B, "global_batch": world * B,
                       "sharding": f"independent proposals, {B} per rank, one NCCL all-gather of lml per step",
                       "l2": "256 MiB buffer written between timed steps (L2 flush, untimed); per-CTA factor workspace "
                             "also exceeds L2", "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA and DFMA share one pipe on sm_100a; tools/pipe_mix.cu)",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu_traffic(),
                         "peak_source": peak_src, "kernel": dom_name,
                         "kernel_launches_per_step": launches_dom,
                         "kernel_avg_launch_ms": float(kms[dom]) / launches_dom,
                         "kernel_share_of_step": float(kms[dom] / max(kms.sum(), 1e-12)),
                         "algorithmic_flops_per_launch": kf[names[dom]] / launches_dom,
                         "per_kernel_ms_per_step": {nm: float(v) for nm, v in zip(names, kms)},
                         "whole_step": {"achieved": step_ach, "frac": step_ach / peak, "ms": kern_ms,
                                        "algorithmic_flops": flops}},
            "e2e": {"value": world * B / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e_ms,
                    "api": "gpl_lml_batched (caller-owned host buffers, blocking; the library packs them into one pinned block: one H2D + one D2H copy per step) via ctypes"},
            "gpu_launches": int(launches), "clocks": clocks, "not_pd_items": bad, "arms_max_rel_diff": parity,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "golden"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
