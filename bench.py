#!/usr/bin/env python3
"""Headline benchmark: FP64 GP log-marginal-likelihood evals/sec (n=512, batch 4096) — BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|golden] [--grad]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: B = 4096 hyperparameter proposals of
`SqExp(:x)+OU(:x)+Noise` at n = 512 (BASELINE config[1]), shared X and y; with --grad the step
also returns the analytic gradient (the mcmc inner loop, CLI/src/mcmc.jl:31-37).  The headline number is weak
scaling: every rank owns its own 4096 proposals; the only collective is the NCCL all-gather of the per-rank
log-densities.

  value     whole-job evals/s with inputs resident in HBM (device-pointer C-ABI entry, CUDA events on the
            launching stream, L2 flushed between timed steps, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call (gpl_lml_batched): H2D of X, y, Theta and D2H
            of lml/info inside the timed region
  roofline  the dominant kernel against the FP64 pipe peak (measured with tools/fp64_peak.cu)
  cpu_baseline  the CPU oracle (C + OpenBLAS dpotrf/dtrtrs, the reference's LAPACK path) on the host cores
  oracle_max_rel_err  sampled items of the GATHERED vector (every rank's block) against the C oracle, outside the timed region
  strong    the same global batch (4096 proposals; 2000 C3 features) split over the ranks: strong scaling
  configs   (N = 1 only) the other BASELINE configs, each with its own CUDA-event time and an oracle check:
            C2 with gradient, C3 lml and lml+gradient, golden n=923, C1 call latency, C4 fit + predict, C5 n=8192

`--impl reference` times the CPU path alone (the reference itself is Julia; no julia binary exists here,
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
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PHASES = ["diag", "potrf", "below", "winv", "minv", "alpha", "gradc"]


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload(name: str, rank: int):
    """The synthetic inputs of one rank (deterministic: any rank can rebuild any other rank's block for checking)."""
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
        g = W.make_golden(GOLDEN_DIR, "3206")
        return dict(name="golden 3206 n=923 200 chain rows", ops=g["ops"], X=g["X"], Y=g["y"], Theta=np.tile(g["Theta"], (2, 1)),
                    sigma2=np.array([0.0]), jitter=g["jitter"], known=np.tile(g["lml_known"], 2))
    raise SystemExit(f"unknown workload {name}")


def config_of(wl, grad: bool, world: int):
    """The `config` object: identical keys and values in both arms (ours / reference)."""
    n = wl["X"].shape[0]
    B = len(wl["Theta"])
    return {"workload": wl["name"] + (" +gradient" if grad else ""), "n": n, "batch_per_gpu": B, "global_batch": world * B,
            "gradient": bool(grad)}


def algorithmic_flops(n: int, grad: bool = False) -> float:
    """lml = Cholesky + one triangular solve + K build as n^2; with the gradient n^3 (adds K^-1)."""
    return (n ** 3 if grad else n ** 3 / 3.0) + 2.0 * n * n


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_lml(wl, idx, threads: int):
    """C-oracle log-densities of the items idx of a workload."""
    from oracle import c_oracle as CO
    CO.use_openblas(1)
    Th = wl["Theta"][idx]
    Y = wl["Y"] if wl["Y"].ndim == 1 else wl["Y"][idx]
    return CO.lml_batched(wl["ops"], wl["X"], Y, Th, wl["sigma2"], wl["jitter"], threads=threads)[0]


def cpu_rate(wl, sample: int, threads: int) -> float:
    t0 = time.perf_counter()
    cpu_lml(wl, np.arange(sample), threads)
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
            "C oracle + OpenBLAS dpotrf/dtrtrs threads=1 per eval; Julia is not installed, so this is the port"
            + ("; log-density only (the reference gets its gradient from ~5 ForwardDiff passes over the same factorisation)"
               if args.grad else ""))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl, args.grad, args.gpus),
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

    def window(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        for t, line in list(self.rows):
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

    def stop(self):
        if not self.proc:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()


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


def ncu_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            per = j.get("per_kernel", {})
            if kernel in per:
                return per[kernel].get("dram_bytes_per_launch")
            return j.get("dram_bytes_per_launch") if j.get("kernel", kernel) == kernel else None
        except Exception:
            return None
    return None


def kernel_flops(n: int, B: int, grad: bool):
    """Algorithmic FLOPs of one step split over the kernels of the lockstep schedule:
    n^3/3 = sum over 64x64 tile operations: potrf t^3/3 per diagonal tile, trsm t^3 and gemm 2 t^3 j per tile
    below the diagonal of column j, syrk t^3 j per diagonal tile; the 2 n^2 of the covariance build and solve go
    to the kernels in proportion to the tiles they generate.  Gradient phases: M = L^-1 is n^3/3 (the diagonal-tile
    inverses of lk_winv are nt t^3/3 of it), K^-1 = M'M is n^3/3 (lk_gradc; the dK contraction is not counted)."""
    t, nt = 64.0, (n + 63) // 64
    below_tiles = nt * (nt - 1) // 2
    gemm = sum(j * (nt - 1 - j) for j in range(nt))
    syrk = sum(range(nt))
    ntri = nt * (nt + 1) // 2
    total = algorithmic_flops(n)
    tile3 = (n / nt) ** 3  # n need not be a multiple of 64: scale the tile cube so the parts sum to n^3/3
    scale = (n ** 3 / 3.0) / (tile3 * (nt / 3.0 + below_tiles + 2 * gemm + syrk))
    f_below = scale * tile3 * (below_tiles + 2 * gemm + syrk) + 2.0 * n * n * (below_tiles + nt - 1) / ntri
    f_diag = 2.0 * n * n * 1 / ntri
    f_potrf = total - f_below - f_diag
    out = {"below": B * f_below, "diag": B * f_diag, "potrf": B * f_potrf, "winv": 0.0, "minv": 0.0, "alpha": 0.0, "gradc": 0.0}
    if grad:
        f_winv = scale * tile3 * nt / 3.0
        out.update(winv=B * f_winv, minv=B * (n ** 3 / 3.0 - f_winv), alpha=0.0, gradc=B * n ** 3 / 3.0)
    return out


class DevArm:
    """One workload resident in HBM, evaluated through the device-pointer entry on torch's current stream."""

    def __init__(self, ctx, prog, wl, dev, grad: bool, lo: int = 0, hi: int | None = None):
        import torch
        self.ctx, self.prog, self.wl, self.grad = ctx, prog, wl, grad
        X = np.asfortranarray(wl["X"])
        self.n, self.d = X.shape
        hi = len(wl["Theta"]) if hi is None else hi
        Theta = np.ascontiguousarray(wl["Theta"][lo:hi])
        self.B, self.p = Theta.shape
        Y = wl["Y"] if wl["Y"].ndim == 1 else wl["Y"][lo:hi]
        Y = np.ascontiguousarray(Y)
        self.y_batched = Y.ndim == 2
        self.hX, self.hY, self.hTheta = wl["X"], Y, Theta
        self.dX = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)           # column-major n x d == C-order d x n
        self.dY = torch.from_numpy(Y).to(dev)
        self.dTh = torch.from_numpy(Theta).to(dev)
        self.dS2 = torch.from_numpy(wl["sigma2"]).to(dev)
        self.dlml = torch.empty(max(self.B, 1), dtype=torch.float64, device=dev)
        self.dinfo = torch.zeros(max(self.B, 1), dtype=torch.int32, device=dev)
        self.ddth = torch.empty(max(self.B, 1) * max(self.p, 1), dtype=torch.float64, device=dev) if grad else None
        self.ddy = torch.empty(max(self.B, 1) * self.n, dtype=torch.float64, device=dev) if grad else None

    def launch(self, stream):
        if self.B == 0:
            return
        self.ctx.lml_batched_dev(self.prog, self.n, self.d, self.dX.data_ptr(), False, self.dY.data_ptr(), self.y_batched,
                                 self.dTh.data_ptr(), self.p, self.dS2.data_ptr(), False, self.wl["jitter"], self.B,
                                 self.dlml.data_ptr(), self.ddth.data_ptr() if self.grad else 0,
                                 self.ddy.data_ptr() if self.grad else 0, self.dinfo.data_ptr(), stream.cuda_stream)


def timed_steps(arm: DevArm, steps: int, warmup: int, flush, world: int, gathered=None, pad_to: int | None = None):
    """W warm-up + K timed steps (CUDA events on the launching stream, L2 flush between steps, max over ranks).
    Returns (ms per step incl. the all-gather, ms per step of the library call alone, launches)."""
    import torch
    import torch.distributed as dist
    stream = torch.cuda.current_stream()
    send = arm.dlml if pad_to is None else torch.zeros(pad_to, dtype=torch.float64, device=arm.dlml.device)

    def gather():
        if world > 1:
            if pad_to is not None:
                send[: arm.B] = arm.dlml[: arm.B]
            dist.all_gather_into_tensor(gathered, send)

    for _ in range(warmup):
        arm.launch(stream)
        gather()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = arm.ctx.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()                                   # L2 flush between timed steps (not timed)
        ev[i][0].record(stream)
        arm.launch(stream)
        ev[i][1].record(stream)
        gather()
        ev[i][2].record(stream)
    torch.cuda.synchronize()
    launches = arm.ctx.launch_count() - l0
    if world > 1:
        dist.barrier()
    step_ms = sum(e[0].elapsed_time(e[2]) for e in ev) / steps
    kern_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=arm.dlml.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = t.tolist()
    return step_ms, kern_ms, int(launches)


def phase_times(arm: DevArm, flush, reps: int):
    """Per-kernel device time (CUDA events around every launch inside the library: option profile_events)."""
    import torch
    stream = torch.cuda.current_stream()
    arm.ctx.set_option("profile_events", 1)
    kms = np.zeros(7)
    kl = np.zeros(7, dtype=np.int64)
    for _ in range(reps):
        flush.zero_()
        arm.launch(stream)
        ms, ln = arm.ctx.last_timing()
        kms += ms
        kl += ln
    arm.ctx.set_option("profile_events", 0)
    torch.cuda.synchronize()
    return kms / reps, kl // reps


def roofline_of(n, B, grad, kms, kl, kern_ms):
    peak, peak_src = fp64_peak()
    flops = B * algorithmic_flops(n, grad)
    step_ach = flops / (kern_ms * 1e-3) * 1e-12
    kf = kernel_flops(n, B, grad)
    dom = int(np.argmax(kms))            # dominant kernel by device time
    dom_name = "lk_%s_kernel" % PHASES[dom]
    launches_dom = max(int(kl[dom]), 1)
    ach = kf[PHASES[dom]] / (kms[dom] * 1e-3) * 1e-12 if kms[dom] > 0 else 0.0
    return {"bound": "tensor", "pipe": "fp64 (DMMA and DFMA share one pipe on sm_100a; tools/pipe_mix.cu)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu_traffic(dom_name),
            "peak_source": peak_src, "kernel": dom_name, "kernel_launches_per_step": launches_dom,
            "kernel_avg_launch_ms": float(kms[dom]) / launches_dom,
            "kernel_share_of_step": float(kms[dom] / max(kms.sum(), 1e-12)),
            "algorithmic_flops_per_launch": kf[PHASES[dom]] / launches_dom,
            "per_kernel_ms_per_step": {nm: float(v) for nm, v in zip(PHASES, kms) if v > 0},
            "per_kernel_tflops": {nm: kf[nm] / (v * 1e-3) * 1e-12 for nm, v in zip(PHASES, kms) if v > 0 and kf[nm] > 0},
            "whole_step": {"achieved": step_ach, "frac": step_ach / peak, "ms": kern_ms, "algorithmic_flops": flops}}


def oracle_check(name: str, world: int, full: np.ndarray, per_rank: int, samples_per_rank: int = 0):
    """Max relative error of sampled items of the gathered vector (every rank's block) against the C oracle."""
    k = samples_per_rank or max(2, -(-8 // world))
    worst, count = 0.0, 0
    for r in range(world):
        wl = workload(name, r)
        idx = np.unique(np.linspace(0, per_rank - 1, k).astype(int))
        ref = cpu_lml(wl, idx, min(host_cores(), len(idx)))
        got = full[r * per_rank + idx]
        worst = max(worst, float(np.max(np.abs(got - ref) / np.abs(ref))))
        count += len(idx)
    return worst, count


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from gaplac_b200 import _lib
    from gaplac_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload(args.workload, rank)
    ctx = _lib.Context(local_rank)
    prog = ctx.program(wl["ops"])
    grad = bool(args.grad)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2

    # ---- device-resident arm (weak scaling: this rank's own batch) -----------------------------------------------------
    arm = DevArm(ctx, prog, wl, dev, grad)
    n, B, p = arm.n, arm.B, arm.p
    gathered = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None
    t0 = time.perf_counter()
    step_ms, kern_ms, launches = timed_steps(arm, args.steps, args.warmup, flush, world, gathered)
    t1 = time.perf_counter()
    bad = int((arm.dinfo != 0).sum().item())
    full = (gathered if world > 1 else arm.dlml).cpu().numpy()
    kms, kl = phase_times(arm, flush, max(2, min(args.steps, 5)))

    # ---- end-to-end arm: host buffers through gpl_lml_batched -----------------------------------------------------------
    hres = None

    def step_e2e():
        nonlocal hres
        hres = ctx.lml_batched(prog, wl["X"], wl["Y"], arm.hTheta, wl["sigma2"], wl["jitter"], grad=grad)
        if world > 1:
            g = torch.from_numpy(hres[0]).to(dev)
            dist.all_gather_into_tensor(gathered, g)
            if rank == 0:
                gathered.cpu()

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
    t2 = time.perf_counter()
    parity = float(np.max(np.abs(hres[0] - arm.dlml.cpu().numpy()) / np.abs(hres[0])))   # the two arms compute the same thing
    h2d = arm.hX.nbytes + arm.hY.nbytes + arm.hTheta.nbytes + wl["sigma2"].nbytes
    d2h = B * 8 + B * 4 + (B * p * 8 + B * n * 8 if grad else 0)

    # ---- strong scaling: the SAME global batch split over the ranks (C2: 4096 proposals; C3: 2000 features) -------------
    strong = {}
    for sname in ("c2", "c3"):
        swl = workload(sname, 0)                         # every rank builds the same global problem, takes its block
        Bg = len(swl["Theta"])
        lo, hi = shard.shard_range(Bg, rank, world)
        sprog = prog if sname == args.workload else ctx.program(swl["ops"])
        sarm = DevArm(ctx, sprog, swl, dev, False, lo, hi)
        m = max(shard.shard_sizes(Bg, world))
        sg = torch.empty(world * m, dtype=torch.float64, device=dev) if world > 1 else None
        s_ms, s_kern, _ = timed_steps(sarm, max(3, min(args.steps, 10)), 3, flush, world, sg, pad_to=m if world > 1 else None)
        if world > 1:
            sizes = shard.shard_sizes(Bg, world)
            sfull = torch.cat([sg[r * m: r * m + sizes[r]] for r in range(world)]).cpu().numpy()
        else:
            sfull = sarm.dlml.cpu().numpy()
        skms, _ = phase_times(sarm, flush, 2)
        entry = {"workload": swl["name"], "global_batch": Bg, "per_rank": hi - lo, "ms_per_step": s_ms,
                 "library_ms": s_kern, "value": Bg / (s_ms * 1e-3), "unit": UNIT,
                 "per_kernel_ms_rank0": {nm: float(v) for nm, v in zip(PHASES, skms) if v > 0}}
        if rank == 0:
            idx = np.unique(np.linspace(0, Bg - 1, 8).astype(int))
            ref = cpu_lml(swl, idx, min(host_cores(), 8))
            entry["oracle_max_rel_err"] = float(np.max(np.abs(sfull[idx] - ref) / np.abs(ref)))
            entry["oracle_items_checked"] = int(len(idx))
        strong[sname] = entry
        del sarm
    t3 = time.perf_counter()

    # ---- one process, all N devices, ONE C-ABI call (gpl_multi_lml_batched): what a Julia host would `ccall` ---------------
    multi = None
    if world > 1:
        # the other ranks must wait on the HOST: an NCCL barrier spins a kernel on their GPUs, which rank 0 is about to use
        host_group = dist.new_group(backend="gloo")
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                mwl = workload("c2", 0)
                mc = _lib.MultiContext(list(range(world)))
                mprog = mc.program(mwl["ops"])
                for _ in range(2):
                    mres = mc.lml_batched(mprog, mwl["X"], mwl["Y"], mwl["Theta"], mwl["sigma2"], mwl["jitter"])
                m0 = time.perf_counter()
                for _ in range(5):
                    mres = mc.lml_batched(mprog, mwl["X"], mwl["Y"], mwl["Theta"], mwl["sigma2"], mwl["jitter"])
                m_ms = (time.perf_counter() - m0) * 1e3 / 5
                idx = np.unique(np.linspace(0, len(mwl["Theta"]) - 1, 8).astype(int))
                ref = cpu_lml(mwl, idx, min(host_cores(), 8))
                multi = {"api": "gpl_multi_lml_batched: one host call, the 4096 proposals split over the devices, every device "
                                "writes its slice into the caller's buffer (no collective)", "devices": world,
                         "global_batch": len(mwl["Theta"]), "ms_per_call": m_ms, "value": len(mwl["Theta"]) / (m_ms * 1e-3),
                         "unit": UNIT, "timing": "wall clock of 5 blocking calls (host buffers, copies inside)",
                         "oracle_max_rel_err": float(np.max(np.abs(mres[0][idx] - ref) / np.abs(ref)))}
                mc.close()
            except Exception as e:          # never lose the headline line to the extra measurement
                multi = {"error": repr(e)}
        dist.barrier(group=host_group)

    out = None
    if rank == 0:
        err, cnt = oracle_check(args.workload, world, full, B)
        out = {
            "metric": METRIC, "value": world * B / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, grad, world),
            "notes": {"sharding": f"independent proposals, {B} per rank, one NCCL all-gather of lml per step",
                      "l2": "256 MiB buffer written between timed steps (L2 flush, untimed); per-item factor workspace "
                            "also exceeds L2", "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "roofline": roofline_of(n, B, grad, kms, kl, kern_ms),
            "e2e": {"value": world * B / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e_ms,
                    "api": "gpl_lml_batched (caller-owned host buffers, blocking) via ctypes"},
            "gpu_launches": int(launches), "clocks": sampler.window(t0, t1), "clocks_e2e": sampler.window(t1, t2),
            "not_pd_items": bad, "arms_max_rel_diff": parity,
            "oracle_max_rel_err": err, "oracle_items_checked": cnt,
            "strong": dict(strong, clocks=sampler.window(t2, t3),
                           note="same global batch split over the ranks (contiguous blocks, padded all-gather); value = global batch / max-over-ranks step time"),
        }
        if multi is not None:
            out["multi_abi"] = multi
    if world == 1 and not args.no_configs:
        from bench_configs import run_configs
        tc0 = time.perf_counter()
        cfg = run_configs(ctx, dev, flush, quick=args.quick_configs)
        cfg["clocks"] = sampler.window(tc0, time.perf_counter())
        out["configs"] = cfg
    if rank == 0:
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(wl)
        sampler.stop()
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
    ap.add_argument("--grad", action="store_true", help="also return the analytic gradient (the mcmc inner loop)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (other BASELINE configs, N = 1)")
    ap.add_argument("--quick-configs", action="store_true", help="configs block without the slow oracle checks (C5)")
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
