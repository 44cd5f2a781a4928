"""Ozaki-split INT8 rank-K update on tcgen05 (tools/ozaki/ozaki_syrk.cu): exactness, speed, and the effect on the C5 lml.

    python tools/ozaki/ozaki_bench.py [--quick] [--json out.json]

The kernel is the one the library's large-n factorisation uses (gaplac_b200/csrc/int8_syrk.cuh), here on dense matrices.  torch
supplies the FP64 yardsticks (cuBLAS DGEMM for the same update, cuSOLVER for the panel factorisations of the blocked-Cholesky
harness); this harness itself is not on the product path.
"""
import argparse
import ctypes
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libozaki.so"))
lib.ozaki_ws_bytes.restype = ctypes.c_long
lib.ozaki_ws_bytes.argtypes = [ctypes.c_int] * 3
lib.ozaki_split.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
lib.ozaki_update.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                             ctypes.c_void_p]
lib.ozaki_last_debug.argtypes = [ctypes.c_void_p]


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} returned {rc}")


def sync(what):
    """Synchronise and turn a timed-out barrier wait (the kernel leaves early with a breadcrumb) into an exception."""
    torch.cuda.synchronize()
    dbg = (ctypes.c_int * 4)()
    lib.ozaki_last_debug(dbg)
    if dbg[0] != 0:
        raise RuntimeError(f"{what}: barrier wait timed out, (code, cta, parity) = {list(dbg)[:3]}")


def ozaki(A, C, S, mode, ws=None):
    """C (column-major view: unit row stride) -= / = A A^T on lower-triangle 128x128 tiles; A row-major."""
    n, K = A.shape
    assert A.stride(1) == 1 and C.stride(0) == 1
    if ws is None:
        ws = torch.empty(lib.ozaki_ws_bytes(n, K, S), dtype=torch.uint8, device="cuda")
    check(lib.ozaki_split(A.data_ptr(), A.stride(0), n, K, S, ws.data_ptr(), stream()), "split")
    check(lib.ozaki_update(n, K, S, ws.data_ptr(), C.data_ptr(), C.stride(1), mode, stream()), "update")
    return ws


def split_emulated(A, S):
    """The same truncation slices in torch FP64 (every step is exact)."""
    mx = A.abs().amax(dim=1)
    e = torch.where(mx > 0, torch.floor(torch.log2(mx)) + 1, torch.zeros_like(mx))
    e = torch.where(mx >= torch.exp2(e), e + 1, e)  # guard log2 rounding at powers of two
    e = torch.where(mx < torch.exp2(e - 1), e - 1, e)
    r = A * torch.exp2(-e)[:, None]
    qs = []
    for _ in range(S):
        r = r * 128.0
        q = torch.trunc(r)
        r = r - q
        qs.append(q)
    return torch.exp2(e), qs


def colmajor(n, m, fill=None):
    """An n x m FP64 matrix stored column-major (what LAPACK and the product's tiles use)."""
    t = torch.empty(m, n, dtype=torch.float64, device="cuda").t()
    if fill is not None:
        t.fill_(fill)
    return t


def lower_tiles_mask(n):
    t = torch.arange(n, device="cuda") // 128
    return t[:, None] >= t[None, :]


def time_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def exactness(out):
    torch.manual_seed(1)
    res = []
    for (n, K) in [(128, 128), (256, 128), (384, 256), (1024, 512)]:
        A = torch.randn(n, K, dtype=torch.float64, device="cuda") * torch.exp2(torch.randint(-6, 7, (n, 1), device="cuda").double())
        mask = lower_tiles_mask(n)
        # S = 1 and S = 2: integer arithmetic end to end, so the kernel must reproduce the emulation bit for bit
        for S in (1, 2, 3):
            rs, qs = split_emulated(A, S)
            want = torch.zeros(n, n, dtype=torch.float64, device="cuda")
            for g in range(S):
                grp = sum(qs[s] @ qs[g - s].T for s in range(g + 1))
                want += grp * 2.0 ** (-7 * (g + 2))
            want = want * rs[:, None] * rs[None, :]
            C = colmajor(n, n, 7.0)
            ozaki(A, C, S, 1)
            sync(f"exact n={n} K={K} S={S}")
            diff = ((C - want) * mask).abs().max().item()
            res.append(dict(n=n, K=K, S=S, max_abs_diff_vs_emulation=diff, scale=want.abs().max().item()))
            print(f"exact  n={n:5d} K={K:4d} S={S}  max |kernel - emulation| = {diff:.3e}  (|C| max {want.abs().max().item():.3e})", flush=True)
        ref = A @ A.T
        for S in (5, 6, 7, 8):
            C = colmajor(n, n, 0.0)
            ozaki(A, C, S, 1)
            sync(f"fp64 n={n} K={K} S={S}")
            nrm = A.norm(dim=1)
            err = (((C - ref) * mask).abs() / (nrm[:, None] * nrm[None, :])).max().item()
            res.append(dict(n=n, K=K, S=S, max_err_over_row_norms=err))
            print(f"fp64   n={n:5d} K={K:4d} S={S}  max |C - A A^T| / (|a_i| |a_j|) = {err:.3e}", flush=True)
        # mode 0 subtracts in place
        C0 = torch.randn(n, n, dtype=torch.float64, device="cuda")
        C = colmajor(n, n)
        C.copy_(C0)
        ozaki(A, C, 8, 0)
        sync("mode 0")
        err = ((C - (C0 - ref)) * mask).abs().max().item() / ref.abs().max().item()
        print(f"update n={n:5d} K={K:4d} S=8  max |(C0 - A A^T) - C| / max|A A^T| = {err:.3e}", flush=True)
        res.append(dict(n=n, K=K, S=8, update_rel_err=err))
    out["exactness"] = res


def speed(out, quick):
    res = []
    shapes = [(8192, 512)] if quick else [(8192, 512), (8192, 1024), (4096, 512), (8192, 256)]
    for (n, K) in shapes:
        A = torch.randn(n, K, dtype=torch.float64, device="cuda")
        C = colmajor(n, n, 0.0)
        flops = (n // 128) * (n // 128 + 1) / 2 * 2 * 128 * 128 * K  # the lower-triangle tiles actually computed
        t_gemm = time_ms(lambda: torch.addmm(C, A, A.T, beta=1.0, alpha=-1.0, out=C))
        row = dict(n=n, K=K, dgemm_full_ms=t_gemm, dgemm_tf=2.0 * n * n * K / t_gemm * 1e-9)
        print(f"speed  n={n} K={K}: cuBLAS DGEMM (full square, 2x the flops) {t_gemm:.3f} ms = {row['dgemm_tf']:.1f} TF", flush=True)
        for S in (8, 7, 6, 5):
            ws = torch.empty(lib.ozaki_ws_bytes(n, K, S), dtype=torch.uint8, device="cuda")
            t_split = time_ms(lambda: check(lib.ozaki_split(A.data_ptr(), A.stride(0), n, K, S, ws.data_ptr(), stream()), "split"))
            t_upd = time_ms(lambda: check(lib.ozaki_update(n, K, S, ws.data_ptr(), C.data_ptr(), C.stride(1), 0, stream()), "update"))
            sync("speed")
            pairs = S * (S + 1) // 2
            row[f"S{S}"] = dict(split_ms=t_split, update_ms=t_upd, fp64_equiv_tf=flops / (t_split + t_upd) * 1e-9,
                                int8_pops=pairs * flops / t_upd * 1e-12)
            print(f"       S={S}: split {t_split:.3f} ms + update {t_upd:.3f} ms  = {row[f'S{S}']['fp64_equiv_tf']:.1f} FP64-equivalent TF"
                  f"   ({pairs} INT8 products at {row[f'S{S}']['int8_pops']:.2f} POP/s)", flush=True)
        res.append(row)
    out["speed"] = res


def chol_blocked(Kmat, y, nb, S):
    """Right-looking blocked Cholesky; S = 0 uses cuBLAS FP64 for the trailing update, S > 0 the INT8 split."""
    n = Kmat.shape[0]
    A = colmajor(n, n)
    A.copy_(Kmat)
    ws = None
    for p in range(0, n, nb):
        e = min(p + nb, n)
        L11 = torch.linalg.cholesky(A[p:e, p:e])
        A[p:e, p:e] = L11
        if e < n:
            L21 = torch.linalg.solve_triangular(L11, A[e:, p:e].T, upper=False).T.contiguous()
            A[e:, p:e] = L21
            T = A[e:, e:]
            if S == 0:
                T -= L21 @ L21.T
            else:
                if ws is None:
                    ws = torch.empty(lib.ozaki_ws_bytes(n, nb, S), dtype=torch.uint8, device="cuda")
                ozaki(L21, T, S, 0, ws)
    L = torch.tril(A)
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum().item()
    z = torch.linalg.solve_triangular(L, y[:, None], upper=False)[:, 0]
    quad = (z * z).sum().item()
    return -0.5 * (quad + logdet + n * math.log(2 * math.pi)), logdet, quad


def lml_effect(out, quick):
    n = 4096 if quick else 8192
    rng = np.random.default_rng(5)  # the C5 workload of gaplac_b200/workloads.py: SqExp(l = 1) + 0.1 I on U(-50, 50)
    x = torch.tensor(rng.uniform(-50, 50, n), device="cuda")
    y = torch.tensor(rng.standard_normal(n), device="cuda")
    d = x[:, None] - x[None, :]
    Kmat = torch.exp(-d * d / 2.0) + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
    ref = chol_blocked(Kmat, y, 512, 0)
    L = torch.linalg.cholesky(Kmat)
    z = torch.linalg.solve_triangular(L, y[:, None], upper=False)[:, 0]
    lml_lapack = -0.5 * ((z * z).sum().item() + 2.0 * torch.log(torch.diagonal(L)).sum().item() + n * math.log(2 * math.pi))
    res = dict(n=n, lml_cusolver=lml_lapack, lml_blocked_fp64=ref[0], rel_blocked_vs_cusolver=abs(ref[0] - lml_lapack) / abs(lml_lapack))
    print(f"lml    n={n}: cuSOLVER {lml_lapack:.10f}; blocked FP64 harness {ref[0]:.10f} (rel {res['rel_blocked_vs_cusolver']:.2e})", flush=True)
    for S in (8, 7, 6, 5, 4):
        got = chol_blocked(Kmat, y, 512, S)
        sync(f"lml S={S}")
        res[f"S{S}"] = dict(lml=got[0], rel_err_lml=abs(got[0] - lml_lapack) / abs(lml_lapack),
                            rel_err_logdet=abs(got[1] - ref[1]) / abs(ref[1]), rel_err_quad=abs(got[2] - ref[2]) / abs(ref[2]))
        print(f"       S={S}: lml {got[0]:.10f}  rel err lml {res[f'S{S}']['rel_err_lml']:.2e}  logdet {res[f'S{S}']['rel_err_logdet']:.2e}"
              f"  quad {res[f'S{S}']['rel_err_quad']:.2e}", flush=True)
    out["lml"] = res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default="")
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    out = dict(device=torch.cuda.get_device_name(0))
    if "exact" not in a.skip:
        exactness(out)
    if "speed" not in a.skip:
        speed(out, a.quick)
    if "lml" not in a.skip:
        lml_effect(out, a.quick)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)
    print("done", flush=True)
