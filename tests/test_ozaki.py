"""tools/ozaki: the Ozaki-split INT8 rank-K update on tcgen05 (microbenchmark).

CPU: the product schedule the host hands to the kernel (every slice pair once, group flags, TMEM slot reuse).
GPU: the kernel against an integer emulation (bit for bit) and against FP64, and its effect on a blocked-Cholesky lml.
"""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "ozaki"))


@pytest.fixture(scope="module")
def lib():
    from tools.ozaki.build import build
    return ctypes.CDLL(build())


def schedule(lib, S):
    buf = ctypes.create_string_buffer(1 << 14)
    lib.ozaki_schedule_text(S, buf, len(buf))
    lines = buf.value.decode().splitlines()
    order = [int(v) for v in lines[0].split("order=")[1].split()]
    steps = []
    for ln in lines[1:]:
        rows, cols, prods, issue = ln.split(":", 1)[1].split("|")
        steps.append(dict(rows=[int(v) for v in rows.split()[1:]], cols=[int(v) for v in cols.split()[1:]],
                          prods=[(int(a), int(b), int(g), f) for a, b, g, f in re.findall(r"\((\d+),(\d+)\)->g(\d+)([FL]*)", prods)],
                          issues=[(int(a), int(b), w == "w", int(slot)) for a, b, w, slot in re.findall(r"(\d+):(\d+)(w?)@(\d+)", issue)]))
    return order, steps


@pytest.mark.parametrize("S", range(1, 10))
def test_schedule_covers_every_slice_pair_once(lib, S):
    order, steps = schedule(lib, S)
    seen = {}
    for st in steps:
        assert 1 <= len(st["rows"]) <= 2 and 1 <= len(st["cols"]) <= 2 and 1 <= len(st["prods"]) <= 4
        for a, b, g, _ in st["prods"]:
            assert a in st["rows"] and b in st["cols"], "a product may only use slices its stage loads"
            assert g == a + b and (a, b) not in seen
            seen[(a, b)] = g
    assert set(seen) == {(a, b) for a in range(S) for b in range(S) if a + b <= S - 1}
    assert sorted(order) == list(range(S))


@pytest.mark.parametrize("S", range(1, 10))
def test_schedule_group_flags_and_tmem_slots(lib, S):
    """First / last flags bracket each group, the epilogue order is the commit order, and a TMEM slot (4 of them, slot =
    g mod 4) is never handed to a new group before the group that held it has been committed."""
    order, steps = schedule(lib, S)
    first, last, committed = {}, {}, []
    for i, st in enumerate(steps):
        committed += sorted((g for _, _, g, f in st["prods"] if "L" in f), key=lambda g: g % 4)  # commits go by ascending slot
        for a, b, g, f in st["prods"]:
            if "F" in f:
                assert g not in first
                first[g] = i
            assert g in first, "accumulation before the overwrite"
            if "L" in f:
                assert g not in last
                last[g] = i
            else:
                assert g not in last or last[g] >= i
    assert committed == order and set(first) == set(last) == set(range(S))
    for g in range(S):
        for other in range(S):
            if other != g and other % 4 == g % 4 and first[other] > first[g]:
                assert first[other] > last[g], f"groups {g} and {other} would share a slot while both are live"


def _torch():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test without a CUDA device")
    return torch


@pytest.mark.gpu
def test_int8_products_match_the_integer_emulation_bit_for_bit(lib):
    torch = _torch()
    import ozaki_bench as ob
    torch.manual_seed(3)
    for (n, K) in [(128, 128), (384, 256), (640, 384)]:
        A = torch.randn(n, K, dtype=torch.float64, device="cuda") * torch.exp2(torch.randint(-8, 9, (n, 1), device="cuda").double())
        mask = ob.lower_tiles_mask(n)
        for S in (1, 2, 3):
            rs, qs = ob.split_emulated(A, S)
            want = torch.zeros(n, n, dtype=torch.float64, device="cuda")
            for g in range(S):
                want += sum(qs[s] @ qs[g - s].T for s in range(g + 1)) * 2.0 ** (-7 * (g + 2))
            want = want * rs[:, None] * rs[None, :]
            C = ob.colmajor(n, n, 7.0)
            ob.ozaki(A, C, S, 1)
            ob.sync("exact")
            assert torch.equal((C * mask), (want * mask)), (n, K, S)


@pytest.mark.gpu
def test_eight_slices_reach_fp64_and_update_in_place(lib):
    torch = _torch()
    import ozaki_bench as ob
    torch.manual_seed(4)
    n, K = 512, 512
    A = torch.randn(n, K, dtype=torch.float64, device="cuda")
    A[::7] *= 1e-3  # rows of very different magnitude: the split scales each row by its own maximum
    ref = A @ A.T
    nrm = A.norm(dim=1)
    mask = ob.lower_tiles_mask(n)
    for S, tol in ((8, 1e-14), (7, 1e-12), (6, 1e-10)):
        C0 = torch.randn(n, n, dtype=torch.float64, device="cuda")
        C = ob.colmajor(n, n)
        C.copy_(C0)
        ob.ozaki(A, C, S, 0)
        ob.sync("update")
        # error of the product relative to the row norms; the subtraction from C0 adds its own FP64 rounding of |C0|
        err = (((C - (C0 - ref)) * mask).abs() / (nrm[:, None] * nrm[None, :] + C0.abs())).max().item()
        assert err < tol, (S, err)


@pytest.mark.gpu
def test_blocked_cholesky_lml_with_int8_trailing_updates(lib):
    """The quantity north_star bounds: lml relative error (1e-9) when every trailing update of a blocked Cholesky goes
    through the INT8 split.  8 slices are indistinguishable from FP64; 6 still meet the bound on this kernel matrix."""
    torch = _torch()
    import math

    import ozaki_bench as ob
    n = 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 40 - 20
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    d = x[:, None] - x[None, :]
    Kmat = torch.exp(-d * d / 2.0) + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(Kmat)
    z = torch.linalg.solve_triangular(L, y[:, None], upper=False)[:, 0]
    ref = -0.5 * ((z * z).sum().item() + 2.0 * torch.log(torch.diagonal(L)).sum().item() + n * math.log(2 * math.pi))
    for S, tol in ((8, 1e-13), (6, 1e-9)):
        got = ob.chol_blocked(Kmat, y, 256, S)[0]
        ob.sync("lml")
        assert abs(got - ref) / abs(ref) < tol, (S, got, ref)


@pytest.mark.parametrize("S", range(1, 10))
def test_schedule_issue_list_matches_the_products(lib, S):
    """The chains actually issued (some fused to N = 256 over two adjacent TMEM slots) cover the step's products exactly."""
    _, steps = schedule(lib, S)
    for st in steps:
        covered = []
        for a, b, wide, slot in st["issues"]:
            assert slot == (a + b) % 4
            covered.append((a, b))
            if wide:
                assert slot != 3 and b + 1 in st["cols"] and b == st["cols"][0]
                covered.append((a, b + 1))
        assert sorted(covered) == sorted((a, b) for a, b, _, _ in st["prods"])
        # a fused chain overwrites or accumulates both halves alike
        flags = {(a, b): f for a, b, _, f in st["prods"]}
        for a, b, wide, _ in st["issues"]:
            if wide:
                assert ("F" in flags[(a, b)]) == ("F" in flags[(a, b + 1)])
