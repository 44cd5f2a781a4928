#!/usr/bin/env python3
"""Opcode histogram per kernel of libgaplac_b200.so (cuobjdump -sass): the evidence for which hardware paths the kernels
use - DMMA.8x8x4 (FP64 tensor-core path; tcgen05 has no f64 kind), LDGSTS (cp.async), UBLKCP + SYNCS (1-D TMA bulk copies
on mbarriers), UTCIMMA / UTCBAR / LDTM / UTMALDG (tcgen05.mma.kind::i8, tcgen05.commit, tcgen05.ld, tensor-map TMA loads of the
INT8 trailing update), plus registers / stack / shared memory per kernel (cuobjdump -res-usage).

    python tools/sass_histogram.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaplac_b200", "libgaplac_b200.so")
KEY = ["DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MUFU", "ATOMS",
       "ATOMG", "RED", "UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    hist = collections.OrderedDict()
    cur = None
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", ln)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
            if m.group(1).startswith("DMMA"):
                hist[cur]["shape:" + m.group(1)] += 1
    print(f"{os.path.relpath(LIB, ROOT)}: cubin architectures {sorted(archs)}")
    print(f"{'kernel':44s} {'instr':>6s} {'REG':>4s} {'STACK':>5s} {'SMEM':>6s}  " + " ".join(f"{k:>6s}" for k in KEY))
    tot = collections.Counter()
    for k, h in hist.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0].replace("gpl::", "")
        u = usage.get(k, (0, 0, 0))
        n = sum(v for kk, v in h.items() if not kk.startswith("shape:"))
        print(f"{name[:44]:44s} {n:6d} {u[0]:4d} {u[1]:5d} {u[2]:6d}  " + " ".join(f"{h.get(x, 0):6d}" for x in KEY))
        tot.update(h)
    print(f"{'TOTAL':44s} {'':6s} {'':4s} {'':5s} {'':6s}  " + " ".join(f"{tot.get(x, 0):6d}" for x in KEY))
    print("DMMA shapes:", {k[6:]: v for k, v in tot.items() if k.startswith("shape:")})
    print("tcgen05 / TMEM / tensor-map TMA opcodes (UTCMMA, LDTM, UTMALDG): none expected - FP64 has no tcgen05 kind (DESIGN.md section 5)")


if __name__ == "__main__":
    main()
