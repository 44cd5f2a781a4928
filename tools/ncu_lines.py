#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all gaplac_b200/libgaplac_b200.so ; nvdisasm --print-line-info X.cubin > X.dis
    python tools/ncu_lines.py src.csv X.dis <kernel-substring> [top]

Prints the share of warp-stall samples (and executed instructions) per file:line, i.e. where the kernel's
time goes, using the -lineinfo tables in the cubin.
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# address -> (file, line) from nvdisasm: lines like  //## File "x.cu", line 123   then   /*0040*/ INSTR
addr2line = {}
cur = None
in_k = False
for ln in open(dis, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if ".text." in ln and ln.strip().startswith(".section"):
        in_k = kern in ln
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m and in_k and cur:
        addr2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    s = int(r[isamp] or 0)
    agg[key][0] += s
    agg[key][1] += int(r[iexec] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            agg[key][2][hdr[i]] += v
    tot += s
print(f"total samples {tot}")
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * s / max(tot, 1):6.2f}%  {key[0]}:{key[1]:<5d} inst={e:<12d} {tops}")
