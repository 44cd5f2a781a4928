#!/usr/bin/env python3
"""Per-launch opcode / source-line histogram of an `ncu --page source --csv` dump (several launches in one file).

    python tools/ncu_ops.py src.csv X.dis <kernel-substring> <launch-index> [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern, which = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
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
# split into launches at header rows
launches, hdrs, curl = [], [], None
for r in rows:
    if r and r[0] == "Address":
        hdrs.append(r)
        curl = []
        launches.append(curl)
    elif curl is not None and r and re.match(r"^[0-9a-fx]+$", r[0]):
        curl.append(r)
L = launches[which]
hdr = hdrs[which]
isrc, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(L[0][0], 16)
ops = defaultdict(int)
lines = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int)])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
allst = defaultdict(int)
tot = tots = 0
for r in L:
    e, s = int(r[iex] or 0), int(r[isamp] or 0)
    txt = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
    op = (txt.split()[0] if txt else "?").split(".")[0]
    ops[op] += e
    key = addr2line.get(int(r[0], 16) - base, ("?", 0))
    lines[key][0] += e
    lines[key][1] += s
    lines[key][2][op] += e
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            lines[key][3][hdr[i][6:]] += v
            allst[hdr[i][6:]] += v
    tot += e
    tots += s
print(f"launch {which}: {tot} warp instructions, {tots} samples")
print("  ".join(f"{k}={100 * v / tot:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
print("stalls: " + "  ".join(f"{k}={100 * v / max(tots, 1):.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:10]))
bysamp = "--by-samples" in sys.argv
for key, (e, s, o, st) in sorted(lines.items(), key=lambda kv: -(kv[1][1] if bysamp else kv[1][0]))[:top]:
    mix = " ".join(f"{k}:{v * 100 // max(e, 1)}" for k, v in sorted(o.items(), key=lambda kv: -kv[1])[:5])
    sts = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * e / tot:6.2f}% inst {100 * s / max(tots, 1):6.2f}% samp  {key[0]}:{key[1]:<5d} {mix} | {sts}")
