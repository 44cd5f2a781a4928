#!/usr/bin/env python3
"""Summarise an ncu per-launch CSV of the lockstep kernels (see profiles/README.md)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ik, im, iv, iid, ig = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID'), H.index('Grid Size')
d = OrderedDict()
for r in data:
    if len(r) <= iv: continue
    d.setdefault(r[iid], {'k': r[ik].split('(')[0].replace('gpl::', ''), 'g': r[ig]})[r[im]] = float(r[iv].replace(',', ''))
tot = 0; per = {}
for k, v in d.items():
    t = v['gpu__time_duration.sum'] / 1e3; tot += t; per[v['k']] = per.get(v['k'], 0) + t
    print(f"{k:>4} {v['k']:18s} grid {v['g']:>16s} {t:8.1f} us  dmma {v.get('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  fp64 {v.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  dram R {v['dram__bytes_read.sum'] / 1e6:7.1f} MB W {v['dram__bytes_write.sum'] / 1e6:7.1f} MB")
print('sum us', round(tot, 1), {k: round(v, 1) for k, v in per.items()})
