#!/bin/bash
# Ozaki INT8 prototype v5: full record (exactness, speed, lml) + ncu of the update at n=8192 K=512 S=8
timeout 500 python tools/ozaki/ozaki_bench.py --json gpurun_out/oz34.json > gpurun_out/oz34.log 2>&1; echo rc=$?; grep "^lml\|S=" gpurun_out/oz34.log | tail -8
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ozaki_syrk -s 1 -c 1 -o gpurun_out/oz34 -f python tools/ozaki/ozaki_one.py 8192 512 8 3 > gpurun_out/oz34_ncu.log 2>&1; echo rc=$?
