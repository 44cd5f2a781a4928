#!/bin/bash
# ncu of the Ozaki update kernel (v3) at n=8192 K=512 S=8, source-level samples
timeout 120 python tools/ozaki/ozaki_one.py 8192 512 8 3 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ozaki_syrk -s 1 -c 1 -o gpurun_out/oz30 -f python tools/ozaki/ozaki_one.py 8192 512 8 3 > gpurun_out/oz30_ncu.log 2>&1; echo rc=$?; tail -3 gpurun_out/oz30_ncu.log
