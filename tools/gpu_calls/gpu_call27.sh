#!/bin/bash
# Ozaki INT8 prototype v2 (2x2 slice blocks, four TMEM accumulators): exactness, speed, lml
timeout 500 python tools/ozaki/ozaki_bench.py --json gpurun_out/oz27.json > gpurun_out/oz27.log 2>&1; echo rc=$?; grep -v "^fp64\|^update" gpurun_out/oz27.log | tail -60
