#!/bin/bash
# Ozaki INT8 prototype, first contact: exactness on small shapes only (bounded waits, no traps)
timeout 240 python tools/ozaki/ozaki_bench.py --skip speed,lml > gpurun_out/oz25.log 2>&1; echo rc=$?; tail -40 gpurun_out/oz25.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
