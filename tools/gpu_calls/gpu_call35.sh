#!/bin/bash
# Ozaki INT8 prototype v6 (N=256 fused chains): GPU tests + speed
timeout 300 python -m pytest tests/test_ozaki.py -x -q -m gpu > gpurun_out/t35.log 2>&1; tail -3 gpurun_out/t35.log
timeout 500 python tools/ozaki/ozaki_bench.py --json gpurun_out/oz35.json > gpurun_out/oz35.log 2>&1; echo rc=$?; grep "lml\|S=8" gpurun_out/oz35.log | tail
