#!/bin/bash
# Ozaki INT8 prototype v1: speed against cuBLAS DGEMM and the effect on the C5 lml
timeout 500 python tools/ozaki/ozaki_bench.py --skip exact --json gpurun_out/oz26.json > gpurun_out/oz26.log 2>&1; echo rc=$?; tail -40 gpurun_out/oz26.log
