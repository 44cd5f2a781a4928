#!/bin/bash
# Ozaki INT8 prototype v5 (C prefetch, shuffled scales, early slot release): exactness + speed
timeout 500 python tools/ozaki/ozaki_bench.py --skip lml --json gpurun_out/oz33.json > gpurun_out/oz33.log 2>&1; echo rc=$?; grep -v "^fp64\|^update" gpurun_out/oz33.log | tail -40
