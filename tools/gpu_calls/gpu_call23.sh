#!/bin/bash
for n in 2048 8192; do GAPLAC_B200_LIB=$PWD/gaplac_b200/libgaplac_b200_bigprof.so timeout 120 python tools/run_c5.py $n 2>&1 | tail -4; done
