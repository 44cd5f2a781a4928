#!/bin/bash
# first contact of option trail_int8 in gpl_lml_large
timeout 600 python tools/i8_large.py 4096 8192 16384 > gpurun_out/i8_37.log 2>&1; echo rc=$?; tail -16 gpurun_out/i8_37.log
