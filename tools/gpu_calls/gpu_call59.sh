#!/bin/bash
timeout 600 python tools/i8_large.py 5120 6144 7168 > gpurun_out/i8_59.log 2>&1; grep "trail_int8=8\|trail_int8=0" gpurun_out/i8_59.log
