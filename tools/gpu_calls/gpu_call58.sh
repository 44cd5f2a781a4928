#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_int8_trail.py -x -q -m gpu > gpurun_out/t58.log 2>&1; tail -2 gpurun_out/t58.log
timeout 600 python tools/i8_large.py 4096 8192 16384 > gpurun_out/i8_58.log 2>&1; grep "trail_int8=8\|trail_int8=0" gpurun_out/i8_58.log
