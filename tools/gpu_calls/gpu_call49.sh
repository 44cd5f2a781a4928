#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_int8_trail.py tests/test_ozaki.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/t49.log 2>&1; tail -3 gpurun_out/t49.log
