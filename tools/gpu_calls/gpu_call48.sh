#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_int8_trail.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/t48.log 2>&1; tail -5 gpurun_out/t48.log
