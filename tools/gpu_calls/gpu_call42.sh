#!/bin/bash
# tests of the INT8 trailing path + the large-n / hygiene / multi tests that go through big_factor
timeout 900 python -m pytest tests/test_gpu_int8_trail.py tests/test_gpu_fullsize.py tests/test_gpu_hygiene.py tests/test_gpu_multi.py tests/test_ozaki.py -x -q -m gpu > gpurun_out/t42.log 2>&1; tail -25 gpurun_out/t42.log
