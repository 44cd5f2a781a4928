#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/t64.log 2>&1; tail -4 gpurun_out/t64.log
