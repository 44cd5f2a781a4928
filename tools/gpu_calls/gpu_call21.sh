#!/bin/bash
python -m pytest tests/test_gpu_fuzz.py -x -q > gpurun_out/t21.log 2>&1; tail -5 gpurun_out/t21.log
python -m pytest tests -m gpu -x -q > gpurun_out/t21b.log 2>&1; tail -3 gpurun_out/t21b.log
