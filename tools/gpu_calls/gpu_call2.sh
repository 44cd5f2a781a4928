#!/bin/bash
# sampler bring-up: device chains vs the host reference
python -m pytest tests/test_gpu_mcmc.py -x -q -s > gpurun_out/t2.log 2>&1; tail -40 gpurun_out/t2.log
