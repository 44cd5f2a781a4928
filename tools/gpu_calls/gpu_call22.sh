#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_hygiene.py tests/test_gpu_multi.py -x -q > gpurun_out/t22.log 2>&1; tail -5 gpurun_out/t22.log
timeout 300 python tools/sweep_large.py > gpurun_out/sweep_large_v2.txt 2> gpurun_out/sweep_large.err; cat gpurun_out/sweep_large_v2.txt | head -12; tail -c 300 gpurun_out/sweep_large.err
timeout 120 python tools/run_c5.py 8192 2>&1 | tail -3
