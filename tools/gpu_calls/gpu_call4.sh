#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t4.log 2>&1; tail -4 gpurun_out/t4.log
python tools/cov_bench.py > gpurun_out/cov_bench_r02b.json 2> gpurun_out/cov_bench.err; tail -c 300 gpurun_out/cov_bench.err
python tools/sanitize.py memcheck > gpurun_out/sanitize_memcheck_stdout.txt 2>&1; tail -15 gpurun_out/sanitize_memcheck_stdout.txt
