#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t12.log 2>&1; tail -4 gpurun_out/t12.log
python tools/small_n_ab.py > gpurun_out/small_n_ab.json 2> gpurun_out/small_n_ab.err; cat gpurun_out/small_n_ab.json | tr -d '\n' | cut -c1-3000; tail -c 300 gpurun_out/small_n_ab.err
python tools/sweep_large.py > gpurun_out/sweep_large_r02.txt 2> gpurun_out/sweep_large.err; head -4 gpurun_out/sweep_large_r02.txt
python bench.py --steps 10 > gpurun_out/b12.json 2> gpurun_out/b12.err; tail -c 300 gpurun_out/b12.err
