#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t11.log 2>&1; tail -4 gpurun_out/t11.log
python tools/sweep_large.py > gpurun_out/sweep_large_r02.txt 2> gpurun_out/sweep_large.err; cat gpurun_out/sweep_large_r02.txt; tail -c 300 gpurun_out/sweep_large.err
python bench.py --grad --no-cpu --no-configs --steps 5 > gpurun_out/b11g.json 2> gpurun_out/b11g.err; tail -c 200 gpurun_out/b11g.err
