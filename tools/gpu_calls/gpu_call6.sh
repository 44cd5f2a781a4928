#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t6.log 2>&1; tail -4 gpurun_out/t6.log
python bench.py --steps 10 > gpurun_out/b6.json 2> gpurun_out/b6.err; tail -c 400 gpurun_out/b6.err
python bench.py --grad --no-cpu --no-configs --steps 5 > gpurun_out/b6g.json 2> gpurun_out/b6g.err; tail -c 300 gpurun_out/b6g.err
