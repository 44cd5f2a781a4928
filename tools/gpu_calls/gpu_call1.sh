#!/bin/bash
# first GPU pass of round 2: parity tests, headline bench, gradient bench + launch list
python -m pytest tests -m gpu -x -q > gpurun_out/t1.log 2>&1; tail -5 gpurun_out/t1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/b1.json 2> gpurun_out/b1.err; tail -c 600 gpurun_out/b1.err
python bench.py --grad --no-cpu --no-configs --steps 5 > gpurun_out/b1g.json 2> gpurun_out/b1g.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_grad.csv \
    python bench.py --grad --no-cpu --no-configs --steps 1 --warmup 3 > gpurun_out/ncu1.log 2>&1
tail -c 400 gpurun_out/b1g.err
