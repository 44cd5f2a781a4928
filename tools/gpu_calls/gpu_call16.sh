#!/bin/bash
python bench.py --no-cpu --no-configs --steps 2 > gpurun_out/b16.json 2> gpurun_out/b16.err && \
ncu --set full --clock-control none --import-source on -k regex:lk_below -s 21 -c 4 -o gpurun_out/below_r02 python bench.py --no-cpu --no-configs --steps 1 > gpurun_out/ncu_below.log 2>&1
tail -2 gpurun_out/ncu_below.log
