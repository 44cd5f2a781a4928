#!/bin/bash
python bench.py --no-cpu --no-configs --steps 2 > gpurun_out/b10.json 2> gpurun_out/b10.err && \
ncu --set full --clock-control none --import-source on -k regex:lk_potrf_warp -s 26 -c 2 -o gpurun_out/potrf_r02 python bench.py --no-cpu --no-configs --steps 1 > gpurun_out/ncu_potrf.log 2>&1
tail -2 gpurun_out/ncu_potrf.log
