#!/bin/bash
# ncu evidence: covariance build kernels (after the exp fix), the dominant kernels of the lml step and of the gradient step
python tools/cov_bench.py > gpurun_out/cov_bench_r02c.json 2> gpurun_out/cov_bench.err
python tools/cov_bench.py --profile > gpurun_out/cov_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cov_ -c 4 -o gpurun_out/cov_r02b python tools/cov_bench.py --profile > gpurun_out/ncu_cov.log 2>&1
tail -2 gpurun_out/ncu_cov.log
python bench.py --grad --no-cpu --no-configs --steps 3 > gpurun_out/b8g.json 2> gpurun_out/b8g.err && \
ncu --set full --clock-control none --import-source on -k regex:"lk_below|lk_gradc|lk_minv" -s 60 -c 15 -o gpurun_out/grad_r02 python bench.py --grad --no-cpu --no-configs --steps 1 > gpurun_out/ncu_grad.log 2>&1
tail -2 gpurun_out/ncu_grad.log; tail -c 300 gpurun_out/b8g.err
python bench.py --steps 10 > gpurun_out/b8.json 2> gpurun_out/b8.err; tail -c 300 gpurun_out/b8.err
