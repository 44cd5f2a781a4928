#!/bin/bash
# cov kernels (tests + bandwidth + one ncu capture), then the full bench line with the configs block
python -m pytest tests -m gpu -x -q > gpurun_out/t3.log 2>&1; tail -3 gpurun_out/t3.log
python tools/cov_bench.py > gpurun_out/cov_bench_r02.json 2> gpurun_out/cov_bench.err; tail -c 300 gpurun_out/cov_bench.err
python tools/cov_bench.py --profile > gpurun_out/cov_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cov_ -c 4 -o gpurun_out/cov_r02 python tools/cov_bench.py --profile > gpurun_out/ncu_cov.log 2>&1
tail -3 gpurun_out/ncu_cov.log
python bench.py --steps 10 > gpurun_out/b3.json 2> gpurun_out/b3.err; tail -c 500 gpurun_out/b3.err
