#!/bin/bash
python bench.py --steps 20 --warmup 5 > gpurun_out/b17.json 2> gpurun_out/b17.err; tail -c 300 gpurun_out/b17.err
python bench.py --grad --no-cpu --no-configs --steps 10 > gpurun_out/b17g.json 2> gpurun_out/b17g.err; tail -c 300 gpurun_out/b17g.err
python tools/cov_bench.py > gpurun_out/cov_bench_r02d.json 2>/dev/null
