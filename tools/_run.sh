set -x
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r01.err | tail -1 > gpurun_out/bench_r01.json || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r01_ref.json
timeout 600 python tools/bench_configs.py > gpurun_out/configs_r01.json 2>gpurun_out/configs_r01.err
timeout 300 python tools/sweep_small.py > gpurun_out/sweep_small.txt 2>&1
CHOLV=3 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:big_ -c 400 --csv --log-file gpurun_out/launches_r01_c5.csv python tools/run_c5.py 8192 1 > gpurun_out/ncu_c5.log 2>&1
