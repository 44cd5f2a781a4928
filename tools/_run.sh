set -x
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lk_below|lk_potrf_warp" -s 22 -c 2 -o gpurun_out/r01_final -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/r01_final.ncu-rep
