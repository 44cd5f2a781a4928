set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lk_potrf_warp -s 10 -c 1 -o gpurun_out/pw2 -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/pw2.ncu-rep
