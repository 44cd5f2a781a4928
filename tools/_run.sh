set -x
CHOLV=3 timeout 300 python tools/run_c5.py 8192 1 > gpurun_out/c5_v3.log 2>&1 || exit 1
CHOLV=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:big_trail -s 3 -c 1 -o gpurun_out/trail2 -f python tools/run_c5.py 8192 1 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/trail2.ncu-rep
