#!/bin/bash
# ncu launch list of gpl_lml_large n=8192 with the INT8 trailing passes, look-ahead without the worker CTA (ncu serialises kernels)
timeout 120 python tools/run_c5.py 8192 chol_variant=3 > gpurun_out/c5_57.log 2>&1 || exit 1; tail -1 gpurun_out/c5_57.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c5_int8.csv python tools/run_c5.py 8192 chol_variant=3 > gpurun_out/ncu57.log 2>&1; echo rc=$?; tail -1 gpurun_out/ncu57.log
