#!/bin/bash
timeout 600 python tools/i8_large.py 8192 12288 16384 > gpurun_out/i8_41.log 2>&1; echo rc=$?; tail -14 gpurun_out/i8_41.log
