#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t65.log 2>&1; tail -3 gpurun_out/t65.log
