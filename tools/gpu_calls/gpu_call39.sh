#!/bin/bash
# option trail_int8 with the look-ahead split of the INT8 pass: sweep of the persistent CTA count of the deferred part
for c in 84 104; do echo "== GPL_I8_CTAS=$c"; GPL_I8_CTAS=$c timeout 300 python tools/i8_large.py 4096 6144 8192 2>&1 | grep -v "trail_int8=7" | tail -9; done > gpurun_out/i8_39.log 2>&1
cat gpurun_out/i8_39.log
