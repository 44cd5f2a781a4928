#!/bin/bash
for b in 8 16 32; do for c in 64 98 128; do GPL_I8_BLOCK=$b GPL_I8_CTAS=$c timeout 120 python tools/_i8t.py 8192 8 2>&1 | tail -1 | sed "s/^/block=$b /"; done; done
GPL_I8_BLOCK=8 timeout 120 python tools/_i8t.py 6144 8 2>&1 | tail -1
GPL_I8_BLOCK=8 timeout 120 python tools/_i8t.py 4096 8 2>&1 | tail -1
