#!/bin/bash
for c in -1 -2 -3 -4 98; do GPL_I8_CTAS=$c timeout 120 python tools/_i8t.py 8192 8 2>&1 | tail -1; done
GPL_I8_CTAS=-2 timeout 120 python tools/_i8t.py 16384 8 2>&1 | tail -1
GPL_I8_CTAS=98 timeout 120 python tools/_i8t.py 16384 8 2>&1 | tail -1
