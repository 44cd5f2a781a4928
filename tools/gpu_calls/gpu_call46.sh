#!/bin/bash
for c in 120 136 148; do GPL_I8_CTAS=$c timeout 120 python tools/_i8t.py 16384 8 2>&1 | tail -1; done
for c in 120 136; do GPL_I8_CTAS=$c timeout 120 python tools/_i8t.py 12288 8 2>&1 | tail -1; done
GPL_I8_CTAS=98 timeout 120 python tools/_i8t.py 12288 8 2>&1 | tail -1
