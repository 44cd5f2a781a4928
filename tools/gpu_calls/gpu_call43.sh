#!/bin/bash
for sk in 0 1 2 4 6 7; do GPL_I8_SKIP=$sk timeout 120 python tools/_i8t.py 8192 8 2>&1 | tail -1; done
