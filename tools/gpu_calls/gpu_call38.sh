#!/bin/bash
timeout 120 python tools/_i8dbg.py 4096 2>&1 | tail -20
