#!/bin/bash
timeout 120 python tools/_i8dbg.py 8192 2>&1 | tail -36
