timeout 120 python tools/run_c5.py 1500 2 2>&1 | tail -2
timeout 120 python tools/run_c5.py 8192 4 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
