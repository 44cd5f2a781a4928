timeout 60 python tools/run_c5.py 2048 2 2>&1 | tail -1
timeout 60 python tools/run_c5.py 8192 3 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
