timeout 60 python tools/run_c5.py 8192 3 2>&1 | tail -1
timeout 60 python tools/run_c5.py 2048 2 2>&1 | tail -1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do timeout 60 python tools/run_c5.py 8192 2 2>&1 | tail -1 | cut -c1-60; done
