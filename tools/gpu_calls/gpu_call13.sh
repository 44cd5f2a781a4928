#!/bin/bash
# round-end rehearsal: what the driver runs (smoke, gpu tests, both bench arms)
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke13.log 2>&1; tail -2 gpurun_out/smoke13.log
python -m pytest tests -m gpu -x -q > gpurun_out/t13.log 2>&1; tail -3 gpurun_out/t13.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/b13_ref.json 2> gpurun_out/b13_ref.err; cut -c1-300 gpurun_out/b13_ref.json
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b13.json 2> gpurun_out/b13.err ) 2>&1 | grep real; tail -c 300 gpurun_out/b13.err
