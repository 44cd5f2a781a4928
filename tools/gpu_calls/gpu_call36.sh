#!/bin/bash
# rehearsal of the round-end sequence after the tools/ozaki addition
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke36.log 2>&1; tail -1 gpurun_out/smoke36.log
python -m pytest tests -m gpu -x -q > gpurun_out/t36.log 2>&1; tail -3 gpurun_out/t36.log
