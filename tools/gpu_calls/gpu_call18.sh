#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t18.log 2>&1; tail -3 gpurun_out/t18.log
python bench.py --grad --no-cpu --no-configs --steps 10 > gpurun_out/b18g.json 2> gpurun_out/b18g.err; tail -c 300 gpurun_out/b18g.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b18g.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['per_kernel_ms_per_step'].items()}, j['oracle_max_rel_err'])
PY
