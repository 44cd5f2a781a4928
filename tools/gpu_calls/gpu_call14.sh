#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t14.log 2>&1; tail -3 gpurun_out/t14.log
python bench.py --steps 10 --no-configs --no-cpu > gpurun_out/b14.json 2> gpurun_out/b14.err; tail -c 300 gpurun_out/b14.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b14.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['per_kernel_ms_per_step'].items()}, j['strong']['c3']['ms_per_step'])
PY
