#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t20.log 2>&1; tail -3 gpurun_out/t20.log
python bench.py --steps 10 > gpurun_out/b20.json 2> gpurun_out/b20.err; tail -c 300 gpurun_out/b20.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b20.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']))
for k in ('c3_mcmc','c1_mcmc'):
    v=j['configs'][k]; print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','timing','note')})
PY
