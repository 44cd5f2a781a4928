#!/bin/bash
python bench.py --grad --no-cpu --no-configs --steps 10 > gpurun_out/b55g.json 2> gpurun_out/b55g.err
python - <<'PY'
import json
for f in ('gpurun_out/b47g.json','gpurun_out/b55g.json'):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['value']), round(j['ms_per_step'],3), json.dumps(j.get('phase_ms') or j['roofline'].get('phases') or {k:v for k,v in j['roofline'].items() if 'ms' in k})[:400])
    except Exception as e: print(f, 'ERR', e)
PY
