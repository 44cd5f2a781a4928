#!/bin/bash
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/b53.json 2> gpurun_out/b53.err; tail -c 200 gpurun_out/b53.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b53.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), round(j['roofline']['whole_step']['frac'],3), j['oracle_max_rel_err'])
c=j['configs']
for k in ('c3_lml','c3_grad','golden_n923','c2_grad'):
    e=c.get(k,{}); print(k, {kk:(round(v,4) if isinstance(v,float) else v) for kk,v in e.items() if kk in ('ms','evals_per_s','frac_fp64_peak')})
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/t53.log 2>&1; tail -2 gpurun_out/t53.log
