#!/bin/bash
# zero-tile skipping in the gradient phases: tests, then the bench configs (C3 lml / grad, sampler)
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t54.log 2>&1; tail -3 gpurun_out/t54.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/b54.json 2> gpurun_out/b54.err; tail -c 200 gpurun_out/b54.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b54.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), round(j['roofline']['whole_step']['frac'],3), j['oracle_max_rel_err'])
c=j['configs']
for k in ('c3_lml','c3_grad','golden_n923','c2_grad','c3_mcmc','c1_mcmc'):
    e=c.get(k,{}); print(k, {kk:(round(v,4) if isinstance(v,float) else v) for kk,v in e.items() if not isinstance(v,(dict,list,str))})
PY
