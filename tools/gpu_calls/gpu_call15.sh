#!/bin/bash
python -m pytest tests/test_gpu_ou_separable.py -x -q > gpurun_out/t15a.log 2>&1; tail -15 gpurun_out/t15a.log
python -m pytest tests -m gpu -x -q > gpurun_out/t15.log 2>&1; tail -3 gpurun_out/t15.log
python bench.py --steps 10 --no-configs --no-cpu > gpurun_out/b15.json 2> gpurun_out/b15.err; tail -c 300 gpurun_out/b15.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b15.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), {k:round(v,3) for k,v in j['roofline']['per_kernel_ms_per_step'].items()}, j['oracle_max_rel_err'], j['arms_max_rel_diff'])
PY
