#!/bin/bash
# lk_below: bulk-copy + mbarrier ring instead of LDGSTS + block barriers
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/b51.json 2> gpurun_out/b51.err; tail -c 200 gpurun_out/b51.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/b51.json').read().strip().splitlines()[-1])
print(round(j['value']), round(j['ms_per_step'],3), 'frac', round(j['roofline']['frac'],3), round(j['roofline']['whole_step']['frac'],3), j['oracle_max_rel_err'], j.get('phases') or j['roofline'].get('per_kernel_ms'))
PY
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_grad.py tests/test_gpu_ou_separable.py tests/test_gpu_hygiene.py -x -q -m gpu > gpurun_out/t51.log 2>&1; tail -3 gpurun_out/t51.log
