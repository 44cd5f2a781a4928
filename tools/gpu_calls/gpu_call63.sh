#!/bin/bash
# rehearsal of the round-end sequence with the INT8 trailing path in the product
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke63.log 2>&1; tail -1 gpurun_out/smoke63.log
python -m pytest tests -m gpu -x -q > gpurun_out/t63.log 2>&1; tail -3 gpurun_out/t63.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/b63_ref.json 2> gpurun_out/b63_ref.err; cut -c1-163 gpurun_out/b63_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b63.json 2> gpurun_out/b63.err; tail -c 300 gpurun_out/b63.err
python bench.py --grad --no-cpu --no-configs --steps 10 > gpurun_out/b63g.json 2> gpurun_out/b63g.err
python - <<'PY'
import json
for f in ('gpurun_out/b63.json','gpurun_out/b63g.json'):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],3), round(j['roofline']['whole_step']['frac'],3), j['oracle_max_rel_err'])
PY
