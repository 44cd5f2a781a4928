#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/b50_n8.json 2> gpurun_out/b50_n8.err; tail -c 300 gpurun_out/b50_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/b50_n4.json 2> gpurun_out/b50_n4.err; tail -c 300 gpurun_out/b50_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b50_n2.json 2> gpurun_out/b50_n2.err; tail -c 300 gpurun_out/b50_n2.err
python - <<'PY'
import json
for n in (8,4,2):
    try:
        j=json.loads(open(f'gpurun_out/b50_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(j['value']), j['ms_per_step'], 'strong', {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else None) for k,v in j.get('strong',{}).items()} , 'multi_abi', j.get('multi_abi',{}).get('value'), 'err', j.get('oracle_max_rel_err'))
    except Exception as e:
        print(n, 'ERR', e)
PY
