timeout 600 python tools/gpu_debug.py > gpurun_out/debug15.log 2>&1; grep -E "ERR|ok " gpurun_out/debug15.log | cut -c1-150 | grep -E "ERR|lml n|grad|c2|posterior n=700"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -2 gpurun_out/bench15.err
python -c "
import json,sys; j=json.load(open('gpurun_out/bench15.json')); print('lockstep', round(j['value']), round(j['roofline']['frac'],4), j['arms_max_rel_diff'], j['not_pd_items'], j['gpu_launches'])"
