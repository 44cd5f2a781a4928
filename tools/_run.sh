timeout 600 python tools/gpu_debug.py > gpurun_out/debug8.log 2>&1; grep -E "ERR|c2 B" gpurun_out/debug8.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python -c "
import json,sys; j=json.load(open('gpurun_out/bench8.json')); print('default', round(j['value']), round(j['roofline']['frac'],4), j['arms_max_rel_diff'], j['not_pd_items'])"
