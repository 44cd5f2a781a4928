python bench.py --steps 10 --warmup 3 > gpurun_out/bench20.json 2> gpurun_out/bench20.err; tail -2 gpurun_out/bench20.err
python -c "
import json,sys; j=json.load(open('gpurun_out/bench20.json')); print(json.dumps({k:j[k] for k in ['value','ms_per_step','gpu_launches','clocks','cpu_baseline','e2e']},indent=0)); print(json.dumps(j['roofline'],indent=0))"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench20_ref.json; cut -c1-300 gpurun_out/bench20_ref.json
