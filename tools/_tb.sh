timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_pw.json
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_pw.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['gpu_launches'])
PY
