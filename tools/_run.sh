for v in "" _nosync; do
  if [ -n "$v" ]; then export GAPLAC_B200_LIB=$PWD/gaplac_b200/libgaplac_b200$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_v$v.json
  python - <<PY
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('$v', d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms_per_step'], d['not_pd_items'])
PY
done
