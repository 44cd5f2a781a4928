timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -2
timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a: round(b,3) for a,b in v.items()})"
