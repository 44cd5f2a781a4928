for v in "" _nostream_zg; do
  GAPLAC_B200_LIB=$PWD/gaplac_b200/libgaplac_b200$v.so python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench14$v.json 2> gpurun_out/bench14$v.err
  python -c "
import json,sys; j=json.load(open('gpurun_out/bench14$v.json')); print('variant[$v]', round(j['value']), round(j['roofline']['frac'],4), j['arms_max_rel_diff'], j['not_pd_items'])"
done
