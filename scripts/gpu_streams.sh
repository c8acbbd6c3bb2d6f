#!/bin/bash
# bench.py with different numbers of independent batches in flight
mkdir -p gpurun_out
for st in 4 2 8 16; do
  echo "== bench --streams $st"
  timeout 600 python bench.py --no-cpu-baseline --no-scale --streams $st > gpurun_out/bench_st$st.json 2> gpurun_out/bench_st$st.err || tail -5 gpurun_out/bench_st$st.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_st$st.json').read().strip().splitlines()[-1])
    print('value %.1fM  ms/step %.5f  single %s  e2e %.1fM  performed/nominal %d/%d  roofline %.3f timed-region %.3f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['single_stream'] and '%.1fM' % (d['single_stream']['value']/1e6), d['e2e']['value']/1e6, d['env_steps']['performed'], d['env_steps']['nominal'], d['roofline']['frac'], d['roofline']['timed_region_frac'], d['clocks']))
except Exception as ex:
    print('bench failed', ex)
PY
done
