#!/bin/bash
# GPU iteration: parity tests, then bench variants (prefetch interval, e2e batches in flight). Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -4
for pe in 4 1 8; do
  echo "== bench --prefetch-every $pe"
  timeout 600 python bench.py --no-cpu-baseline --no-scale --prefetch-every $pe --e2e-batches 8 > gpurun_out/bench_pe$pe.json 2> gpurun_out/bench_pe$pe.err || tail -5 gpurun_out/bench_pe$pe.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_pe$pe.json').read().strip().splitlines()[-1])
    print('value %.1fM  ms/step %.5f  e2e %.1fM (single %.1fM)  env_steps %s  roofline %.3f  launches %d' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['single_batch_blocking']/1e6, d['env_steps'], d['roofline']['frac'], d['gpu_launches']))
except Exception as ex:
    print('bench failed', ex)
PY
done
for eb in 2 4 16 32; do
  echo "== bench --e2e-batches $eb (short)"
  timeout 600 python bench.py --no-cpu-baseline --no-scale --steps 2560 --warmup 256 --e2e-batches $eb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e %.1fM single %.1fM' % (d['e2e']['value']/1e6, d['e2e']['single_batch_blocking']/1e6))"
done
