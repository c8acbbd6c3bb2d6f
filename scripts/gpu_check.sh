#!/bin/bash
# parity + probe + short bench after a kernel change
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -3
echo "== probe"; timeout 300 build_probe/probe 12 | grep -A2 "^B=" | tee gpurun_out/probe_check.txt
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err || tail -5 gpurun_out/bench_check.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1])
print('value %.1fM  single-stream %.1fM  kernel %.2f us  e2e %.1fM (blocking %.1fM)  1Mi-env %.0f us (%.3f of HBM peak)' % (d['value']/1e6, d['single_stream']['value']/1e6, d['roofline']['avg_launch_us'], d['e2e']['value']/1e6, d['e2e']['single_batch_blocking']/1e6, d['scale']['us_per_launch'], d['scale']['roofline_frac']))
PY
