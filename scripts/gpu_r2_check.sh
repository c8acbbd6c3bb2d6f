#!/bin/bash
# tests + default bench after a change (one GPU)
T=${1:-r2chk}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -12 | tee gpurun_out/${T}_pytest.txt
echo "== bench driver flags"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_driverflags.json 2>gpurun_out/bench.err || tail -20 gpurun_out/bench.err
python - $T <<'PY'
import json, sys
f = 'gpurun_out/%s_bench_driverflags.json' % sys.argv[1]
d = json.loads(open(f).read().strip().splitlines()[-1])
print('value %.1fM  ms/step %.5f  replays %d  rounds-median %.1fM  single %.1fM (rot %.1fM)  roof launch %.2f us frac %.4f (1-step %.2f us frac %.4f)  e2e %.1fM (python rr %.1fM, blocking %.1fM, %d B down)  1Mi %.0f us (%.3f)  cpu %.1fM on %s' % (
    d['value']/1e6, d['ms_per_step'], d['timed_region']['replays'], d['rounds']['median_value']/1e6, d['single_batch']['value']/1e6, d['single_batch']['rotating_value']/1e6,
    d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['single_step_kernel']['avg_launch_us'], d['roofline']['single_step_kernel']['frac'],
    d['e2e']['value']/1e6, d['e2e']['python_round_robin']/1e6, d['e2e']['single_batch_blocking']/1e6, d['e2e']['d2h_bytes_per_step'], d['scale']['us_per_launch'], d['scale']['roofline_frac'],
    d['cpu_baseline']['value']/1e6, d['cpu_baseline']['cores']))
print('   env_steps', d['env_steps']['performed'], d['env_steps']['nominal'], 'parity', d['parity_500_cases'].get('match'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['samples'], 'launches', d['gpu_launches'])
PY
