#!/bin/bash
# last GPU call of a round, at the final commit: everything the driver will run + the evidence files of profiles/
mkdir -p gpurun_out
echo "== pytest -m gpu -x -q (the driver's command)"; timeout 1800 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_final_pytest.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r2_final_smoke.txt
echo "== reference arm (driver flags)"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2>gpurun_out/bench_ref.err || tail -5 gpurun_out/bench_ref.err
echo "== bench (driver flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2>gpurun_out/bench.err || tail -20 gpurun_out/bench.err
echo "== bench, BASELINE config 4 (20 humans, square crossing)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --humans 20 --rule square_crossing --no-cpu-baseline --no-scale > gpurun_out/r02_bench_cfg4_n1.json 2>gpurun_out/bench4.err || tail -5 gpurun_out/bench4.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_cfg4_n1.json').read().strip().splitlines()[-1])
    print('config 4: value %.1fM single %.1fM roofline %s %.2f us/launch frac %.4f e2e %.1fM' % (d['value']/1e6, d['single_batch']['value']/1e6, d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['e2e']['value']/1e6))
except Exception as e:
    print('config 4 bench unreadable', e)
PY
echo "== configs 1/3/4"; timeout 900 python scripts/measure_misc.py > gpurun_out/r02_configs_1_3_4.json 2>gpurun_out/misc.err || tail -5 gpurun_out/misc.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/r02_bench_reference_arm.json').read().strip().splitlines()[-1])
print('value %.1fM  single %.1fM (rot %.1fM)  roofline %.2f us/launch frac %.4f traffic %s  e2e %.1fM  1Mi %.0f us (%.3f)  cpu_baseline %.1fM/%s  reference arm %.1fM/%s threads' % (
    d['value']/1e6, d['single_batch']['value']/1e6, d['single_batch']['rotating_value']/1e6, d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['traffic'],
    d['e2e']['value']/1e6, d['scale']['us_per_launch'], d['scale']['roofline_frac'], d['cpu_baseline']['value']/1e6, d['cpu_baseline']['cores'], r['value']/1e6, r['cpu_baseline']['cores']))
print('ratio e2e/reference %.1f  device/reference %.1f  parity %s  clocks %s %s' % (d['e2e']['value']/r['value'], d['value']/r['value'], d['parity_500_cases'].get('match'), d['clocks']['sm_mhz'], d['clocks']['reasons']))
print(open('gpurun_out/r02_configs_1_3_4.json').read()[:1500])
PY
