#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2c6_pytest.txt
echo "== probe new kernel"; timeout 300 build_probe/probe 12 | grep -E -A5 "^B=" | tee gpurun_out/r2c6_probe_new.txt
echo "== bench driver flags"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c6_bench_driverflags.json 2>gpurun_out/bench.err || tail -20 gpurun_out/bench.err
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2c6_bench_default.json 2>gpurun_out/bench2.err || tail -20 gpurun_out/bench2.err
python - <<'PY'
import json
for f in ('r2c6_bench_driverflags', 'r2c6_bench_default'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value %.1fM  ms/step %.5f  replays %d  rounds-median %.1fM  single %.1fM (rot %.1fM)  roof launch %.2f us frac %.4f (1-step %.2f us frac %.4f)  e2e %.1fM (blocking %.1fM, %d B down)  1Mi %.0f us (%.3f)' % (
        d['value']/1e6, d['ms_per_step'], d['timed_region']['replays'], d['rounds']['median_value']/1e6, d['single_batch']['value']/1e6, d['single_batch']['rotating_value']/1e6,
        d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['single_step_kernel']['avg_launch_us'], d['roofline']['single_step_kernel']['frac'],
        d['e2e']['value']/1e6, d['e2e']['single_batch_blocking']/1e6, d['e2e']['d2h_bytes_per_step'], d['scale']['us_per_launch'], d['scale']['roofline_frac']))
    print('   env_steps', d['env_steps']['performed'], d['env_steps']['nominal'], 'parity', d['parity_500_cases'].get('match'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['samples'])
PY
echo "== variants"
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.1fM single %.1fM (rot %.1fM) roof %.2f us/launch frac %.4f e2e %.1fM' % (d['value']/1e6, d['single_batch']['value']/1e6, d['single_batch']['rotating_value']/1e6, d['roofline']['avg_launch_us'], d['roofline']['frac'], d['e2e']['value']/1e6))
PY
}
for c in 1 2 4 16; do timeout 600 python bench.py --no-cpu-baseline --no-scale --chunk $c > gpurun_out/r2c6_bench_chunk$c.json 2>/dev/null; summ gpurun_out/r2c6_bench_chunk$c.json; done
for mb in 4 5; do CROWDSIM_B200_LIB=$PWD/build_probe/lib_multi_mb$mb.so timeout 600 python bench.py --no-cpu-baseline --no-scale > gpurun_out/r2c6_bench_multi_mb$mb.json 2>/dev/null; summ gpurun_out/r2c6_bench_multi_mb$mb.json; done
for st in 4 8 32; do timeout 600 python bench.py --no-cpu-baseline --no-scale --streams $st > gpurun_out/r2c6_bench_streams$st.json 2>/dev/null; summ gpurun_out/r2c6_bench_streams$st.json; done
