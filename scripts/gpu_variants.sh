#!/bin/bash
# A/B of prebuilt library variants (build_probe/lib_*.so, built in the dev container) through bench.py.
mkdir -p gpurun_out
for v in ${VARIANTS:-default wpb8}; do
  if [ $v = default ]; then unset CROWDSIM_B200_LIB; else export CROWDSIM_B200_LIB=$PWD/build_probe/lib_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --e2e-batches 4 2>gpurun_out/var_$v.err > gpurun_out/var_$v.json || tail -3 gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1])
    print('%-10s value %.1fM  single-stream %.1fM  kernel %.2f us  1Mi-env %.0f us (%.3f of HBM peak)' % ('$v', d['value']/1e6, d['single_stream']['value']/1e6, d['roofline']['avg_launch_us'], d['scale']['us_per_launch'], d['scale']['roofline_frac']))
except Exception as ex:
    print('$v failed', ex)
PY
done
