#!/bin/bash
mkdir -p gpurun_out
for mb in 5 6; do echo "== MULTI MINBLOCKS=$mb"; timeout 300 build_probe/probe_multi_mb$mb 12 | grep -E "^B=|step_n|dense" | tee gpurun_out/r2_probe_multi_mb$mb.txt; done
echo "== pytest"; timeout 1800 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== misc"; timeout 900 python scripts/measure_misc.py > gpurun_out/r02_configs_1_3_4.json 2>gpurun_out/misc.err || tail -5 gpurun_out/misc.err; tail -12 gpurun_out/r02_configs_1_3_4.json
