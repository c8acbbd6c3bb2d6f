#!/bin/bash
# Quick GPU iteration: parity subset + bench + packing sweep. Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest step parity" ; timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout=120 -k "random_scenes or full_suites_from" 2>&1 | tail -5
echo "== bench" ; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== tune" ; timeout 900 python scripts/tune_epw.py 5 2>&1 | tee gpurun_out/tune_epw_n5.txt | tail -12
