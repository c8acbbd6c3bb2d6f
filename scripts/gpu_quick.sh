#!/bin/bash
# Quick GPU iteration: parity + bench. Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -8
echo "== bench" ; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
