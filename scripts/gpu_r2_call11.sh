#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -12 | tee gpurun_out/r2c11_pytest.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; tail -4 gpurun_out/r2_sanitizer_$tool.txt
done
