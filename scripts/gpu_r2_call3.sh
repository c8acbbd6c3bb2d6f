#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2c3_pytest.txt
echo "== probe round-1 kernel"; timeout 300 build_probe/probe_r1 12 | grep -A2 "^B=" | tee gpurun_out/r2c3_probe_r1.txt
echo "== probe new kernel"; timeout 300 build_probe/probe 12 | grep -E -A5 "^B=" | tee gpurun_out/r2c3_probe_new.txt
