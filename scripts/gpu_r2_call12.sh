#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout=300 -x -k "human_times" 2>&1 | tail -40 | tee gpurun_out/r2c13_ht.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -12 | tee gpurun_out/r2c13_pytest.txt
