#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2c7_pytest.txt
echo "== crowd kernel timing"; timeout 600 python scripts/time_crowd.py 20 | tee gpurun_out/r2c7_time_crowd.txt; timeout 300 python scripts/time_crowd.py 10 4096 | tee -a gpurun_out/r2c7_time_crowd.txt
