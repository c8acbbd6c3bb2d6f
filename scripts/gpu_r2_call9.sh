#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2c10_pytest.txt
echo "== crowd kernel timing (5 blocks/SM)"; timeout 600 python scripts/time_crowd.py 20 | tee gpurun_out/r2c10_time_crowd.txt
for mb in 4 6; do echo "== $mb blocks/SM"; CROWDSIM_B200_LIB=$PWD/build_probe/lib_mid_mb$mb.so timeout 600 python scripts/time_crowd.py 20 | grep crowd | tee -a gpurun_out/r2c10_time_crowd.txt; done
timeout 300 python scripts/time_crowd.py 10 4096 | tee -a gpurun_out/r2c10_time_crowd.txt
