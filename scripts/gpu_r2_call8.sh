#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/r2_prof_crowd python scripts/eager_crowd.py 20 4096 > gpurun_out/ncu_crowd.log 2>&1; tail -2 gpurun_out/ncu_crowd.log
ls -la gpurun_out/*.ncu-rep
