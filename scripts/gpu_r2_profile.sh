#!/bin/bash
# round-2 ncu evidence (one GPU): launch list of the bench sequence, full captures of the multi-step kernel (4096 envs, steady
# state), the single-step kernel at full occupancy (262144 envs) and the crowd kernel (4096 envs x 20 humans)
mkdir -p gpurun_out
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_|scene_" -s 3200 -c 384 --csv --log-file gpurun_out/r2_launches.csv python scripts/eager_loop.py > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
echo "== full: multi-step kernel, 4096 envs, steady state"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 1600 -c 1 -o gpurun_out/r2_prof_step_n python scripts/eager_loop.py > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
echo "== full: single-step kernel, 262144 envs"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 14 -c 1 -o gpurun_out/r2_prof_step_big python scripts/eager_big.py > gpurun_out/ncu_big.log 2>&1; tail -1 gpurun_out/ncu_big.log
echo "== full: crowd kernel, 4096 envs x 20 humans"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/r2_prof_crowd_final python scripts/eager_crowd.py 20 4096 > gpurun_out/ncu_crowd.log 2>&1; tail -1 gpurun_out/ncu_crowd.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
