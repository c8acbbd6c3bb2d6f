#!/bin/bash
# Last call of the round: full GPU test suite, smoke, default bench, DRAM traffic of the step kernel at 2^20 envs.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout=180 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "== ncu dram bytes at 2^20 envs"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:step_flat -s 14 -c 1 --csv --log-file gpurun_out/step_1Mi_dram.csv python scripts/eager_big.py 1048576 > gpurun_out/ncu_1Mi.log 2>&1; tail -1 gpurun_out/ncu_1Mi.log; cat gpurun_out/step_1Mi_dram.csv | tail -6
