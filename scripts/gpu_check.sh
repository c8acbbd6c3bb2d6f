#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench (+ optional ncu launch list). Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_|reset_|lookahead|pack_" -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 400 --warmup 100 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
  echo "== ncu full on step kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 100 -c 3 -o gpurun_out/prof_step python bench.py --steps 60 --warmup 50 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
