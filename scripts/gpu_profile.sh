#!/bin/bash
# Round-end evidence run (1 GPU): tests, bench (+ reference arm), ncu launch list + full capture of the step kernel,
# latency probe, packing sweep. Outputs under gpurun_out/; summaries are copied to profiles/ by hand afterwards.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q --timeout=180 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
echo "== ncu launch list (eager launches of the same step/prefetch sequence)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_|scene_|lookahead|pack_" -s 600 -c 800 --csv --log-file gpurun_out/launches.csv python scripts/eager_loop.py > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full on the step kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 300 -c 2 -o gpurun_out/prof_step python scripts/eager_loop.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log

echo "== ncu full on the step kernel at full occupancy (262144 envs, mid-episode)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_flat -s 14 -c 1 -o gpurun_out/prof_step_big python scripts/eager_big.py > gpurun_out/ncu_big.log 2>&1; tail -2 gpurun_out/ncu_big.log
echo "== latency probe (mid-episode states)" ; ./build_probe/probe 12 | tee gpurun_out/latency_probe.txt
echo "== configs 1/3/4" ; timeout 600 python scripts/measure_misc.py | tee gpurun_out/measure_misc.json
