#!/bin/bash
# round 2, call 2: new small-crowd kernel (lane-parallel lp3, step_n): parity + stage probe vs the round-1 kernel on the same box
mkdir -p gpurun_out
echo "== pytest -m gpu (parity file first)"; timeout 1500 python -m pytest tests/test_cuda_0_parity.py -m gpu -q --timeout=300 2>&1 | tail -15 | tee gpurun_out/r2c2_pytest_parity.txt
echo "== probe round-1 kernel"; timeout 300 build_probe/probe_r1 12 | grep -A2 "^B=" | tee gpurun_out/r2c2_probe_r1.txt
echo "== probe new kernel (80 regs)"; timeout 300 build_probe/probe 12 | grep -E -A5 "^B=" | tee gpurun_out/r2c2_probe_new.txt
for mb in 4 5; do echo "== probe new kernel MINBLOCKS=$mb"; timeout 300 build_probe/probe_mb$mb 12 | grep -A2 "^B=" | tee gpurun_out/r2c2_probe_new_mb$mb.txt; done
