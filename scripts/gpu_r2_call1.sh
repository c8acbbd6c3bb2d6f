#!/bin/bash
# round 2, call 1: full GPU suite at HEAD (no -x: every failure listed), parity of the f32x2 / no-ROT variants, reference arm
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
echo "== pytest -m gpu (HEAD)"; timeout 1200 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2c1_pytest_head.txt
for v in f32x2 norot_f32x2; do
  echo "== pytest -m gpu on variant $v"
  CROWDSIM_B200_LIB=$PWD/build_probe/lib_$v.so timeout 900 python -m pytest tests -m gpu -q --timeout=300 -k "not external_rot and not unicycle and not rot" 2>&1 | tail -4 | tee gpurun_out/r2c1_pytest_$v.txt
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --no-python-loop > gpurun_out/r2c1_bench_ref_driverflags.json 2>gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/r2c1_bench_ref_driverflags.json
echo "== bench driver flags"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench_driverflags.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/r2c1_bench_driverflags.json
