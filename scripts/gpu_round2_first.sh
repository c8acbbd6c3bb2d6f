#!/bin/bash
# Suggested FIRST gpurun call of the next round (build the variants first: bash scripts/build_variants.sh):
#  1. the new CPU arm (one OpenMP region) on the GPU box's cores -- not measured there yet
#  2. parity suite on the f32x2 / no-ROT variants (the ROT tests are expected to fail on a no-ROT build: deselected)
#  3. bench A/B of the variants
mkdir -p gpurun_out
echo "== reference arm (CPU, one OpenMP region)"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; cut -c1-700 gpurun_out/bench_ref.json
for v in f32x2 norot_f32x2; do
  echo "== pytest -m gpu on variant $v"
  CROWDSIM_B200_LIB=$PWD/build_probe/lib_$v.so timeout 900 python -m pytest tests -m gpu -q --timeout=180 -k "not external_rot and not unicycle" 2>&1 | tail -4
done
VARIANTS="default f32x2 norot norot_f32x2" bash scripts/gpu_variants.sh
