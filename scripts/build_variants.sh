#!/bin/bash
# Build A/B variants of libcrowdsim_b200.so into build_probe/ (git-ignored, travels with gpurun); run them with
#   VARIANTS="default f32x2" bash scripts/gpu_variants.sh      (bench.py through CROWDSIM_B200_LIB)
# Each variant must also pass `CROWDSIM_B200_LIB=$PWD/build_probe/lib_<v>.so python -m pytest tests -m gpu` before it is adopted.
set -e
mkdir -p build_probe
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC -shared -cudart shared"
S="crowdnav_b200/csrc/step_kernel.cu crowdnav_b200/csrc/reset_kernel.cu crowdnav_b200/csrc/pack_kernel.cu"
build() { nvcc $F $2 $S -o build_probe/lib_$1.so & }
build f32x2 "-DCS_F32X2"                       # packed FADD2/FMUL2 for the (x, y) arithmetic (unmeasured)
build f32x2_mb5 "-DCS_F32X2 -DCS_FLAT_MINBLOCKS=5"
build norot "-DCS_FLAT_NO_ROT"                   # unicycle code compiled out: only for holonomic / ORCA robots
build norot_f32x2 "-DCS_FLAT_NO_ROT -DCS_F32X2"
wait
ls -la build_probe/*.so
