#!/bin/bash
# Builds libwrt_b200.so with extra nvcc -D flags (A/B experiments): tools/build_variant.sh -DWRT_REFILL_THRESHOLD=30
cd "$(dirname "$0")/../winmad-s-raytracer-v1.0_b200"
for f in scene_upload trace_kernels pt_wavefront bdpt_wavefront; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -diag-suppress 177 "$@" -c csrc/$f.cu -o build/$f.cu.o || exit 1
done
nvcc -shared -o libwrt_b200.so build/*.o 2>/dev/null
