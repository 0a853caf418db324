#!/bin/bash
# Builds a variant of the product library with extra nvcc -D flags (A/B experiments):
#   tools/build_variant.sh [-o libname.so] -DWRT_TRACE_SCHED=5 -DWRT_LANE_RAYS=3
# Default output is winmad-s-raytracer-v1.0_b200/libwrt_b200.so (overwrites the default build; rebuild with build.py --force).
cd "$(dirname "$0")/../winmad-s-raytracer-v1.0_b200"
out=libwrt_b200.so
if [ "$1" = "-o" ]; then out=$2; shift 2; fi
od=build/variant_$$; mkdir -p $od
for f in scene_upload trace_kernels pt_wavefront bdpt_wavefront debug_kernels multi_gpu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -diag-suppress 177 "$@" -c csrc/$f.cu -o $od/$f.cu.o &
done
wait
nvcc -shared -o $out $od/*.cu.o build/scene_layout.cpp.o build/kd_build.cpp.o build/scene_io.cpp.o build/host_api.cpp.o 2>/dev/null; rc=$?
rm -rf $od
exit $rc
