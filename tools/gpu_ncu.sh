#!/bin/bash
# ncu --set full of one full-pool k_pt_extend launch (the third: second iteration of sub-pool 0; after a plain run exits 0).  Usage: tools/gpu_ncu.sh <tag> [bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$tag.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_pt_extend$ -s 2 -c 1 -f -o gpurun_out/prof_extend_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_extend_$tag.ncu-rep
