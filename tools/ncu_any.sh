#!/bin/bash
# usage: ncu_any.sh <kernel regex> <bench args...>: one ncu --set full capture (2 launches) of the matching kernels
mkdir -p gpurun_out
K=$1; shift
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 3 -o gpurun_out/ncu_$K -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | tail -3
