#!/bin/bash
# one ncu --set full capture of kernels A, B, C (second batch of the run), source view included
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 3 -o gpurun_out/fused_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --rot-per-step 512 "$@" > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
