#!/bin/bash
# compute-sanitizer over every fused pipeline (round 2 kernels: TMA tiles, TMEM stash, staged rows)
mkdir -p gpurun_out
out=gpurun_out/sanitizer.txt
: > $out
run() { echo "### $*" >> $out; timeout 1500 "$@" 2>&1 | grep -E "fused|class|shapes|ERROR SUMMARY|RACECHECK SUMMARY|Hazard|hazard|Invalid|error" | head -40 >> $out; }
run compute-sanitizer --tool memcheck python tools/sanitize_scan.py 64 128 192 256
run compute-sanitizer --tool racecheck python tools/sanitize_scan.py 64 128
run env PFB_NO_PRUNE=1 compute-sanitizer --tool racecheck python tools/sanitize_scan.py 64 128
run compute-sanitizer --tool racecheck python tools/sanitize_scan.py 192
# per-axis fused pipeline: 4-lane pencils (32, 96), 8-lane (64, 128), with and without the TMEM stash
run compute-sanitizer --tool memcheck python tools/sanitize_scan.py 96x128x64b 32x64x96 128x32x96b 64x96x32 96x96x96b 32x32x32b
run compute-sanitizer --tool racecheck python tools/sanitize_scan.py 96x128x64b 32x64x96 64x96x128b
cat $out
