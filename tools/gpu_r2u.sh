#!/bin/bash
# class-path kernel B: TMA-staged rows (256^3) on / off
mkdir -p gpurun_out
show='
import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], "rot/s %.0f  e2e %.0f  frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["step_frac"]), {k: round(v["us_per_rotation"],2) for k,v in d["roofline"]["kernels"].items()})'
timeout 900 python -m pytest tests -m gpu -q -x -k "class_path" 2>&1 | tail -3
for st in 0 1 0 1; do
PFB_B_STAGE=$st timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --workload config4 2>gpurun_out/err_config4.txt | python -c "$show" "config4 stage=$st"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --workload config5 2>gpurun_out/err_config5.txt | python -c "$show" "config5"
