#!/bin/bash
# class-path kernel B at 256^3: 256 threads x 232 registers against 512 threads x 128 registers
mkdir -p gpurun_out
show='
import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], "rot/s %.0f  e2e %.0f  frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["step_frac"]), {k: round(v["us_per_rotation"],2) for k,v in d["roofline"]["kernels"].items()})'
PFB_CLS_BT=512 timeout 900 python -m pytest tests -m gpu -q -x -k "class_path" 2>&1 | tail -3
for bt in 256 512 256 512; do
PFB_CLS_BT=$bt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --workload config4 2>gpurun_out/err_config4.txt | python -c "$show" "config4 bt=$bt"
done
