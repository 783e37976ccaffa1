#!/bin/bash
# class-path kernel A: pencils of rows outside the sphere skipped
mkdir -p gpurun_out
show='
import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], "rot/s %.0f  e2e %.0f  frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["step_frac"]), {k: round(v["us_per_rotation"],2) for k,v in d["roofline"]["kernels"].items()})'
timeout 900 python -m pytest tests -m gpu -q -x -k "class_path" 2>&1 | tail -3
for w in config4 config5 config4 config5; do
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras --workload $w 2>gpurun_out/err_$w.txt | python -c "$show" "$w"
done
