#!/bin/bash
# per-axis fused pipeline: pencil blocks, mixed-shape parity (generic path + reference goldens), padding,
# then the whole parity suite and the headline bench line (regression check of the cubes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pencil or mixed or padding or target_prep" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/err.txt | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})"; tail -3 gpurun_out/err.txt
