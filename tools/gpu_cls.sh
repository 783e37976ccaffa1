#!/bin/bash
# class-decimated path (192^3, 256^3): its parity tests, then the 256^3 and 192^3 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "class_path" 2>&1 | tail -15
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $ARGS 2>gpurun_out/err.txt | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; tail -2 gpurun_out/err.txt; }
ARGS="--workload config4" run PFB_X=0
ARGS="--workload config5" run PFB_X=0
