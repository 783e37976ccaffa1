#!/bin/bash
# kernel B3 (plane in registers): parity at 128^3 with PFB_B3=1, then the bench line with and without it
mkdir -p gpurun_out
PFB_B3=1 timeout 900 python -m pytest tests -m gpu -q -x -k "128 or fused_path or shards" 2>&1 | tail -8
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $ARGS 2>gpurun_out/err.txt | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; tail -2 gpurun_out/err.txt; }
ARGS="" run PFB_B3=1
ARGS="" run PFB_B3=1 PFB_B3_STAGE=0
ARGS="--workload config3" run PFB_B3=1
