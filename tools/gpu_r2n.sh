#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "class or multi" 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: rot/s %.0f  frac %.3f' % (d['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})"; }
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras $BENCH_ARGS 2>gpurun_out/err_$name.txt | tee gpurun_out/bench_$name.json | summ $name; tail -1 gpurun_out/err_$name.txt; }
BENCH_ARGS="--workload config5" run c5 PFB_NOP=1
BENCH_ARGS="--workload config4" run c4 PFB_NOP=1
BENCH_ARGS="--workload config5" run c5b PFB_NOP=1
BENCH_ARGS="--workload config4" run c4b PFB_NOP=1
