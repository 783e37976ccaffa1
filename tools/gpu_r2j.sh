#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or fused or class or multi or shards or pencil" 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: rot/s %.0f  frac %.3f' % (d['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})"; }
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras $BENCH_ARGS 2>gpurun_out/err_$name.txt | tee gpurun_out/bench_$name.json | summ $name; tail -1 gpurun_out/err_$name.txt; }
run default1 PFB_NOP=1
run default2 PFB_NOP=1
BENCH_ARGS="--workload config3" run cw PFB_NOP=1
BENCH_ARGS="--workload config1" run c1 PFB_NOP=1
