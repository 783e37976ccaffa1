#!/bin/bash
# round 2 experiment: parity suite, then the headline bench under kernel-C / overlap variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; }
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras $BENCH_ARGS 2>gpurun_out/err_$name.txt | tee gpurun_out/bench_$name.json | summ $name; tail -2 gpurun_out/err_$name.txt; }
run default PFB_NOP=1
run c_tma0 PFB_C_TMA=0
run c_tma1 PFB_C_TMA=1
run c_tma3 PFB_C_TMA=3
run no_overlap PFB_OVERLAP=0
BENCH_ARGS="--workload config3" run cw PFB_NOP=1
BENCH_ARGS="--workload config1" run c1 PFB_NOP=1
