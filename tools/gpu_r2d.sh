#!/bin/bash
# parity suite + bench lines of all configs with the kernel variants of this round
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})"; }
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras $BENCH_ARGS 2>gpurun_out/err_$name.txt | tee gpurun_out/bench_$name.json | summ $name; tail -2 gpurun_out/err_$name.txt; }
run default PFB_NOP=1
run nostage PFB_B_STAGE=0
BENCH_ARGS="--workload config3" run cw PFB_NOP=1
BENCH_ARGS="--workload config3" run cw_nostage PFB_B_STAGE=0
BENCH_ARGS="--workload config1" run c1 PFB_NOP=1
BENCH_ARGS="--workload config1" run c1_nostage PFB_B_STAGE=0
BENCH_ARGS="--workload config5" run c5 PFB_NOP=1
BENCH_ARGS="--workload config5" run c5_oldc PFB_CLS_C_TMA=0
BENCH_ARGS="--workload config4" run c4 PFB_NOP=1
BENCH_ARGS="--workload config4" run c4_oldc PFB_CLS_C_TMA=0
