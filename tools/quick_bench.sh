#!/bin/bash
# quick GPU check: parity suite + headline bench + per-kernel split
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/err.txt | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  step_frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']))
for k,v in d['roofline']['kernels'].items(): print('  %-20s %.3f ms/step  %.2f us/rot' % (k, v['ms_per_step'], 1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu']))"
tail -3 gpurun_out/err.txt
