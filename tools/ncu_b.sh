#!/bin/bash
# quick bench (no parity suite) + one ncu --set full capture of the kernels matching $1 (default: fused_fftyz)
mkdir -p gpurun_out
K=${1:-fused_fftyz}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/err.txt | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  step_frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']))
for k,v in d['roofline']['kernels'].items(): print('  %-20s %.3f ms/step  %.2f us/rot' % (k, v['ms_per_step'], 1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu']))"
tail -3 gpurun_out/err.txt
ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -o gpurun_out/ncu_$K -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
