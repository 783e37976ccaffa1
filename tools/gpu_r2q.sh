#!/bin/bash
# throughput of the per-axis fused pipeline on a CLI-like non-cubic grid
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --workload cli_96x128x64 2>gpurun_out/err_cli.txt | tee gpurun_out/bench_cli_96x128x64.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})"; tail -3 gpurun_out/err_cli.txt
PFB_FUSED=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --workload cli_96x128x64 --rot-per-step 512 2>>gpurun_out/err_cli.txt | tee gpurun_out/bench_cli_96x128x64_generic.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('generic: rot/s %.0f  e2e %.0f  frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']))"
