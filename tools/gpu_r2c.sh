#!/bin/bash
# parity suite, full default bench line, ncu launch list + full capture of the fused kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_err.txt; tail -3 gpurun_out/bench_err.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_config2.json'))
print('rot/s %.0f e2e %.0f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']))
print({k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})
print('cpu', d.get('cpu_baseline'))
print('strong', d.get('strong'))
for k,v in d.get('configs',{}).items(): print(k, v.get('value'), v.get('step_frac'), v.get('us_per_rotation'), v.get('error'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --rot-per-step 512 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 3 -o gpurun_out/fused_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --rot-per-step 512 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -5
