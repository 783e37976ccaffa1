#!/bin/bash
# round-end evidence on one GPU: smoke, full parity suite, the default bench line (all extras), the other configs'
# own lines, ncu launch list + full capture of the fused kernels (128^3) and of the class-path kernels (192^3 / 256^3)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_err.txt; tail -2 gpurun_out/bench_err.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n1.json 2>> gpurun_out/bench_err.txt
for c in config1 config3 config4 config5 cli_96x128x64; do timeout 600 python bench.py --steps 3 --warmup 3 --workload $c --no-cpu-baseline --no-extras > gpurun_out/bench_$c.json 2>> gpurun_out/bench_err.txt; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_config2.json'))
print('rot/s %.0f e2e %.0f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['step_frac']), {k: round(v['us_per_rotation'],2) for k,v in d['roofline']['kernels'].items()})
print('cpu', d.get('cpu_baseline',{}).get('value'), 'strong', d['strong']['seconds'], d['strong']['rotations_per_s'])
print('multi', d.get('multi_template'))
for c in ('config1','config3','config4','config5','cli_96x128x64'):
    x=json.load(open('gpurun_out/bench_%s.json'%c)); print(c, round(x['value']), round(x['e2e']['value']), round(x['roofline']['step_frac'],3), {k: round(v['us_per_rotation'],2) for k,v in x['roofline']['kernels'].items()})
print('reference arm', json.load(open('gpurun_out/bench_reference_n1.json'))['value'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --rot-per-step 1024 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 3 -o gpurun_out/fused_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --rot-per-step 1024 > gpurun_out/ncu_full.log 2>&1
for w in 4 5; do
ncu --set full --clock-control none --import-source on -k regex:cls_ -s 3 -c 3 -o gpurun_out/cls_full_config$w -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --workload config$w > gpurun_out/ncu_full$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
