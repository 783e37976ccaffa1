#!/bin/bash
# round-end evidence: smoke, full parity suite, all bench lines, ncu launch list + full capture of every fused kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_err.txt
for c in 1 3 4 5; do python bench.py --steps 3 --warmup 3 --workload config$c --no-cpu-baseline > gpurun_out/bench_config$c.json 2>> gpurun_out/bench_err.txt; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rot-per-step 512 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 30 -c 6 -o gpurun_out/fused_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rot-per-step 512 > gpurun_out/ncu_full.log 2>&1
for w in 4 5; do
ncu --set full --clock-control none --import-source on -k regex:cls_ -s 3 -c 3 -o gpurun_out/cls_full_config$w -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload config$w > gpurun_out/ncu_full$w.log 2>&1
done
ls -la gpurun_out | tail -12
