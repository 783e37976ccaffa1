#!/bin/bash
# one GPU box visit: parity suite, headline bench, ncu launch list, ncu --set full of the fused kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_err.txt
cat gpurun_out/bench_config2.json
python bench.py --steps 5 --warmup 3 --workload config3 --no-cpu-baseline > gpurun_out/bench_config3.json 2>> gpurun_out/bench_err.txt
python bench.py --steps 5 --warmup 3 --workload config1 --no-cpu-baseline > gpurun_out/bench_config1.json 2>> gpurun_out/bench_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 30 -c 6 -o gpurun_out/fused_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
