#!/bin/bash
# 2-GPU check: the driver's launch line for both arms (ours, reference) at N=2, and the reference arm at N=1
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/ref_err.txt
cat gpurun_out/bench_reference_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/n2_err.txt
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/n2_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_reference_n2.json 2>> gpurun_out/ref_err.txt
cat gpurun_out/bench_reference_n2.json; tail -5 gpurun_out/ref_err.txt
