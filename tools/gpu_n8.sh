#!/bin/bash
# the driver's N-GPU launch line for our arm (N from $1, default 8)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/n${N}_err.txt
cat gpurun_out/bench_n$N.json | cut -c1-900; tail -3 gpurun_out/n${N}_err.txt
