#!/bin/bash
# N GPUs (default 2): the 2-GPU NCCL parity test, then the driver's torchrun bench line (merge_check + strong inside)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests -m gpu -q -x -k "two_gpu" 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/n${N}_err.txt
tail -3 gpurun_out/n${N}_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N=%d rot/s %.0f e2e %.0f frac %.3f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['step_frac']))
print('merge_check', d.get('merge_check'))
print('strong', d.get('strong'))
print('multi', d.get('multi_template'))
print('config4_sharded', d.get('config4_sharded'))
PY
