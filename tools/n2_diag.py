"""Where does a 2-GPU bench step spend its time? scan vs all-reduce, timed with CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from powerfit_b200 import CUDACorrelator, synth
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
case = synth.config2(seed=0)
rots = synth.random_rotations(512 * 8 * world + 8, seed=1)
corr = CUDACorrelator(case.target, device=dev, laplace=True)
corr.template, corr.mask, corr.rotations = case.template, case.mask, rots
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for i in range(8):
    lo = (i * world + rank) * 512
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record()
    best = corr.scan_device(lo, lo + 512, reset=(i == 0))
    ev[1].record()
    t1 = time.perf_counter()
    dist.all_reduce(best, op=dist.ReduceOp.MAX)
    ev[2].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("rank %d step %d: scan %.2f ms, allreduce %.2f ms (device) | host: launch scan %.2f, launch ar %.2f, total %.2f ms; best ptr %x"
          % (rank, i, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3, best.data_ptr()), flush=True)
dist.destroy_process_group()
