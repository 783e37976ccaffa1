"""2-GPU diagnostic: topology, peer access, NCCL transport and all-reduce timing for the packed best grid."""
import os, sys, time, subprocess
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print("can_access_peer 0->1:", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)
    print("df /dev/shm:", subprocess.run(["df", "-h", "/dev/shm"], capture_output=True, text=True).stdout)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for dtype in (torch.int64, torch.float32):
    x = torch.zeros(2 * 1024 * 1024, dtype=dtype, device="cuda") + rank
    for op in (dist.ReduceOp.MAX, dist.ReduceOp.SUM):
        for _ in range(3):
            dist.all_reduce(x, op=op)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dist.all_reduce(x, op=op)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        if rank == 0:
            print("all_reduce %s %s %d MB: %.3f ms" % (dtype, op, x.numel() * x.element_size() >> 20, dt * 1e3))
dist.destroy_process_group()
