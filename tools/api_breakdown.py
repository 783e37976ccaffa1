"""Where the wall-clock of one user-level search goes (CUDACorrelator from host arrays), per preparation mode."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from powerfit_b200 import CUDACorrelator, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
case = synth.config2(seed=0) if n == 128 else synth.config4(seed=0)
rots = synth.random_rotations(7416 if n == 128 else 256, seed=1)
def sync(): torch.cuda.synchronize()
for rep in range(3):
    for mode in ("device", "host"):
        sync(); t = [time.perf_counter()]
        c = CUDACorrelator(case.target, laplace=True, prep=mode); sync(); t.append(time.perf_counter())
        c.template = case.template; sync(); t.append(time.perf_counter())
        c.mask = case.mask; sync(); t.append(time.perf_counter())
        c.rotations = rots; c.shard = False
        c.scan(); sync(); t.append(time.perf_counter())
        del c; sync(); t.append(time.perf_counter())
        d = np.diff(t)
        print("%d %-6s ctor %.3f  template %.3f  mask %.3f  scan %.3f  del %.3f  total %.3f" % ((rep, mode) + tuple(d) + (t[-1] - t[0],)))
