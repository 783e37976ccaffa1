"""Tiny fused-path scan for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import powerfit_b200
from powerfit_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
case = synth.make_case(n=n, voxelspacing=2.0, resolution=8.0, n_res=60, rg=9.0, n_copies=2, seed=3,
                       core_weighted=True)
c = powerfit_b200.CUDACorrelator(case.target, laplace=False, batch=4)
c.template, c.mask, c.rotations = case.template, case.mask, synth.random_rotations(5, seed=1)
c.scan()
print("fused", c.plan_info(6), "rs", c.plan_info(8), "max lcc", float(c.lcc.max()))
