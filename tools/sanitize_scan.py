"""compute-sanitizer target: one small search per fused pipeline (64^3, 128^3, 192^3 / 256^3 class path) plus the
device preparation and shape kernels.  Run as: compute-sanitizer --tool memcheck python tools/sanitize_scan.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from powerfit_b200 import CUDACorrelator, shapes, synth

# arguments: cube edges, or shapes as ZxYxX (per-axis fused pipeline); "-b" after a shape: binary mask
args = sys.argv[1:] or ["64", "128", "256"]
for a in args:
    if "x" in a:
        binary = a.endswith("b")
        shape = tuple(int(v) for v in a.rstrip("b").split("x"))
        small = min(shape) == 32
        case = synth.make_case(shape=shape, voxelspacing=3.0, resolution=9.0, n_res=40 if small else 120,
                               rg=5.0 if small else 10.0, n_copies=2, seed=3, core_weighted=not binary)
        c = CUDACorrelator(case.target, laplace=True, batch=4)
        c.template, c.mask, c.rotations = case.template, case.mask, synth.random_rotations(5, seed=2)
        c.scan()
        print(a, "fused", c.plan_info(6), "class", c.plan_info(9), "max lcc %.4f" % c.lcc.max(), flush=True)
sizes = [int(a) for a in args if "x" not in a]
for n in sizes:
    case = synth.make_case(n=n, voxelspacing=2.8, resolution=9.0, n_res=120, rg=11.0, n_copies=2, seed=3,
                           core_weighted=(n not in (128, 192)))      # 128 / 192: binary mask -> TMEM spectrum stash
    c = CUDACorrelator(case.target, laplace=True, batch=4)
    c.template, c.mask, c.rotations = case.template, case.mask, synth.random_rotations(5, seed=2)
    c.scan()
    print(n, "fused", c.plan_info(6), "class", c.plan_info(9), "max lcc %.4f" % c.lcc.max(), flush=True)
xyz = synth.random_walk_trace(40, 7.0, 1).T.copy()
grid = ((24, 28, 20), 2.5, (0.0, 0.0, 0.0))
t = shapes.structure_to_shape_like(grid, xyz, resolution=9.0, weights=np.full(40, 6.0), shape="vol")
m = shapes.determine_core_indices(shapes.structure_to_shape_like(grid, xyz, resolution=9.0, shape="mask"))
print("shapes", t.sum(), m.max(), flush=True)
