"""One-off sweep of the per-axis fused pipeline against the any-shape pipeline: shapes outside the test list, templates
from small to box-filling (support radius clamped to rmax -> unstaged kernel B), binary and core-weighted masks, odd
rotation counts, small batches.   python tools/mixed_sweep.py"""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

CASES = [((32, 128, 64), 14.0, False, True), ((96, 64, 96), 9.0, True, False), ((128, 96, 32), 13.0, False, False),
         ((64, 32, 32), 4.0, True, True), ((96, 96, 128), 30.0, False, True), ((64, 128, 96), 22.0, True, True),
         ((32, 32, 96), 12.0, False, False), ((128, 128, 32), 6.0, True, False), ((96, 128, 128), 16.0, False, True)]

def run(mode, idx):
    from powerfit_b200 import CUDACorrelator, synth
    shape, rg, cw, laplace = CASES[idx]
    case = synth.make_case(shape=shape, voxelspacing=3.0, resolution=9.0, n_res=int(20 + rg * 10), rg=rg, n_copies=2,
                           seed=100 + idx, core_weighted=cw)
    rots = synth.random_rotations(5 + idx % 3, seed=idx)
    c = CUDACorrelator(case.target, laplace=laplace, batch=2 + 2 * (idx % 3))
    c.template, c.mask, c.rotations = case.template, case.mask, rots
    c.scan()
    np.savez("/tmp/sweep_%s_%d.npz" % (mode, idx), lcc=c.lcc, rot=c.rot, fused=c.plan_info(6), rs=c.plan_info(8))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]))
        sys.exit(0)
    bad = 0
    for i, (shape, rg, cw, laplace) in enumerate(CASES):
        for mode, env in (("generic", {"PFB_FUSED": "0"}), ("fused", {"PFB_FUSED": "1"})):
            e = dict(os.environ); e.update(env)
            subprocess.run([sys.executable, __file__, mode, str(i)], env=e, check=True)
        g, f = np.load("/tmp/sweep_generic_%d.npz" % i), np.load("/tmp/sweep_fused_%d.npz" % i)
        d = float(np.abs(g["lcc"] - f["lcc"]).max())
        same = float((g["rot"] == f["rot"]).mean())
        ok = int(f["fused"]) == 1 and int(g["fused"]) == 0 and d < 2e-5 and same > 0.999
        bad += not ok
        print(shape, "rg", rg, "cw", cw, "laplace", laplace, "rs", int(f["rs"]), "rmax", min(shape) // 2,
              "max|dLCC| %.2e" % d, "rot equal %.5f" % same, "nonzero", int((g["lcc"] > 0).sum()), "OK" if ok else "FAIL")
    print("failures:", bad)
    sys.exit(1 if bad else 0)
