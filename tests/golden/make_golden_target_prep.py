#!/usr/bin/env python
"""Golden vectors for the CLI's target preparation (powerfit.py:219-233), produced by the REAL
reference functions resample / trim / extend / nearest_multiple2357.

    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden_target_prep.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))

from powerfit_em.volume import Volume, resample, trim, extend, nearest_multiple2357   # noqa: E402  (reference)
from powerfit_b200 import synth                                                          # noqa: E402

case = synth.make_case(shape=(44, 52, 38), voxelspacing=1.2, resolution=9.0, n_res=60, rg=8.0, n_copies=2, seed=71,
                       noise=0.01)
target = Volume(case.target.astype(np.float32).astype(np.float64), 1.2, (5.0, -3.0, 11.0))
resolution, rate = 9.0, 2
out = {"map": target.array, "voxelspacing": np.array(1.2), "origin": np.array(target.origin), "resolution": np.array(resolution)}
factor = 2 * rate * target.voxelspacing / resolution
assert factor < 0.9
t = resample(target, factor)
out["resampled"] = t.array
t = trim(t, t.array.max() / 10)
out["trimmed"] = t.array
out["trim_origin"] = np.array(t.origin)
shape = [nearest_multiple2357(n) for n in t.shape]
t = extend(t, shape)
out["final"] = t.array
out["final_vs"] = np.array(t.voxelspacing)
out["smooth"] = np.array([nearest_multiple2357(n) for n in range(1, 300)])
print(target.shape, "->", out["resampled"].shape, "->", out["trimmed"].shape, "->", out["final"].shape, t.voxelspacing, t.origin)
np.savez_compressed(os.path.join(HERE, "target_prep.npz"), **out)
