#!/usr/bin/env python
"""Golden vectors for the image-pyramid row, produced by the REAL reference functions
powerfit_em.volume.lower_resolution and powerfit_em.volume.resample (what scripts/__init__.py:93-103 calls).

    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden_pyramid.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))

from powerfit_em.volume import Volume, lower_resolution, resample   # noqa: E402  (reference)
from powerfit_b200 import synth                                       # noqa: E402

case = synth.make_case(shape=(24, 30, 20), voxelspacing=2.0, resolution=8.0, n_res=40, rg=7.0, n_copies=2, seed=61)
vol = Volume(case.target.astype(np.float32).astype(np.float64), 2.0, (0, 0, 0))
out = {"map": vol.array, "voxelspacing": np.array(2.0), "resolution": np.array(8.0)}
targets = [12.0, 20.0, 30.0]
for i, res in enumerate(targets):
    low = lower_resolution(vol, 8.0, res)
    new_vs = res / (2 * 2)
    factor = vol.voxelspacing / new_vs
    rs = resample(low, factor, order=1)
    out["low_%d" % i] = low.array
    out["res_%d" % i] = rs.array
    out["vs_%d" % i] = np.array(rs.voxelspacing)
    print(res, low.array.shape, rs.array.shape, rs.voxelspacing)
out["targets"] = np.array(targets)
np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **out)
