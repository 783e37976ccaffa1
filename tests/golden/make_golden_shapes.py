#!/usr/bin/env python
"""Golden vectors for template / mask synthesis, produced by the REAL reference:
powerfit_em.volume.structure_to_shape_like (over _powerfit.blur_points / dilate_points) and
powerfit_em.helpers.determine_core_indices.  Run in the build container only:

    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden_shapes.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))

from powerfit_em.volume import Volume, structure_to_shape_like   # noqa: E402  (reference)
from powerfit_em.helpers import determine_core_indices           # noqa: E402  (reference)
from powerfit_b200 import synth                                   # noqa: E402

out = {}
cases = [
    # name, shape, voxelspacing, origin, n_res, rg, resolution, seed
    ("small", (20, 24, 28), 2.5, (3.0, -4.0, 7.5), 25, 6.0, 9.0, 1),
    ("tiny_box", (6, 7, 8), 3.0, (0.0, 0.0, 0.0), 5, 3.0, 12.0, 2),       # the 4 sigma box exceeds the grid
    ("config1", (64, 64, 64), 2.0, (10.0, 20.0, 30.0), 300, 14.0, 8.0, 0),
]
for name, shape, vs, origin, n_res, rg, res, seed in cases:
    rng = np.random.default_rng(seed)
    xyz = synth.random_walk_trace(n_res, rg, seed).T.copy()      # (3, n) centred on 0
    xyz += np.asarray(origin).reshape(3, 1)                      # the CLI moves the model to the map origin
    weights = rng.integers(6, 17, size=n_res).astype(np.float64)  # atomic numbers C..S
    vol = Volume(np.zeros(shape), vs, origin)
    t = structure_to_shape_like(vol, xyz.copy(), resolution=res, weights=weights, shape="vol").array
    m = structure_to_shape_like(vol, xyz.copy(), resolution=res, shape="mask").array
    radii = rng.uniform(3.0, 7.0, size=n_res)
    m2 = structure_to_shape_like(vol, xyz.copy(), resolution=res, radii=radii.copy(), shape="mask").array
    core = determine_core_indices(m)
    out.update({name + "_shape": np.array(shape), name + "_vs": np.array(vs), name + "_origin": np.array(origin),
                name + "_res": np.array(res), name + "_xyz": xyz, name + "_weights": weights, name + "_radii": radii,
                name + "_vol": t, name + "_mask": m.astype(np.uint8), name + "_mask_radii": m2.astype(np.uint8),
                name + "_core": core.astype(np.uint8)})
    print(name, shape, "vol max %.4f sum %.3f" % (t.max(), t.sum()), "mask", int(m.sum()), int(m2.sum()), "core max", core.max())
out["names"] = np.array([c[0] for c in cases])
np.savez_compressed(os.path.join(HERE, "shapes.npz"), **out)
