#!/usr/bin/env python
"""Golden vectors of the solution extraction, produced by the REAL reference
`Analyzer` (/root/reference/src/powerfit_em/analyzer.py, loaded by path -- it needs only numpy
and scipy) on LCC / rotation-index grids already stored in tests/golden/.

    python tests/golden/make_golden_analyzer.py        (build container only)

Stored per case: the feature-maximum positions the reference found and its `solutions` rows.
"""
import importlib.util
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "ref_analyzer", os.environ.get("REF_ANALYZER", "/root/reference/src/powerfit_em/analyzer.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def run(name, lcc, rot, rotations, steps, voxelspacing, origin, z_sigma):
    a = ref.Analyzer(lcc, rotations, rot, steps=steps, voxelspacing=voxelspacing, origin=origin, z_sigma=z_sigma)
    sol = np.array(a.solutions, dtype=np.float64)
    pos = np.array(sorted(a._positions), dtype=np.int64)
    with tempfile.TemporaryDirectory() as d:
        a.tofile(os.path.join(d, "solutions.out"))
        text = open(os.path.join(d, "solutions.out")).read()
    print(name, "features:", len(pos), "top cc %.4f" % sol[0, 0])
    return {name + "_positions": pos, name + "_solutions": sol, name + "_text": np.array(text),
            name + "_params": np.array([steps, voxelspacing, origin[0], origin[1], origin[2], z_sigma], dtype=np.float64)}


out = {}
for case, steps, vs, origin, zs in [("scan_config1_64", 5, 2.0, (0.0, 0.0, 0.0), 0.021),
                                    ("scan_32_plain", 5, 3.0, (-12.5, 4.0, 7.25), 1.0),
                                    ("scan_24_laplace_cw", 8, 1.0, (0, 0, 0), 0.1),
                                    ("scan_config2_128_subset", 5, 2.8, (10.0, -20.0, 30.0), 0.0085)]:
    g = np.load(os.path.join(HERE, case + ".npz"))
    out.update(run(case, g["lcc"], g["rot"], g["rotations"], steps, vs, origin, zs))
# a rough synthetic surface with plateaus and many shallow features
rng = np.random.default_rng(3)
from scipy.ndimage import gaussian_filter
rough = gaussian_filter(rng.normal(size=(40, 36, 44)), 2.0).astype(np.float32)
rough = (rough / np.abs(rough).max() * 0.9).astype(np.float32)
rough[rough < 0] = 0
rot = rng.integers(0, 50, size=rough.shape).astype(np.int32)
rots = rng.normal(size=(50, 3, 3))
out.update(run("rough", rough, rot, rots, 5, 1.5, (1.0, 2.0, 3.0), 0.05))
out["rough_lcc"] = rough
out["rough_rot"] = rot
out["rough_rotations"] = rots
np.savez_compressed(os.path.join(HERE, "analyzer_solutions.npz"), **out)
print("wrote analyzer_solutions.npz", os.path.getsize(os.path.join(HERE, "analyzer_solutions.npz")) >> 10, "KiB")
