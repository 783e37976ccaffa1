#!/usr/bin/env python
"""Golden files of the map IO, written and read back by the REAL reference
(`powerfit_em.volume`: Volume.tofile / to_mrc, parse_volume) in the build container:

    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden_volume_io.py

Stored in tests/golden/volume_io.npz: the bytes of every file the reference wrote (float64 / float32 / int16 / int8
volumes as .mrc, .ccp4 and .map), and what the reference's own parser returns for each of them (density,
voxel spacing, origin).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))
from powerfit_em.volume import Volume, parse_volume          # noqa: E402  (reference)

rng = np.random.default_rng(17)
cases = {
    "f64_mrc": (rng.normal(size=(6, 9, 11)), 1.25, (3.75, -7.5, 12.5), "mrc"),
    "f32_ccp4": (rng.normal(size=(5, 4, 7)).astype(np.float32), 2.0, (4.0, -6.0, 10.0), "ccp4"),
    "f64_map": (rng.random((4, 6, 5)) * 10, 3.0, (-9.0, 3.0, 0.0), "map"),
    "i16_mrc": (rng.integers(-300, 300, size=(3, 5, 8)).astype(np.int16), 1.5, (0.0, 1.5, -3.0), "mrc"),
    "i8_ccp4": (rng.integers(-100, 100, size=(4, 4, 6)).astype(np.int8), 1.0, (2.0, 3.0, -4.0), "ccp4"),
}
out = {"names": np.array(sorted(cases))}
with tempfile.TemporaryDirectory() as d:
    for name, (arr, vs, origin, ext) in cases.items():
        path = os.path.join(d, name + "." + ext)
        Volume(arr, vs, origin).tofile(path)
        raw = open(path, "rb").read()
        dens, pvs, porigin = parse_volume(path)
        out[name + "_array"] = arr
        out[name + "_meta"] = np.array([vs, origin[0], origin[1], origin[2]])
        out[name + "_ext"] = np.array(ext)
        out[name + "_bytes"] = np.frombuffer(raw, dtype=np.uint8)
        out[name + "_density"] = dens
        out[name + "_voxelspacing"] = np.array(pvs)
        out[name + "_origin"] = np.asarray(porigin, dtype=np.float64)
        print(name, ext, len(raw), "bytes; parsed", dens.dtype, dens.shape, pvs, list(porigin))
np.savez_compressed(os.path.join(HERE, "volume_io.npz"), **out)
