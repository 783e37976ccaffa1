#!/usr/bin/env python
"""Full-search golden of BASELINE configs[1]: 128^3, Laplace, the WHOLE 10 degree rotation set
(7416 rotations), produced by the REAL reference (`powerfit_em.powerfitter.CPUCorrelator` for the
scan, `powerfit_em.analyzer.Analyzer` for solutions.out) in the build container:

    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden_full128.py [nproc]

The rotation list is split into contiguous blocks, one reference CPUCorrelator per process, and
the blocks are merged in order with the reference's own rule (powerfitter.py:95-108, 146-163:
strict '>' of a later block against the merged grid), additionally keeping the runner-up LCC so
that the test can tell where the arg-max is decided by less than the tolerance.

Stored (the full grids would be 25 MB): the rotation set, lcc / rot / lcc2 on eight z planes,
the global maximum and its position, and the reference Analyzer's solutions (all rows + the
solutions.out text of the top 100) computed on the FULL reference grids.
"""
import multiprocessing as mp
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))
warnings.simplefilter("ignore")

from powerfit_em.powerfitter import CPUCorrelator            # noqa: E402  (reference)
from powerfit_em.analyzer import Analyzer                    # noqa: E402  (reference)
from powerfit_em.rotations import proportional_orientations, quat_to_rotmat  # noqa: E402

from powerfit_b200 import synth                               # noqa: E402


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def block_scan(job):
    target, template, mask, rotations, laplace = job
    c = CPUCorrelator(target, laplace=laplace)
    c.template = template
    c.mask = mask
    c.rotations = rotations
    lcc = np.zeros(target.shape)
    rot = np.zeros(target.shape)
    lcc2 = np.full(target.shape, -np.inf)
    for n in range(c._rotations.shape[0]):
        c._translational_scan(c._rotations[n])
        scan = c._lcc_scan
        ind = scan > lcc
        cand = np.where(ind, lcc, scan)
        np.fmax(lcc2, cand, out=lcc2)
        lcc[ind] = scan[ind]
        rot[ind] = n
    return lcc, rot, lcc2, float(c._norm_factor), c._lcc_mask.copy()


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    laplace = True
    q, w, a = proportional_orientations(10.0)
    rotations = quat_to_rotmat(q)
    R = len(rotations)
    case = synth.config2(seed=0)
    target, template, mask = f32(case.target), f32(case.template), f32(case.mask)
    per = R // nproc
    bounds = [(i * per, (i + 1) * per if i + 1 < nproc else R) for i in range(nproc)]
    t0 = time.time()
    with mp.get_context("fork").Pool(nproc) as pool:
        parts = pool.map(block_scan, [(target, template, mask, rotations[lo:hi], laplace) for lo, hi in bounds])
    dt = time.time() - t0
    print("reference scan: %d rotations, %d processes, %.1f s (%.2f rot/s)" % (R, nproc, dt, R / dt))
    lcc = np.zeros(target.shape)
    rot = np.zeros(target.shape)
    lcc2 = np.full(target.shape, -np.inf)
    for (lo, hi), (l, r, l2, norm, lmask) in zip(bounds, parts):
        ind = l > lcc                                   # powerfitter.py:157-159
        loser = np.where(ind, lcc, l)
        np.fmax(lcc2, loser, out=lcc2)
        np.fmax(lcc2, l2, out=lcc2)
        lcc[ind] = l[ind]
        rot[ind] = r[ind] + lo
    an = Analyzer(lcc, rotations, rot, voxelspacing=case.voxelspacing, origin=(10.0, -20.0, 30.0),
                  z_sigma=1.0 / np.sqrt(lcc.size))      # default steps=5, as powerfit.py:299-303
    sol = np.array(an.solutions, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        an.tofile(os.path.join(d, "solutions.out"))
        text = open(os.path.join(d, "solutions.out")).read()
    top_text = "".join(text.splitlines(keepends=True)[:101])
    planes = np.array([0, 1, 37, 64, 65, 90, 126, 127])
    d = dict(rotations=rotations, laplace=np.array(1), seed=np.array(0), planes=planes,
             lcc=lcc[planes].astype(np.float32), rot=rot[planes].astype(np.int32),
             lcc2=lcc2[planes].astype(np.float32), lcc64_max=np.array(lcc.max()),
             argmax=np.array(np.unravel_index(np.argmax(lcc), lcc.shape)),
             norm_factor=np.array(norm), lcc_mask=np.packbits(lmask[planes].astype(bool)),
             nonzero=np.array(int((lcc > 0).sum())), lcc_sum=np.array(lcc.sum()),
             solutions=sol, solutions_text=np.array(top_text),
             analyzer_params=np.array([5, case.voxelspacing, 10.0, -20.0, 30.0, 1.0 / np.sqrt(lcc.size)]),
             ref_seconds=np.array(dt), ref_nproc=np.array(nproc))
    np.savez_compressed(os.path.join(HERE, "scan_config2_128_full.npz"), **d)
    print("scan_config2_128_full: lcc max %.5f at %s, %d solutions, top:" % (lcc.max(), d["argmax"], len(sol)))
    print(top_text[:600])


if __name__ == "__main__":
    main()
