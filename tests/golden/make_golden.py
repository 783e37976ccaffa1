#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by running the REAL reference.

Run in the build container only (the GPU box has no /root/reference):

    cp -r /root/reference /tmp/ref_build && (cd /tmp/ref_build && python setup.py build_ext --inplace)
    REF_SRC=/tmp/ref_build/src python tests/golden/make_golden.py

Everything written here is an input/output pair of the reference's own functions
(`powerfit_em._extensions.rotate_grid3d`, `powerfit_em.powerfitter.CPUCorrelator`);
no reference source is copied.  Inputs are rounded to float32 before they are fed
to the reference so the stored float32 copies reproduce them exactly.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("REF_SRC", "/tmp/ref_build/src"))
warnings.simplefilter("ignore")

from powerfit_em._extensions import rotate_grid3d            # noqa: E402  (reference)
from powerfit_em.powerfitter import CPUCorrelator            # noqa: E402  (reference)
from powerfit_em.rotations import proportional_orientations, quat_to_rotmat  # noqa: E402
from powerfit_em.helpers import determine_core_indices       # noqa: E402

from powerfit_b200 import synth                               # noqa: E402


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def rotation_set(angle):
    q, w, a = proportional_orientations(angle)
    return quat_to_rotmat(q)


# --------------------------------------------------------------------------- #
def rotate_vectors():
    rng = np.random.default_rng(11)
    cases = []
    # the reference's own unit-test inputs (tests/test_extensions.py:10-37)
    grid = np.zeros((4, 5, 6))
    for idx in [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 0, 2), (0, 0, -1), (-1, 0, 0)]:
        grid[idx] = 1
    eye = np.eye(3)
    rz90 = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    cases.append((grid, eye, 2, 1))
    cases.append((grid, rz90, 2, 0))
    rots = synth.random_rotations(6, seed=5)
    for shape, radius in [((8, 8, 8), 4), ((6, 7, 9), 3), ((12, 10, 16), 5), ((16, 16, 16), 8),
                          ((9, 9, 9), 4), ((5, 6, 7), 3)]:
        g = f32(rng.normal(size=shape))
        for R in rots[:4]:
            for nearest in (0, 1):
                cases.append((g, R, radius, nearest))
    # half-integer coordinates (exact ties of round()) : 60 deg about z has cos = 0.5
    c, s = 0.5, np.around(np.sqrt(3) / 2, 8)
    Rz60 = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    g = f32(rng.normal(size=(10, 10, 10)))
    cases.append((g, Rz60, 5, 1))
    cases.append((g, Rz60, 5, 0))
    out = {}
    for i, (g, R, radius, nearest) in enumerate(cases):
        o = np.zeros_like(g)
        rotate_grid3d(g, np.ascontiguousarray(R, dtype=np.float64), radius, o, nearest)
        out["grid_%d" % i] = g
        out["rotmat_%d" % i] = R
        out["meta_%d" % i] = np.array([radius, nearest])
        out["out_%d" % i] = o
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "rotate_vectors.npz"), **out)
    print("rotate_vectors:", len(cases), "cases")


def run_reference_scan(target, template, mask, rotations, laplace):
    """Drive the reference CPUCorrelator rotation by rotation, keeping the runner-up."""
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
    # and once through the reference's own scan() to be sure the harness loop is it
    c.scan()
    assert np.array_equal(np.nan_to_num(c.lcc), np.nan_to_num(lcc)) and np.array_equal(c.rot, rot)
    return lcc, rot, lcc2, c


def save_scan(name, case, rotations, laplace, store_inputs=True, extra=None):
    target, template, mask = f32(case.target), f32(case.template), f32(case.mask)
    lcc, rot, lcc2, c = run_reference_scan(target, template, mask, rotations, laplace)
    d = dict(rotations=rotations, laplace=np.array(int(laplace)),
             lcc=lcc.astype(np.float32), lcc64_max=np.array(lcc.max()),
             rot=rot.astype(np.int32), lcc2=lcc2.astype(np.float32),
             norm_factor=np.array(float(c._norm_factor)), rmax=np.array(int(c._rmax)),
             prepped_template=c._template.astype(np.float32),
             lcc_mask=np.packbits(c._lcc_mask.astype(bool)))
    if store_inputs:
        d.update(target=target.astype(np.float32), template=template.astype(np.float32),
                 mask=mask.astype(np.float32))
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, target.shape, "R=%d" % len(rotations), "lcc max %.4f" % lcc.max(),
          "nonzero %d" % (lcc > 0).sum())


def scans():
    # (a) non-cubic 2.3.5.7-smooth shape, plain LCC, 24 rotations
    case = synth.make_case(shape=(16, 18, 20), voxelspacing=3.0, resolution=9.0, n_res=40, rg=7.0,
                           n_copies=2, seed=3)
    save_scan("scan_16x18x20_plain", case, rotation_set(62.8), False)
    # (b) radix 3/5/7 axes, odd length, laplace
    case = synth.make_case(shape=(12, 15, 14), voxelspacing=3.0, resolution=9.0, n_res=20, rg=5.0,
                           n_copies=1, seed=4)
    save_scan("scan_12x15x14_laplace", case, rotation_set(62.8), True)
    # (c) 24^3 laplace + core-weighted, 60 rotations
    case = synth.make_case(n=24, voxelspacing=3.0, resolution=9.0, n_res=60, rg=8.0, n_copies=2,
                           seed=5, core_weighted=True)
    save_scan("scan_24_laplace_cw", case, rotation_set(44.48), True)
    # (d) 32^3 plain, 216 rotations
    case = synth.make_case(n=32, voxelspacing=3.0, resolution=9.0, n_res=80, rg=9.0, n_copies=2, seed=6)
    save_scan("scan_32_plain", case, rotation_set(36.47), False)
    # (e) BASELINE config 1: 64^3, 8 A, 300 residues, full 20 deg set (648 rotations)
    case = synth.config1(seed=0)
    save_scan("scan_config1_64", case, rotation_set(20.0), False)
    # (f) BASELINE config 2 (subset): 128^3, laplace, 24 rotations strided through the
    #     10 deg set (first and last included); indices refer to the subset order.
    full = rotation_set(10.0)
    idx = np.unique(np.r_[np.arange(0, len(full), len(full) // 22)[:23], len(full) - 1])
    case = synth.config2(seed=0)
    save_scan("scan_config2_128_subset", case, full[idx], True, store_inputs=False,
              extra=dict(subset_index=idx, seed=np.array(0)))
    # (g) BASELINE config 3 (subset): 128^3, core-weighted, 12 rotations
    idx = np.unique(np.r_[np.arange(0, len(full), len(full) // 11)[:11], len(full) - 1])
    case = synth.config2(seed=0, core_weighted=True)
    save_scan("scan_config3_128_cw_subset", case, full[idx], False, store_inputs=False,
              extra=dict(subset_index=idx, seed=np.array(0)))


def scans_mixed():
    """CLI-like non-cubic grids whose axes are fused lengths (32 / 64 / 96 / 128): they take the per-axis fused
    pipeline.  Inputs are regenerated from the seeded generator (the keyword arguments travel in the file)."""
    import json
    full = rotation_set(20.0)
    for name, kw, laplace, nrot in [
            ("scan_96x128x64_laplace", dict(shape=(96, 128, 64), voxelspacing=3.0, resolution=9.0, n_res=150, rg=11.0,
                                            n_copies=3, seed=31), True, 14),
            ("scan_32x64x96_cw", dict(shape=(32, 64, 96), voxelspacing=3.0, resolution=9.0, n_res=50, rg=5.5,
                                      n_copies=3, seed=32, core_weighted=True), False, 15)]:
        idx = np.unique(np.r_[np.arange(0, len(full), len(full) // (nrot - 1))[:nrot - 1], len(full) - 1])
        case = synth.make_case(**kw)
        save_scan(name, case, full[idx], laplace, store_inputs=False,
                  extra=dict(subset_index=idx, case_kwargs=np.array(json.dumps(kw))))


def scan_256():
    """BASELINE config 4 (subset): 256^3, Laplace + core-weighted, 6 rotations of the 4.71 deg set + 2 true poses.
    Only eight z planes of the result grids are stored (the full grids would be 200 MB)."""
    full = rotation_set(4.71)
    idx = np.unique(np.r_[np.arange(0, len(full), len(full) // 5)[:5], len(full) - 1])
    case = synth.config4(seed=0)
    target, template, mask = f32(case.target), f32(case.template), f32(case.mask)
    # plus the true orientations of two of the copies in the map, so that real peaks are in the grids
    rotations = np.concatenate([full[idx[:3]], case.poses[0][0][None], full[idx[3:]], case.poses[4][0][None]])
    lcc, rot, lcc2, c = run_reference_scan(target, template, mask, rotations, True)
    planes = np.array([0, 1, 77, 128, 130, 201, 254, 255])
    d = dict(rotations=rotations, laplace=np.array(1), planes=planes, seed=np.array(0), subset_index=idx,
             lcc=lcc[planes].astype(np.float32), rot=rot[planes].astype(np.int32), lcc2=lcc2[planes].astype(np.float32),
             lcc64_max=np.array(lcc.max()), argmax=np.array(np.unravel_index(np.argmax(lcc), lcc.shape)),
             norm_factor=np.array(float(c._norm_factor)), rmax=np.array(int(c._rmax)),
             lcc_mask=np.packbits(c._lcc_mask[planes].astype(bool)))
    np.savez_compressed(os.path.join(HERE, "scan_config4_256_subset.npz"), **d)
    print("scan_config4_256_subset R=%d lcc max %.4f at" % (len(rotations), lcc.max()), d["argmax"])


def scan_192():
    """BASELINE config 5 shape (subset): 192^3, plain LCC, 6 rotations of the 4.71 deg set + 2 true poses;
    eight z planes of the result grids are stored."""
    full = rotation_set(4.71)
    idx = np.unique(np.r_[np.arange(0, len(full), len(full) // 5)[:5], len(full) - 1])
    case = synth.config5(seed=0)
    target, template, mask = f32(case.target), f32(case.template), f32(case.mask)
    rotations = np.concatenate([full[idx[:3]], case.poses[1][0][None], full[idx[3:]], case.poses[5][0][None]])
    lcc, rot, lcc2, c = run_reference_scan(target, template, mask, rotations, False)
    planes = np.array([0, 1, 50, 96, 97, 150, 190, 191])
    d = dict(rotations=rotations, laplace=np.array(0), planes=planes, seed=np.array(0), subset_index=idx,
             lcc=lcc[planes].astype(np.float32), rot=rot[planes].astype(np.int32), lcc2=lcc2[planes].astype(np.float32),
             lcc64_max=np.array(lcc.max()), argmax=np.array(np.unravel_index(np.argmax(lcc), lcc.shape)),
             norm_factor=np.array(float(c._norm_factor)), rmax=np.array(int(c._rmax)),
             lcc_mask=np.packbits(c._lcc_mask[planes].astype(bool)))
    np.savez_compressed(os.path.join(HERE, "scan_config5_192_subset.npz"), **d)
    print("scan_config5_192_subset R=%d lcc max %.4f at" % (len(rotations), lcc.max()), d["argmax"])


def lcc_chain():
    """tests/test_powerfitter.py:67-79 -- perfect fit gives LCC 1 at index 0."""
    rng = np.random.default_rng(2)
    target = f32(rng.random((5, 6, 7)))
    c = CPUCorrelator(target)
    c._lcc_mask.fill(1)
    c.template = target.copy()
    c.mask = np.ones(target.shape)
    c._rot_template[:] = c._template
    c._rot_mask[:] = c._mask
    c._get_lcc()
    np.savez_compressed(os.path.join(HERE, "lcc_chain.npz"), target=target, lcc_scan=c._lcc_scan)
    print("lcc_chain max", c._lcc_scan.max(), "argmax", c._lcc_scan.argmax())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "256":
        scan_256()
    elif len(sys.argv) > 1 and sys.argv[1] == "192":
        scan_192()
    elif len(sys.argv) > 1 and sys.argv[1] == "mixed":
        scans_mixed()
    else:
        rotate_vectors()
        lcc_chain()
        scans()
        scans_mixed()
        scan_256()
        scan_192()
