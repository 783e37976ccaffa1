"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, FP64) of the reference's exhaustive LCC search path, the
checker for the CUDA product in ``powerfit_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product path never does (it fails loudly when its
CUDA library is missing).

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here against
(a) the reference's own known-answer tests for this path
(``tests/test_extensions.py:10-37``, ``tests/test_powerfitter.py:43-79``,
``tests/test__powerfit.py:13-42``) and (b) golden vectors produced by importing the
real reference (``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` is mounted), including full multi-rotation scans.

What is restated, with the reference lines it follows (paths below
``/root/reference/src/powerfit_em/``):

=========================  ======================================================
``rotate_grid3d``          ``_extensions.c:7-196`` (C restatement in
                           ``oracle/rotate_oracle.c``; pure-numpy twin below)
``conj_multiply``          ``_powerfit.pyx:45-53``
``calc_lcc``               ``_powerfit.pyx:56-72``
``laplace_wrap``           ``powerfitter.py:212-215`` -> scipy.ndimage.laplace(mode='wrap')
``OracleCorrelator``       ``powerfitter.py:166-393`` (BaseCorrelator + CPUCorrelator)
``partition_rotations``    ``powerfitter.py:95-108``
``combine_partials``       ``powerfitter.py:146-163``
``blur_points``            ``_powerfit.pyx:75-138``  (next row N2; pinned by
``dilate_points``          ``_powerfit.pyx:141-206``  tests/golden/shapes.npz)
``determine_core_indices`` ``helpers.py:26-34``
``structure_to_shape_like````volume.py:192-224``
``lower_resolution``       ``volume.py:129-141`` (next row N4; pinned by
``resample``               ``volume.py:66-72``    tests/golden/pyramid.npz)
``watershed_positions``    ``analyzer.py:80-95`` (next row N1; pinned by
``solution_rows``          ``analyzer.py:58-78``  tests/golden/analyzer_solutions.npz)
=========================  ======================================================

The FFT arithmetic itself is third-party in the reference (pyFFTW>=0.12 wrapping
FFTW3, ``pyproject.toml:26``; or its built-in fallback ``numpy.fft``,
``powerfitter.py:13,311-315``).  pyFFTW is not installable here, so -- exactly like
the reference in this image -- the oracle calls ``numpy.fft.rfftn/irfftn``.
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF_EXT = None


# --------------------------------------------------------------------------- #
# native helpers: the C restatement and, when built, the reference's own module
# --------------------------------------------------------------------------- #
def _load_c():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.pfo_rotate_grid3d.argtypes = [dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, dp,
                                          ctypes.c_int, dp, ctypes.c_long, ctypes.c_long,
                                          ctypes.c_long, ctypes.c_int]
        lib.pfo_rotate_grid3d.restype = None
        _LIB = lib
    return _LIB


def load_reference_extension():
    """The reference's own compiled ``_extensions`` module from ``oracle/_ref``
    (built by ``make -C oracle ref``), or None when it was not built."""
    global _REF_EXT
    if _REF_EXT is None:
        hits = glob.glob(os.path.join(_HERE, "_ref", "_extensions*.so"))
        if not hits:
            return None
        spec = importlib.util.spec_from_file_location("_extensions", hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF_EXT = mod
    return _REF_EXT


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------- #
# K1  rotate_grid3d
# --------------------------------------------------------------------------- #
def rotate_grid3d_numpy(grid, rotmat, radius, out, nearest):
    """Vectorised numpy twin of ``rotate_oracle.c`` (same evaluation order per
    voxel; ``_extensions.c:55-181``).  Used to cross-check the C restatement."""
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    R = np.asarray(rotmat, dtype=np.float64).reshape(9)
    nz, ny, nx = grid.shape
    onz, ony, onx = out.shape
    r = np.arange(-radius, radius + 1)
    z, y, x = np.meshgrid(r, r, r, indexing="ij")
    sel = (z * z + y * y + x * x) <= radius * radius
    z, y, x = z[sel], y[sel], x[sel]          # C order == reference visiting order
    sx = (R[6] * z + R[3] * y) + R[0] * x
    sy = (R[7] * z + R[4] * y) + R[1] * x
    sz = (R[8] * z + R[5] * y) + R[2] * x
    wrap = lambda i, n: np.where(i < 0, i + n, i)
    dst = (wrap(z, onz) * ony + wrap(y, ony)) * onx + wrap(x, onx)
    flat = grid.reshape(-1)

    def src(k, j, i):
        return flat[(wrap(k, nz) * ny + wrap(j, ny)) * nx + wrap(i, nx)]

    if nearest:
        rnd = lambda a: (np.sign(a) * np.floor(np.abs(a) + 0.5)).astype(np.int64)
        val = src(rnd(sz), rnd(sy), rnd(sx))
    else:
        i0 = np.floor(sx).astype(np.int64)
        j0 = np.floor(sy).astype(np.int64)
        k0 = np.floor(sz).astype(np.int64)
        wx, wy, wz = sx - i0, sy - j0, sz - k0
        wx1, wy1, wz1 = 1 - wx, 1 - wy, 1 - wz
        c00 = src(k0, j0, i0) * wx1 + src(k0, j0, i0 + 1) * wx
        c10 = src(k0, j0 + 1, i0) * wx1 + src(k0, j0 + 1, i0 + 1) * wx
        c01 = src(k0 + 1, j0, i0) * wx1 + src(k0 + 1, j0, i0 + 1) * wx
        c11 = src(k0 + 1, j0 + 1, i0) * wx1 + src(k0 + 1, j0 + 1, i0 + 1) * wx
        c0 = c00 * wy1 + c10 * wy
        c1 = c01 * wy1 + c11 * wy
        val = c0 * wz1 + c1 * wz
    # duplicates (the +/-radius alias) must resolve to the LAST visited write
    o = out.reshape(-1)
    o[dst] = val      # numpy fancy assignment keeps the last value for repeated indices
    return out


def rotate_grid3d(grid, rotmat, radius, out, nearest=False):
    """``_extensions.c:7-196``: out(r) = interp(grid, R^T r) for |r| <= radius."""
    lib = _load_c()
    if lib is None:
        return rotate_grid3d_numpy(grid, rotmat, radius, out, nearest)
    g = np.ascontiguousarray(grid, dtype=np.float64)
    R = np.ascontiguousarray(rotmat, dtype=np.float64).reshape(9)
    assert out.dtype == np.float64 and out.flags.c_contiguous
    lib.pfo_rotate_grid3d(_dptr(g), *g.shape, _dptr(R), int(radius), _dptr(out), *out.shape,
                          int(bool(nearest)))
    return out


# --------------------------------------------------------------------------- #
# K6 / K8
# --------------------------------------------------------------------------- #
def conj_multiply(in1, in2, out):
    """``_powerfit.pyx:45-53``: out = conj(in1) * in2 (1-D complex128)."""
    np.multiply(np.conj(in1), in2, out=out)
    return out


def calc_lcc(gcc, ave, ave2, mask, lcc):
    """``_powerfit.pyx:56-72``: lcc = gcc / sqrt(ave2 - ave**2) where mask; no guard."""
    ind = mask != 0
    with np.errstate(invalid="ignore", divide="ignore"):
        lcc[ind] = gcc[ind] / np.sqrt(ave2[ind] - ave[ind] ** 2)
    return lcc


def laplace_wrap(a):
    """``powerfitter.py:212-215``: scipy.ndimage.laplace(a, mode='wrap') == periodic
    6-neighbour second difference, accumulated axis by axis like scipy does."""
    out = np.zeros_like(a, dtype=np.float64)
    for ax in range(a.ndim):
        out += np.roll(a, 1, ax) + np.roll(a, -1, ax) - 2.0 * a
    return out


# --------------------------------------------------------------------------- #
# L2  BaseCorrelator + CPUCorrelator
# --------------------------------------------------------------------------- #
class OracleCorrelator:
    """``powerfitter.py:166-393`` restated.  Same setter order and errors."""

    def __init__(self, target, laplace=False, rotate=None):
        target = np.asarray(target, dtype=np.float64)
        self._target = target / target.max()                                   # :170
        self._laplace = laplace
        self._lcc_mask = (self._target > self._target.max() * 0.05).astype(np.uint8)   # :178-180
        self._rmax = min(target.shape) // 2                                    # :176
        self._template = self._mask = self._rotations = None
        self._rotate = rotate or rotate_grid3d
        t = laplace_wrap(self._target) if laplace else self._target           # :268-270
        self._ft_target = np.fft.rfftn(t)                                      # :276
        self._ft_target2 = np.fft.rfftn(t ** 2)                                # :277
        shape = target.shape
        self._rot_template = np.zeros(shape)
        self._rot_mask = np.zeros(shape)
        self._lcc_scan = np.zeros(shape)
        self._lcc = np.zeros(shape)
        self._rot = np.zeros(shape)

    # -- setters ------------------------------------------------------------
    @property
    def template(self):
        return self._template

    @template.setter
    def template(self, template):                                              # :236-243
        if template.shape != self._target.shape:
            raise ValueError("Shape of template does not match the target.")
        self._mask = None
        self._template = np.array(template, dtype=np.float64)

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, mask):                                                      # :190-210
        if self._template is None:
            raise ValueError("First set the template.")
        if self._target.shape != mask.shape:
            raise ValueError("Shape of the mask is different from target.")
        ind = mask != 0
        self._norm_factor = ind.sum()
        if self._norm_factor == 0:
            raise ValueError("Zero-filled mask is not allowed.")
        self._mask = np.array(mask, dtype=np.float64)
        if self._laplace:
            self._template = laplace_wrap(self._template)
        self._template *= self._mask
        self._template[ind] -= self._template[ind].mean()                      # :217-220
        self._template[ind] /= self._template[ind].std()
        self._template *= self._mask

    @property
    def rotations(self):
        return self._rotations

    @rotations.setter
    def rotations(self, rotations):                                            # :226-230
        self._rotations = np.asarray(rotations, dtype=np.float64).reshape(-1, 3, 3)

    @property
    def lcc(self):
        return self._lcc

    @property
    def rot(self):
        return self._rot

    # -- hot loop -----------------------------------------------------------
    def _translational_scan(self, rotmat):                                     # :335-373
        self._rotate(self._template, rotmat, self._rmax, self._rot_template, False)
        self._rotate(self._mask, rotmat, self._rmax, self._rot_mask, True)
        rot_mask2 = self._rot_mask * self._rot_mask
        ft_t = np.fft.rfftn(self._rot_template)
        ft_m = np.fft.rfftn(self._rot_mask)
        ft_m2 = np.fft.rfftn(rot_mask2)
        s = self._target.shape
        ax = (0, 1, 2)
        self._gcc = np.fft.irfftn(np.conj(ft_t) * self._ft_target, s=s, axes=ax)
        self._ave = np.fft.irfftn(np.conj(ft_m) * self._ft_target, s=s, axes=ax)
        self._ave2 = np.fft.irfftn(np.conj(ft_m2) * self._ft_target2, s=s, axes=ax)
        self._ave2 *= self._norm_factor                                        # :369
        calc_lcc(self._gcc.ravel(), self._ave.ravel(), self._ave2.ravel(),
                 self._lcc_mask.ravel(), self._lcc_scan.reshape(-1))
        return self._lcc_scan

    def scan(self, track_second=False, progress=None):                         # :317-333
        if any(req is None for req in (self._template, self._mask, self._rotations)):
            raise ValueError("First set the template, mask, and rotations.")
        self._lcc.fill(0)
        self._rot.fill(0)
        if track_second:
            # harness extra (not in the reference): runner-up LCC per voxel, so tests
            # can tell where the arg-max is decided by less than the tolerance.
            self._lcc2 = np.full(self._target.shape, -np.inf)
        for n in range(self._rotations.shape[0]):
            scan = self._translational_scan(self._rotations[n])
            ind = scan > self._lcc                                             # :327 strict >
            if track_second:
                # candidates that lose, or the dethroned best, feed the runner-up
                cand = np.where(ind, self._lcc, scan)
                np.fmax(self._lcc2, cand, out=self._lcc2)
            self._lcc[ind] = scan[ind]
            self._rot[ind] = n
            if progress is not None:
                progress(n)


# --------------------------------------------------------------------------- #
# L3  rotation sharding + merge
# --------------------------------------------------------------------------- #
def partition_rotations(nrot, njobs):
    """``powerfitter.py:95-108``: contiguous blocks of nrot//njobs, last takes the rest."""
    per = nrot // njobs
    return [(n * per, nrot if n == njobs - 1 else (n + 1) * per) for n in range(njobs)]


def combine_partials(parts, nrot_per_job, shape):
    """``powerfitter.py:146-163``: strict '>' merge in job order, index offset per job."""
    lcc = np.zeros(shape)
    rot = np.zeros(shape)
    for n, (plcc, prot) in enumerate(parts):
        ind = plcc > lcc
        lcc[ind] = plcc[ind]
        rot[ind] = prot[ind] + nrot_per_job * n
    return lcc, rot


def _scan_block(args):
    target, template, mask, rotations, laplace = args
    c = OracleCorrelator(target, laplace=laplace)
    c.template = template
    c.mask = mask
    c.rotations = rotations
    c.scan()
    return c.lcc, c.rot


def parallel_scan(target, template, mask, rotations, laplace=False, nproc=1):
    """``PowerFitter._cpu_scan`` (``powerfitter.py:94-163``) with a process pool in
    place of fork + .npy files; same partition, same merge."""
    import multiprocessing as mp
    rotations = np.asarray(rotations, dtype=np.float64).reshape(-1, 3, 3)
    blocks = partition_rotations(rotations.shape[0], nproc)
    jobs = [(target, template, mask, rotations[a:b], laplace) for a, b in blocks]
    if nproc == 1:
        parts = [_scan_block(jobs[0])]
    else:
        with mp.get_context("fork").Pool(nproc) as pool:
            parts = pool.map(_scan_block, jobs)
    return combine_partials(parts, rotations.shape[0] // nproc, np.asarray(target).shape)


# --------------------------------------------------------------------------- #
# Solution extraction -- dense restatement of the reference Analyzer
def watershed_positions(corr, steps=5):
    """``analyzer.py:80-95``: for ``steps`` cutoffs from the maximum down to half of it, label
    ``corr >= cutoff`` (scipy default 6-connectivity) and collect the position of the maximum
    of every feature."""
    from scipy.ndimage import label, maximum_position
    top = corr.max()
    low = 0.5 * top
    delta = (top - low) / steps
    level = top
    found = set()
    for _ in range(steps):
        level = level - delta
        lab, count = label(corr >= level)
        found.update(tuple(int(c) for c in p) for p in maximum_position(corr, lab, list(range(1, count + 1))))
    return found


def solution_rows(corr, rotmat, rotmat_ind, positions, voxelspacing=1, origin=(0, 0, 0), z_sigma=1):
    """``analyzer.py:58-78``: [cc, Fisher z, relative z, x, y, z, a11..a33] per position,
    sorted by cc descending."""
    rows = []
    for pos in positions:
        cc = corr[pos]
        fz = 0.5 * (np.log(1 + cc) - np.log(1 - cc))
        zyx = [c * voxelspacing + o for c, o in zip(pos, origin[::-1])]
        rows.append([cc, fz, fz / z_sigma, zyx[2], zyx[1], zyx[0]] + list(np.ravel(rotmat[int(rotmat_ind[pos])])))
    return sorted(rows, key=lambda r: r[0], reverse=True)


# --------------------------------------------------------------------------- #
# next row N2: template / mask synthesis from atom coordinates
# --------------------------------------------------------------------------- #
def _axis_range(p, reach, n):
    """The reference's clipped loop range along one axis (wraparound=True)."""
    lo = max(int(np.ceil(p - reach)), -n + 1)
    hi = min(int(np.floor(p + reach)), n - 1)
    return np.arange(lo, hi + 1)


def blur_points(points, weights, sigma, out, wraparound=True):
    """_powerfit.pyx:75-138: out[z,y,x] += w_n exp(-d^2 / (2 sigma^2)) within 4 sigma, atom by atom;
    negative positions index from the end like Python (positions -n+1 .. n-1)."""
    assert wraparound
    nz, ny, nx = out.shape
    extend = 4.0 * sigma
    extend2 = extend * extend
    dsigma2 = 2.0 * sigma * sigma
    for n in range(points.shape[1]):
        xs = _axis_range(points[0, n], extend, nx)
        ys = _axis_range(points[1, n], extend, ny)
        zs = _axis_range(points[2, n], extend, nz)
        if min(len(xs), len(ys), len(zs)) == 0:
            continue
        z2 = (zs - points[2, n]) ** 2
        y2z2 = (ys - points[1, n])[None, :] ** 2 + z2[:, None]
        d2 = (xs - points[0, n])[None, None, :] ** 2 + y2z2[:, :, None]
        val = np.where(d2 <= extend2, weights[n] * np.exp(-d2 / dsigma2), 0.0)
        # a box wider than the grid visits a voxel twice (p and p - n): add.at keeps both, in loop order
        np.add.at(out, np.ix_(zs % nz, ys % ny, xs % nx), val)


def dilate_points(points, radii, out, wraparound=True):
    """_powerfit.pyx:141-206: out = 1 inside the ball of radius radii[n] around every atom."""
    assert wraparound
    nz, ny, nx = out.shape
    for n in range(points.shape[1]):
        radius = radii[n]
        xs = _axis_range(points[0, n], radius, nx)
        ys = _axis_range(points[1, n], radius, ny)
        zs = _axis_range(points[2, n], radius, nz)
        if min(len(xs), len(ys), len(zs)) == 0:
            continue
        z2 = (zs - points[2, n]) ** 2
        y2z2 = (ys - points[1, n])[None, :] ** 2 + z2[:, None]
        d2 = (xs - points[0, n])[None, None, :] ** 2 + y2z2[:, :, None]
        np.maximum.at(out, np.ix_(zs % nz, ys % ny, xs % nx), (d2 <= radius ** 2).astype(np.float64))


def determine_core_indices(mask):
    """helpers.py:26-34."""
    from scipy.ndimage import binary_erosion
    core = np.zeros(mask.shape)
    eroded = mask > 0
    while eroded.sum() > 0:
        core += eroded
        eroded = binary_erosion(eroded)
    return core


def structure_to_shape_like(shape, voxelspacing, origin, xyz, resolution, weights=None, radii=None, kind="vol"):
    """volume.py:192-224 on plain arrays: grid of `shape` with the given voxel spacing and origin."""
    natoms = xyz.shape[1]
    if kind == "vol" and weights is None:
        weights = np.ones(natoms)
    if kind == "mask":
        if radii is None:
            radii = np.empty(natoms, dtype=np.float64)
            radii.fill(5)
        radii = radii / voxelspacing
    sigma = (resolution / (np.sqrt(2.0) * np.pi)) / voxelspacing
    xyz_grid = xyz - np.asarray(origin, dtype=np.float64).reshape(3, 1)
    xyz_grid = xyz_grid / voxelspacing
    out = np.zeros(shape)
    if kind == "vol":
        blur_points(xyz_grid, np.asarray(weights, dtype=np.float64), sigma, out, True)
    else:
        dilate_points(xyz_grid, radii, out, True)
    return out


# --------------------------------------------------------------------------- #
# next row N4: image pyramid
# --------------------------------------------------------------------------- #
def lower_resolution(array, voxelspacing, res_high, res_low):
    """volume.py:129-141."""
    from scipy.ndimage import gaussian_filter
    r2s = lambda r: r / (np.sqrt(2.0) * np.pi)
    sigma_k = np.sqrt(r2s(res_low) ** 2 - r2s(res_high) ** 2) / voxelspacing
    return gaussian_filter(array, sigma_k, mode="constant")


def resample(array, voxelspacing, factor, order=1):
    """volume.py:66-72."""
    import warnings
    from scipy.ndimage import zoom
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = zoom(array, factor, order=order)
    return out, voxelspacing / factor
