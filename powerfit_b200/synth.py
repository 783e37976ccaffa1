"""Deterministic synthetic inputs of the shapes BASELINE.json names.

There is no network and no EM data in the image, so benchmarks, smoke tests and the
golden-vector generator all draw their inputs from here: a seeded random-walk
C-alpha trace stands in for the PDB model, a Gaussian splat of several posed copies
plus noise stands in for the cryo-EM map, and the template / mask pair is built the
way the reference CLI builds it (density of the model centred on voxel (0,0,0) with
periodic wrap-around, mask = union of 5 A... balls around the atoms, optionally
core-weighted; cf. /root/reference/src/powerfit_em/powerfit.py:245-267).  The code is
an independent numpy implementation -- only the *shape* of the data matters for the
search kernels.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Case:
    target: np.ndarray      # (nz,ny,nx) float64 map
    template: np.ndarray    # (nz,ny,nx) float64, centred on voxel 0 (wrapped)
    mask: np.ndarray        # (nz,ny,nx) float64, binary or core-weighted
    voxelspacing: float
    resolution: float
    poses: list             # [(rotmat, shift_voxels)] ground truth of the copies
    name: str = ""


def random_rotations(n, seed=0, decimals=8):
    """n seeded uniform rotation matrices (unit quaternions -> matrices, rounded to 8
    decimals like the reference's quat_to_rotmat, rotations.py:30-61)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    if n:
        R[0] = np.eye(3)
    return np.around(R, decimals=decimals)


def random_walk_trace(n_res, rg, seed):
    """Compact random-walk C-alpha trace (Angstrom), centred, scaled to radius of
    gyration ``rg``."""
    rng = np.random.default_rng(seed)
    steps = rng.normal(size=(n_res, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    # mild pull towards the centre keeps the walk globular
    xyz = np.zeros((n_res, 3))
    for i in range(1, n_res):
        xyz[i] = xyz[i - 1] + 3.8 * steps[i] - 0.02 * xyz[i - 1]
    xyz -= xyz.mean(0)
    xyz *= rg / np.sqrt((xyz ** 2).sum(1).mean())
    return xyz


def splat_gaussians(xyz_vox, sigma_vox, shape, weights=None):
    """Periodic Gaussian splat of points given in voxel units (x,y,z order)."""
    nz, ny, nx = shape
    out = np.zeros(shape)
    cut = int(np.ceil(4 * sigma_vox))
    r = np.arange(-cut, cut + 1)
    if weights is None:
        weights = np.ones(len(xyz_vox))
    for (x, y, z), w in zip(xyz_vox, weights):
        ix, iy, iz = int(np.rint(x)), int(np.rint(y)), int(np.rint(z))
        gx = np.exp(-0.5 * ((ix + r - x) / sigma_vox) ** 2)
        gy = np.exp(-0.5 * ((iy + r - y) / sigma_vox) ** 2)
        gz = np.exp(-0.5 * ((iz + r - z) / sigma_vox) ** 2)
        blk = w * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
        out[np.ix_((iz + r) % nz, (iy + r) % ny, (ix + r) % nx)] += blk
    return out


def splat_balls(xyz_vox, radius_vox, shape):
    """Periodic union of balls (binary mask) around points in voxel units."""
    nz, ny, nx = shape
    out = np.zeros(shape)
    cut = int(np.ceil(radius_vox))
    r = np.arange(-cut, cut + 1)
    for x, y, z in xyz_vox:
        ix, iy, iz = int(np.rint(x)), int(np.rint(y)), int(np.rint(z))
        d2 = ((iz + r - z) ** 2)[:, None, None] + ((iy + r - y) ** 2)[None, :, None] \
            + ((ix + r - x) ** 2)[None, None, :]
        sub = np.ix_((iz + r) % nz, (iy + r) % ny, (ix + r) % nx)
        out[sub] = np.maximum(out[sub], (d2 <= radius_vox ** 2).astype(np.float64))
    return out


def core_weight(mask):
    """Erosion depth of every mask voxel (1 on the surface layer, 2 one voxel in, ...),
    the quantity the reference's ``-cw`` option multiplies into the mask
    (helpers.py:26-34)."""
    from scipy.ndimage import binary_erosion
    depth = np.zeros(mask.shape)
    cur = mask > 0
    while cur.any():
        depth += cur
        cur = binary_erosion(cur)
    return depth


def make_template(shape, voxelspacing, resolution, n_res, rg, seed, core_weighted=False):
    """Template density and mask of a seeded random-walk model, centred on voxel 0 (what make_case builds)."""
    shape = tuple(shape)
    xyz = random_walk_trace(n_res, rg, seed)
    sigma_vox = max(resolution / (np.sqrt(2.0) * np.pi) / voxelspacing, 0.6)
    xv = xyz / voxelspacing
    template = splat_gaussians(xv, sigma_vox, shape)
    mask = splat_balls(xv, max(5.0 / voxelspacing, 1.5), shape)
    if core_weighted:
        mask = core_weight(mask)
    return template, mask


def make_case(n=64, voxelspacing=2.0, resolution=8.0, n_res=300, rg=14.0, n_copies=3,
              seed=0, noise=0.05, core_weighted=False, shape=None, name=""):
    """A map with ``n_copies`` posed copies of a random-walk model, and the model's
    template/mask pair centred on voxel 0."""
    shape = tuple(shape) if shape is not None else (n, n, n)
    rng = np.random.default_rng(seed + 1000)
    xyz = random_walk_trace(n_res, rg, seed)
    sigma_vox = resolution / (np.sqrt(2.0) * np.pi) / voxelspacing
    sigma_vox = max(sigma_vox, 0.6)
    xv = xyz / voxelspacing
    extent = np.sqrt((xv ** 2).sum(1)).max()
    template = splat_gaussians(xv, sigma_vox, shape)
    mask = splat_balls(xv, max(5.0 / voxelspacing, 1.5), shape)
    if core_weighted:
        mask = core_weight(mask)
    target = np.zeros(shape)
    poses = []
    rots = random_rotations(n_copies + 1, seed=seed + 7)[1:]
    dims = np.array(shape[::-1], dtype=np.float64)          # x,y,z extents
    for R in rots:
        lim = np.maximum(dims / 2 - extent - 4 * sigma_vox - 1, 1.0)
        shift = dims / 2 + rng.uniform(-0.6, 0.6, 3) * lim
        pts = xv @ R.T + shift
        target += splat_gaussians(pts, sigma_vox, shape)
        poses.append((R, shift))
    target += noise * target.max() * rng.normal(size=shape)
    return Case(target=target, template=template, mask=mask, voxelspacing=voxelspacing,
                resolution=resolution, poses=poses, name=name)


# named configurations (BASELINE.json:configs; sizes from SURVEY.md section 8)
def config1(seed=0):
    return make_case(n=64, voxelspacing=2.0, resolution=8.0, n_res=300, rg=14.0, n_copies=3,
                     seed=seed, name="64^3 8A 300-res 20deg")


def config2(seed=0, core_weighted=False, n=128):
    # GroEL/GroES-like: 23.5 A map, ~2.8 A voxels, GroES-sized (679 res) template
    return make_case(n=n, voxelspacing=2.8, resolution=23.5, n_res=679, rg=26.0, n_copies=6,
                     seed=seed, core_weighted=core_weighted, name="%d^3 23.5A GroES-sized 10deg" % n)


def config4(seed=0, n=256):
    return make_case(n=n, voxelspacing=1.5, resolution=6.0, n_res=2500, rg=38.0, n_copies=10,
                     seed=seed, core_weighted=True, name="%d^3 6A ribosome-sized 4.71deg" % n)


def config5(seed=0, n=192):
    # one of the four sub-unit templates of BASELINE configs[4]: 192^3 map, plain LCC
    return make_case(n=n, voxelspacing=2.0, resolution=8.0, n_res=1200, rg=30.0, n_copies=8,
                     seed=seed, name="%d^3 8A sub-unit, fine search" % n)
