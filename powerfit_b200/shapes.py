"""Template and mask synthesis on the device.

Host-side mirror of what the reference CLI calls between reading the PDB model and building
the ``PowerFitter`` (/root/reference/src/powerfit_em/powerfit.py:245-267):
``structure_to_shape_like`` (volume.py:192-224, over ``_powerfit.blur_points`` /
``dilate_points``, _powerfit.pyx:75-206) and ``determine_core_indices`` (helpers.py:26-34).
Same names, arguments and error behaviour; the grids are computed by the FP64 kernels in
``csrc/shapes.cu`` and returned as host float64 arrays (what ``Volume.array`` holds).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _grid_of(vol):
    """(shape, voxelspacing, origin) of a Volume-like object (``.array``/``.shape``, ``.voxelspacing``,
    ``.origin``) or of a (shape, voxelspacing, origin) tuple."""
    if isinstance(vol, tuple):
        shape, vs, origin = vol
    else:
        shape = vol.array.shape if hasattr(vol, "array") else vol.shape
        vs, origin = vol.voxelspacing, vol.origin
    return tuple(int(v) for v in shape), float(vs), np.asarray(origin, dtype=np.float64).reshape(3, 1)


def _device(device):
    import torch
    if not torch.cuda.is_available():
        raise _lib.PowerfitB200Error("powerfit_b200.shapes needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cuda", device) if isinstance(device, int) else torch.device(device)


def structure_to_shape_like(vol, xyz, resolution=None, weights=None, radii=None, shape="vol", device=None):
    """volume.py:192-224.  ``xyz`` is the reference's (3, natoms) coordinate array (x, y, z rows, Angstrom).
    ``shape='vol'``: Gaussian density of the atoms at ``resolution``; ``shape='mask'``: union of balls of
    ``radii`` (default 5 Angstrom).  Returns the (nz, ny, nx) float64 grid."""
    import torch
    lib = _lib.load()
    gshape, vs, origin = _grid_of(vol)
    xyz = np.asarray(xyz, dtype=np.float64)
    natoms = xyz.shape[1]
    if resolution is None:
        resolution = vol.resolution
    if shape == "vol":
        if weights is None:
            weights = np.ones(natoms)
        elif np.asarray(weights).size != natoms:
            raise ValueError("weights array is of incorrect size")
    if shape == "mask":
        if radii is None:
            radii = np.empty(natoms, dtype=np.float64)
            radii.fill(5)
        elif np.asarray(radii).size != natoms:
            raise ValueError("weights array is of incorrect size")
        radii = np.asarray(radii, dtype=np.float64) / vs
    sigma = (resolution / (np.sqrt(2.0) * np.pi)) / vs
    xyz_grid = xyz - origin                      # move the coordinates to the origin of the grid
    xyz_grid /= vs
    dev = _device(device)
    nz, ny, nx = gshape
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_pts = torch.from_numpy(np.ascontiguousarray(xyz_grid)).to(dev)
        out = torch.zeros(gshape, dtype=torch.float64, device=dev)
        if shape == "vol":
            d_w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
            _lib.check(lib.pfb_blur_points(d_pts.data_ptr(), d_w.data_ptr(), natoms, float(sigma), nz, ny, nx,
                                           out.data_ptr(), stream))
        elif shape == "mask":
            d_r = torch.from_numpy(np.ascontiguousarray(radii)).to(dev)
            _lib.check(lib.pfb_dilate_points(d_pts.data_ptr(), d_r.data_ptr(), natoms, nz, ny, nx,
                                             out.data_ptr(), stream))
        return out.cpu().numpy()


def determine_core_indices(mask, device=None):
    """helpers.py:26-34: erosion depth of every voxel of ``mask > 0`` (1 on the surface layer, 2 one voxel
    in, ...), the weights of the core-weighted LCC (``-cw``)."""
    import torch
    lib = _lib.load()
    mask = np.ascontiguousarray(mask, dtype=np.float64)
    dev = _device(device)
    nz, ny, nx = mask.shape
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_m = torch.from_numpy(mask).to(dev)
        core = torch.empty(mask.shape, dtype=torch.float64, device=dev)
        scratch = torch.empty(2 * mask.size + 16, dtype=torch.uint8, device=dev)
        _lib.check(lib.pfb_core_indices(d_m.data_ptr(), nz, ny, nx, core.data_ptr(), scratch.data_ptr(), stream))
        return core.cpu().numpy()
