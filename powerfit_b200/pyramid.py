"""Multi-scale image pyramid on the device.

Array-level mirror of the reference's ``image-pyramid`` script
(/root/reference/src/powerfit_em/scripts/__init__.py:93-103): ``lower_resolution``
(volume.py:129-141) and ``resample`` (volume.py:66-72) with the same arguments, on
``(array, voxelspacing)`` instead of ``Volume`` objects (file IO stays with the caller).  The
Gaussian filter reproduces scipy's FP64 result exactly, the linear zoom to the last bit or two.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .shapes import _device


def res_to_sigma(resolution):                       # volume.py:121-122
    return resolution / (np.sqrt(2.0) * np.pi)


def _gaussian_weights(sigma, truncate=4.0):
    """Right half of scipy.ndimage's normalised Gaussian kernel (radius = int(truncate * sigma + 0.5))."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:]), radius


def lower_resolution(array, voxelspacing, res_high, res_low, device=None):
    """volume.py:129-141: blur with the Gaussian that takes a map from res_high to res_low."""
    import torch
    lib = _lib.load()
    sigma_k = np.sqrt(res_to_sigma(res_low) ** 2 - res_to_sigma(res_high) ** 2) / voxelspacing
    a = np.ascontiguousarray(array, dtype=np.float64)
    if sigma_k <= 1e-15:
        # scipy.ndimage.gaussian_filter skips axes with sigma <= 1e-15: the map comes back unchanged
        return a.copy()
    w, radius = _gaussian_weights(sigma_k)
    dev = _device(device)
    nz, ny, nx = a.shape
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_in = torch.from_numpy(a).to(dev)
        d_w = torch.from_numpy(w).to(dev)
        out, tmp = torch.empty_like(d_in), torch.empty_like(d_in)
        _lib.check(lib.pfb_gaussian_filter(d_in.data_ptr(), out.data_ptr(), tmp.data_ptr(), nz, ny, nx,
                                           d_w.data_ptr(), radius, stream))
        return out.cpu().numpy()


def resample(array, voxelspacing, factor, order=1, device=None):
    """volume.py:66-72: scipy.ndimage.zoom(array, factor, order=1); returns (array, new voxelspacing)."""
    import torch
    if order != 1:
        raise ValueError("only order=1 (what the image pyramid uses) is implemented")
    lib = _lib.load()
    a = np.ascontiguousarray(array, dtype=np.float64)
    oshape = tuple(int(round(s * factor)) for s in a.shape)
    dev = _device(device)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_in = torch.from_numpy(a).to(dev)
        out = torch.empty(oshape, dtype=torch.float64, device=dev)
        _lib.check(lib.pfb_zoom_linear(d_in.data_ptr(), a.shape[0], a.shape[1], a.shape[2], out.data_ptr(),
                                       oshape[0], oshape[1], oshape[2], stream))
        return out.cpu().numpy(), voxelspacing / factor


def image_pyramid(array, voxelspacing, resolution, target_resolutions, resampling_rate=2):
    """scripts/__init__.py:93-103 without the file IO: [(array, voxelspacing)] per target resolution."""
    if resolution <= 0:
        raise ValueError("Resolution should be bigger than 0.")
    if resampling_rate < 1:
        raise ValueError("Resampling rate should be bigger than 1 times Nyquist.")
    levels = []
    for res in target_resolutions:
        if res < resolution:
            raise ValueError("Target resolution of image-pyramid should be lower than original data.")
        blurred = lower_resolution(array, voxelspacing, resolution, res)
        new_voxelspacing = res / (2 * resampling_rate)
        levels.append(resample(blurred, voxelspacing, voxelspacing / new_voxelspacing, order=1))
    return levels
